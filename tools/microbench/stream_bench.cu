// Streaming-read microbenchmark (sm_100a): how fast can one B200 pull a large array from HBM through
//   (a) cp.async.bulk (1-D bulk copies into a shared-memory ring, one producer thread per CTA),
//   (b) cp.async.bulk.tensor 2-D boxes of 64 x 64 bf16 with SWIZZLE_128B (the weight-gradient kernel's loads),
//   (c) plain 16-byte LDG from every thread?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bench stream_bench.cu -lcuda
// Nothing here is part of the product; it decides which load path the HBM-bound kernels should use.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    }
}

// (a) 1-D bulk copies: ring of `depth` stages of `stage_bytes`; a consumer warp "reads" one word and releases the stage
__global__ void bulk1d_kernel(const uint8_t* src, size_t total, uint32_t stage_bytes, int depth, unsigned long long* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)depth * stage_bytes);
    uint64_t* empty = full + depth;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t n_stage = total / stage_bytes;
    unsigned long long acc = 0;
    if (threadIdx.x == 0) {
        uint32_t s = 0, ph = 0;
        for (size_t i = blockIdx.x; i < n_stage; i += gridDim.x) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect(&full[s], stage_bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)s * stage_bytes)), "l"(src + i * stage_bytes), "r"(stage_bytes), "r"(smem_u32(&full[s])) : "memory");
            if (++s == (uint32_t)depth) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        uint32_t s = 0, ph = 0;
        for (size_t i = blockIdx.x; i < n_stage; i += gridDim.x) {
            mbar_wait(&full[s], ph);
            acc += *reinterpret_cast<volatile uint32_t*>(smem + (size_t)s * stage_bytes);
            mbar_arrive(&empty[s]);
            if (++s == (uint32_t)depth) { s = 0; ph ^= 1; }
        }
        if (acc == 0x1234567) *sink = acc;
    }
}

// (b) 2-D tensor boxes: tensor [rows, cols] bf16, box 64 cols x box_rows rows, `boxes` column boxes per stage
__global__ void tma2d_kernel(const __grid_constant__ CUtensorMap map, int rows, int col_boxes, int box_rows, int depth, unsigned long long* sink) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t box_bytes = (uint32_t)box_rows * 128u, stage_bytes = box_bytes * col_boxes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)depth * stage_bytes);
    uint64_t* empty = full + depth;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_stage = rows / box_rows;
    unsigned long long acc = 0;
    if (threadIdx.x == 0) {
        uint32_t s = 0, ph = 0;
        // contiguous row range per CTA (like the weight-gradient kernel's row splits)
        const int per = (n_stage + gridDim.x - 1) / gridDim.x;
        const int i0 = blockIdx.x * per, i1 = min(n_stage, i0 + per);
        for (int i = i0; i < i1; ++i) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect(&full[s], stage_bytes);
            for (int c = 0; c < col_boxes; ++c)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem + (size_t)s * stage_bytes + c * box_bytes)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&full[s])), "r"(c * 64), "r"(i * box_rows) : "memory");
            if (++s == (uint32_t)depth) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        uint32_t s = 0, ph = 0;
        const int per = (n_stage + gridDim.x - 1) / gridDim.x;
        const int i0 = blockIdx.x * per, i1 = min(n_stage, i0 + per);
        for (int i = i0; i < i1; ++i) {
            mbar_wait(&full[s], ph);
            acc += *reinterpret_cast<volatile uint32_t*>(smem + (size_t)s * stage_bytes);
            mbar_arrive(&empty[s]);
            if (++s == (uint32_t)depth) { s = 0; ph ^= 1; }
        }
        if (acc == 0x1234567) *sink = acc;
    }
}

// (c) LDG.128 from every thread, `unroll` independent loads in flight per thread
template <int UNROLL>
__global__ void ldg_kernel(const uint4* src, size_t n16, unsigned long long* sink) {
    unsigned long long acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x1234567) *sink = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <class F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const size_t total = (size_t)1 << 30;   // 1 GiB >> L2
    uint8_t* d; unsigned long long* sink;
    CK(cudaMalloc(&d, total)); CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(d, 1, total));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("{\"sms\": %d}\n", sms);
    // (c) LDG
    {
        float ms = time_ms([&] { ldg_kernel<8><<<sms * 8, 256>>>(reinterpret_cast<const uint4*>(d), total / 16, sink); });
        printf("{\"path\": \"ldg128\", \"unroll\": 8, \"ctas_per_sm\": 8, \"GBps\": %.0f}\n", total / ms / 1e6);
        ms = time_ms([&] { ldg_kernel<4><<<sms * 4, 512>>>(reinterpret_cast<const uint4*>(d), total / 16, sink); });
        printf("{\"path\": \"ldg128\", \"unroll\": 4, \"ctas_per_sm\": 4, \"threads\": 512, \"GBps\": %.0f}\n", total / ms / 1e6);
    }
    // (a) 1-D bulk
    CK(cudaFuncSetAttribute(bulk1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    const int cfgs[][3] = {{16384, 4, 1}, {16384, 12, 1}, {32768, 6, 1}, {65536, 3, 1}, {16384, 4, 3}, {8192, 8, 3}, {4096, 16, 3}, {16384, 2, 6}, {8192, 4, 6}};
    for (auto& c : cfgs) {
        const uint32_t sb = c[0]; const int depth = c[1], per_sm = c[2];
        const size_t smem = (size_t)sb * depth + 16 * depth + 64;
        float ms = time_ms([&] { bulk1d_kernel<<<sms * per_sm, 64, smem>>>(d, total, sb, depth, sink); });
        printf("{\"path\": \"bulk1d\", \"stage_bytes\": %u, \"depth\": %d, \"ctas_per_sm\": %d, \"in_flight_KB_per_sm\": %zu, \"GBps\": %.0f}\n", sb, depth, per_sm, (size_t)sb * depth * per_sm / 1024, total / ms / 1e6);
    }
    // (b) 2-D tensor boxes on a [rows, 256] bf16 tensor
    {
        void* fn = nullptr; cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
        CK(cudaFuncSetAttribute(tma2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        const int cols = 256; const int rows = (int)(total / (cols * 2));
        const int tcfg[][3] = {{64, 3, 1}, {64, 6, 1}, {32, 6, 1}, {128, 1, 1}, {256, 1, 1}, {64, 1, 3}, {32, 2, 3}};   // box_rows, depth, ctas/sm (4 column boxes per stage)
        for (auto& c : tcfg) {
            const int box_rows = c[0], depth = c[1], per_sm = c[2];
            CUtensorMap map;
            cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            const size_t stage = (size_t)box_rows * 128 * 4; const size_t smem = stage * depth + 16 * depth + 64 + 1024;
            if (smem > 220 * 1024) continue;
            float ms = time_ms([&] { tma2d_kernel<<<sms * per_sm, 64, smem>>>(map, rows, 4, box_rows, depth, sink); });
            printf("{\"path\": \"tma2d_sw128\", \"box_rows\": %d, \"stage_bytes\": %zu, \"depth\": %d, \"ctas_per_sm\": %d, \"in_flight_KB_per_sm\": %zu, \"GBps\": %.0f}\n", box_rows, stage, depth, per_sm, stage * depth * per_sm / 1024, total / ms / 1e6);
        }
    }
    return 0;
}
