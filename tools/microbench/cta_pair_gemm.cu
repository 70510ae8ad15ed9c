// cta_group::2 feasibility test (sm_100a): D[256 x 256] = A[256 x 64] * B[256 x 64]^T with ONE tcgen05.mma.cta_group::2
// sequence issued by the leader CTA of a 2-CTA cluster.  Each CTA stages its own 128 rows of A and its own 128 rows (N half)
// of B; the accumulator rows of CTA r end up in CTA r's TMEM.  Checks the result against the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cta_pair_gemm cta_pair_gemm.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try(b, parity)) if (clock64() - t0 > 2000000000LL) { printf("timeout block %d thread %d bar %p\n", blockIdx.x, threadIdx.x, b); __trap(); }
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* b, uint32_t cta) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(b)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// the same load, but the mbarrier that receives the complete_tx lives in CTA 0 of the pair (what CUTLASS's SM100_TMA_2SM_LOAD does)
__device__ __forceinline__ void tma_load_2d_leaderbar(const CUtensorMap* m, uint64_t* bar_local_addr, void* dst, int c0, int c1) {
    uint32_t rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(bar_local_addr)), "r"(0u));
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(rbar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

struct Smem {
    alignas(1024) uint8_t a[128 * 128];   // 128 rows x 64 bf16 (SW128)
    alignas(1024) uint8_t b[128 * 128];   // this CTA's 128 N-rows x 64 bf16
    uint64_t full, peer_ready, done;
    uint32_t tmem_base;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* C, int alloc_mode, int tma_mode) {
    extern __shared__ uint8_t raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
    const uint32_t rank = cluster_rank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&s.full, 1);
        mbar_init(&s.peer_ready, 1);
        mbar_init(&s.done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (alloc_mode == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(256u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(256u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;
    if (threadIdx.x == 0) {
        if (tma_mode == 1) {
            // all four tiles (two per CTA) complete on the LEADER's barrier; the peer arms nothing
            if (rank == 0) mbar_expect(&s.full, 4 * 128 * 128);
            tma_load_2d_leaderbar(&tmA, &s.full, s.a, 0, rank * 128);
            tma_load_2d_leaderbar(&tmB, &s.full, s.b, 0, rank * 128);
            if (rank == 0) mbar_wait(&s.full, 0);
        } else {
            mbar_expect(&s.full, 2 * 128 * 128);
            tma_load_2d(&tmA, &s.full, s.a, 0, rank * 128);
            tma_load_2d(&tmB, &s.full, s.b, 0, rank * 128);
            mbar_wait(&s.full, 0);
        }
        if (rank == 1) {
            if (tma_mode == 0) mbar_arrive_remote(&s.peer_ready, 0);       // tell the leader this CTA's operands have landed
        } else {
            if (tma_mode == 0) mbar_wait(&s.peer_ready, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = idesc_bf16(256, 256);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t ad = smem_desc(smem_u32(s.a) + k * 32, 16, 1024);
                const uint64_t bd = smem_desc(smem_u32(s.b) + k * 32, 16, 1024);
                const uint32_t acc = k != 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            }
            const uint16_t mask = 3;
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&s.done)), "h"(mask) : "memory");
        }
    }
    mbar_wait(&s.done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // 4 warps x 32 lanes = 128 rows of this CTA; 256 columns in 8 chunks of 32
    const int row = rank * 128 + warp * 32 + lane;
    for (int c = 0; c < 8; ++c) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c * 32;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                       "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) C[(size_t)row * 256 + c * 32 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 0) {
        if (alloc_mode == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int alloc_mode = argc > 1 ? atoi(argv[1]) : 2;
    const int tma_mode = argc > 2 ? atoi(argv[2]) : 0;
    const int M = 256, N = 256, K = 64;
    __nv_bfloat16 *hA = new __nv_bfloat16[M * K], *hB = new __nv_bfloat16[N * K];
    float* fa = new float[M * K]; float* fb = new float[N * K];
    srand(1);
    for (int i = 0; i < M * K; ++i) { float v = (rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(v); fa[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 4.0f; hB[i] = __float2bfloat16(v); fb[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dC;
    CK(cudaMalloc(&dA, M * K * 2)); CK(cudaMalloc(&dB, N * K * 2)); CK(cudaMalloc(&dC, M * N * 4));
    CK(cudaMemcpy(dA, hA, M * K * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, N * K * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0, M * N * 4));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
    CUtensorMap tmA, tmB;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, 128}; cuuint32_t estr[2] = {1, 1};
    if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode A failed\n"); return 1; }
    if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode B failed\n"); return 1; }
    const size_t smem = sizeof(Smem) + 1024;
    CK(cudaFuncSetAttribute(pair_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pair_gemm<<<2, 128, smem>>>(tmA, tmB, dC, alloc_mode, tma_mode);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    float* hC = new float[M * N];
    CK(cudaMemcpy(hC, dC, M * N * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)fa[m * K + k] * fb[n * K + k];
        double e = fabs(ref - hC[m * N + n]); if (e > maxerr) maxerr = e; if (e > 1e-3) ++bad;
    }
    printf("{\"test\": \"cta_pair_gemm\", \"tma_mode\": %d, \"alloc_mode\": %d, \"max_abs_err\": %.3g, \"bad\": %d, \"C[0]\": %g, \"C[last]\": %g}\n", tma_mode, alloc_mode, maxerr, bad, hC[0], hC[M * N - 1]);
    return bad != 0;
}
