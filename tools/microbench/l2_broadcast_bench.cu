// L2 re-streaming microbenchmark (sm_100a): the fused MLP kernel re-reads the SAME ~0.7 MB of weights from L2 in every
// CTA, 16 KB stage after 16 KB stage.  How fast can 148 CTAs pull an L2-resident working set through a shared-memory ring
// when (a) every CTA walks the same copy, in step, (b) CTAs walk the same copy with a per-CTA phase shift, (c) groups of
// CTAs have their own replica -- and what does a concurrent 1:1 stream of TMA stores to HBM cost?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_broadcast_bench l2_broadcast_bench.cu
// Nothing here is part of the product; it decides how the fused kernel should lay out / address its weights.
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

// every CTA streams `n_stage` stages of `stage_bytes` out of its replica (set_bytes long, walked cyclically from a per-CTA
// start offset) through a `depth`-deep ring; with `store` the consumer also bulk-stores every stage to its own HBM range
__global__ void walk_kernel(const uint8_t* src, size_t set_bytes, int replicas, int shift_stages, uint32_t stage_bytes, int depth,
                            int n_stage, int store, uint8_t* dst, size_t dst_per_cta, unsigned long long* sink,
                            unsigned long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)depth * stage_bytes);
    uint64_t* empty = full + depth;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint8_t* base = src + (size_t)(blockIdx.x % replicas) * set_bytes;
    const int set_stages = (int)(set_bytes / stage_bytes);
    const int start = (int)(((size_t)blockIdx.x * shift_stages) % set_stages);
    unsigned long long acc = 0;
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        uint32_t s = 0, ph = 0;
        int at = start;
        for (int i = 0; i < n_stage; ++i) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect(&full[s], stage_bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)s * stage_bytes)), "l"(base + (size_t)at * stage_bytes), "r"(stage_bytes), "r"(smem_u32(&full[s])) : "memory");
            if (++at == set_stages) at = 0;
            if (++s == (uint32_t)depth) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        uint32_t s = 0, ph = 0;
        uint8_t* out = dst + (size_t)blockIdx.x * dst_per_cta;
        size_t o = 0;
        for (int i = 0; i < n_stage; ++i) {
            mbar_wait(&full[s], ph);
            acc += *reinterpret_cast<volatile uint32_t*>(smem + (size_t)s * stage_bytes);
            if (store) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + o), "r"(smem_u32(smem + (size_t)s * stage_bytes)), "r"(stage_bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                o += stage_bytes;
                if (o + stage_bytes > dst_per_cta) o = 0;
            }
            mbar_arrive(&empty[s]);
            if (++s == (uint32_t)depth) { s = 0; ph ^= 1; }
        }
        if (store) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (acc == 0x1234567) *sink = acc;
        cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const size_t set_bytes = 704 * 1024;      // both nets' bf16 weights, about
    const int max_rep = 37;
    uint8_t *src, *dst; unsigned long long *sink, *cyc;
    CK(cudaMalloc(&src, set_bytes * max_rep)); CK(cudaMemset(src, 1, set_bytes * max_rep));
    const size_t dst_per_cta = 8u << 20;      // 148 x 8 MB of store targets >> L2
    CK(cudaMalloc(&dst, dst_per_cta * sms)); CK(cudaMalloc(&sink, 8)); CK(cudaMalloc(&cyc, 8 * sms));
    CK(cudaFuncSetAttribute(walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    printf("{\"sms\": %d, \"set_KB\": %zu}\n", sms, set_bytes / 1024);
    const int n_stage = 45 * 4 * 8;           // 8 launches' worth of 16 KB stages per CTA
    // {stage_bytes, depth, replicas, shift_stages, store}
    const int cfgs[][5] = {
        {16384, 3, 1, 0, 0}, {16384, 5, 1, 0, 0}, {16384, 8, 1, 0, 0}, {16384, 5, 1, 7, 0}, {16384, 5, 2, 0, 0}, {16384, 5, 8, 0, 0},
        {16384, 5, 37, 0, 0}, {16384, 5, 37, 7, 0}, {32768, 3, 1, 0, 0}, {8192, 10, 1, 0, 0},
        {16384, 3, 1, 0, 1}, {16384, 5, 1, 0, 1}, {16384, 5, 1, 7, 1}, {16384, 5, 8, 0, 1}, {16384, 5, 37, 7, 1}, {16384, 8, 37, 7, 1},
    };
    unsigned long long* h = (unsigned long long*)malloc(8 * sms);
    for (auto& c : cfgs) {
        const uint32_t sb = c[0]; const int depth = c[1];
        const size_t smem = (size_t)sb * depth + 16 * depth + 64;
        const int ns = n_stage * 16384 / (int)sb;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0);
            walk_kernel<<<sms, 64, smem>>>(src, set_bytes, c[2], c[3], sb, depth, ns, c[4], dst, dst_per_cta, sink, cyc);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0 && ms < best) best = ms;
        }
        CK(cudaMemcpy(h, cyc, 8 * sms, cudaMemcpyDeviceToHost));
        double mean = 0; unsigned long long mx = 0;
        for (int i = 0; i < sms; ++i) { mean += (double)h[i]; if (h[i] > mx) mx = h[i]; }
        mean /= sms;
        const double bytes = (double)ns * sb;
        printf("{\"stage_bytes\": %u, \"depth\": %d, \"replicas\": %d, \"shift_stages\": %d, \"store\": %d, \"ms\": %.4f, \"L2_read_GBps\": %.0f, "
               "\"B_per_clk_per_sm_mean\": %.1f, \"B_per_clk_per_sm_slowest\": %.1f, \"cycles_per_16KB\": %.0f}\n",
               sb, depth, c[2], c[3], c[4], best, bytes * sms / best / 1e6, bytes / mean, bytes / (double)mx, mean / (bytes / 16384.0));
    }
    return 0;
}
