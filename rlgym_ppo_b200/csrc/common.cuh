// Shared helpers for librlppo_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rlppo.h"

namespace rlppo {

void set_error(const char* fmt, ...);
int check_device();   // RLPPO_OK or RLPPO_ERR_DEVICE (cached per device)
int num_sms();

#define RLPPO_CHECK_ARG(cond, ...)                 \
    do {                                           \
        if (!(cond)) {                             \
            rlppo::set_error(__VA_ARGS__);         \
            return RLPPO_ERR_ARG;                  \
        }                                          \
    } while (0)

#define RLPPO_CUDA(expr)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            rlppo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                       \
            return RLPPO_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)

#define RLPPO_REQUIRE_DEVICE()                    \
    do {                                          \
        int _d = rlppo::check_device();           \
        if (_d != RLPPO_OK) return _d;            \
    } while (0)

#define RLPPO_LAUNCH_CHECK() RLPPO_CUDA(cudaGetLastError())

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// Kernels of the update chain are launched with cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel's CTAs
// may be scheduled while this one drains, run their prologue (barrier init, TMEM allocation, tensor-map prefetch) and
// then block in pdl_wait() until the previous grid has completed and its writes are visible.  A kernel launched without
// the attribute treats both instructions as no-ops.  pdl_trigger() early in a kernel whose CTAs are all resident lets the
// dependent grid be scheduled as soon as SMs free up instead of when the last CTA exits.
// NOTE: data written by the PREVIOUS kernel must not be read with __ldg (ld.global.nc) after pdl_wait(): the compiler treats
// that as an invariant load and may hoist it above the wait (seen in the optimiser kernel); use __ldcg or a plain load.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
int pdl_mode();   // 0 off, 1 on (RLPPO_PDL), cached

// cluster_x > 1: thread-block clusters of that many CTAs along x (the grid must be a multiple of it)
template <class... KArgs, class... Args>
cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x,
                               Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (pdl_mode()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    return launch_pdl_cluster(kernel, grid, block, smem, s, 1, static_cast<Args&&>(args)...);
}

__device__ __forceinline__ int ld_acquire_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_s32(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ uint16_t f32_to_bf16_bits(float f) {
    return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
__device__ __forceinline__ float bf16_bits_to_f32(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    return (uint32_t)f32_to_bf16_bits(lo) | ((uint32_t)f32_to_bf16_bits(hi) << 16);
}

}  // namespace rlppo
