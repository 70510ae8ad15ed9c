// Whole-network fused MLP kernel for the PPO update and for inference (hidden widths <= 256).
//
// The layer-by-layer path (mlp_tcgen05.cu) writes every hidden activation to HBM and reads it back in the next
// launch; for the 256-wide nets of the example config every one of those GEMMs is HBM/latency bound.  This kernel
// runs a 128-row tile of samples through the WHOLE network without leaving the SM:
//
//   x tile (TMA) -> [GEMM l -> TMEM -> bias+ReLU epilogue -> bf16 tile in shared memory = A operand of GEMM l+1] x L
//     policy: -> head GEMM -> softmax/clamp/log-prob/entropy/ratio/clip/KL epilogue -> d(loss)/d(logits) tile
//     value : -> value head as a register dot product inside the last epilogue -> d(loss)/dH_L tile
//   backward data path in the same kernel: dH_l = (dH_{l+1} W_{l+1}) (.) relu'(H_l), the ReLU masks kept as bit masks
//   in the registers of the thread that owns the row, bias gradients reduced with warp shuffles and accumulated in
//   registers across tiles.
//
// Per tile the only HBM traffic is: read x once; write H_1..H_L, d(logits) and dH_L..dH_1 once (the weight-gradient
// GEMMs, which contract over ALL rows, read them back in rlppo_linear_wgrad).  Weights stream from L2 through a
// 2-stage TMA ring.  In inference mode nothing but x, actions/log-probs or values touches HBM.
//
// Warp roles (320 threads, 1 CTA/SM, persistent over tiles):
//   warp 0      TMA producer: x tile + weight k-blocks
//   warp 1      TMEM owner; one thread issues tcgen05.mma, TMA-stores finished activation tiles, commits barriers
//   warps 2..9  epilogue: warp w owns TMEM lanes 32*(w%4)..+31 (rows), warps 2-5 the lower half of the columns,
//               warps 6-9 the upper half.  A k-block of the next A operand is released to the MMA thread as soon as
//               its columns are written, so GEMM l+1 starts while epilogue l is still running.
// Two TMEM accumulators (2 x 256 columns) alternate between consecutive GEMMs; two 64 KB activation buffers
// alternate between A-operand and epilogue-destination roles.
#include <math.h>

#include "tc_common.cuh"

namespace {

using namespace rlppo;
using namespace rlppo::tc;

constexpr int MAXL = 4;
constexpr int TILE_M = 128;
constexpr int KBLK = 64;
constexpr uint32_t KB_BYTES = TILE_M * KBLK * 2;   // 16 KB: one 64-column k-block of an activation tile
constexpr uint32_t ACT_BYTES = 4 * KB_BYTES;       // 64 KB
constexpr uint32_t WST_BYTES = 256 * KBLK * 2;     // 32 KB: one k-block of a weight operand (<= 256 rows)
constexpr int NWST = 2;
constexpr int kThreads = 320;
constexpr int MAXPH = 2 * MAXL + 1;

constexpr uint32_t OFF_ACT_A = 0;
constexpr uint32_t OFF_ACT_B = ACT_BYTES;
constexpr uint32_t OFF_WRING = 2 * ACT_BYTES;
constexpr uint32_t OFF_BIAS = OFF_WRING + NWST * WST_BYTES;          // (MAXL + 1) * 256 floats
constexpr uint32_t OFF_ROWX = OFF_BIAS + (MAXL + 1) * 256 * 4;      // 2 * 128 floats
constexpr uint32_t OFF_BARS = OFF_ROWX + 2 * 128 * 4;
constexpr uint32_t SMEM_TOTAL = OFF_BARS + 256 + 1024;              // + slack for 1024-byte alignment

enum { PH_FWD = 0, PH_FWD_VALUE = 1, PH_HEAD = 2, PH_DGRAD = 3 };

struct StoreDesc {
    uint8_t map, buf, n_kb, pad;
};
struct PhaseDesc {
    uint8_t kind;      // PH_*
    uint8_t layer;     // FWD: hidden index (0-based) of the layer produced; DGRAD: hidden index whose ReLU mask applies
    uint8_t src;       // A-operand buffer: 0 = A, 1 = B (the epilogue writes the other one)
    uint8_t n_kb;      // k-blocks of the contraction
    uint16_t N;        // accumulator columns = rows of the weight box (multiple of 16, <= 256)
    uint8_t wmap;      // index into Maps::w
    uint8_t n_store;   // tiles of the PREVIOUS epilogue to store while this phase's MMAs are issued
    StoreDesc st[2];
};

struct alignas(64) Maps {
    CUtensorMap x;
    CUtensorMap w[MAXPH];
    CUtensorMap out[MAXPH];
};

struct Params {
    int64_t M;
    int num_tiles, n_ph, L, in_kb;
    int H[MAXL];
    PhaseDesc ph[MAXPH];
    int n_tail_store;
    StoreDesc tail[2];
    const float* bias[MAXL + 1];
    float* gbias[MAXL + 1];
    // policy head
    int n_actions, out_kb;
    const float *actions, *old_logp, *adv;
    float inv_batch, clip, ent_coef;
    const float* u_inject;
    uint64_t seed, offset;
    int deterministic;
    float* actions_out;
    int64_t* actions_i64_out;
    float* logp_out;
    // value head
    const float* w_head;
    const float* targets;
    float* gw_head;
    float* values_out;
    float* metrics;
};

// ---- small PTX helpers local to this kernel -------------------------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// lane j returns sum over the 32 lanes of v[j] (v is destroyed): a 31-shuffle reduce-scatter
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const bool up = (lane & w) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const float send = up ? v[i] : v[i + w];
            const float keep = up ? v[i + w] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
    }
    return v[0];
}

__device__ __forceinline__ float bf16_round(float x) { return bf16_bits_to_f32(f32_to_bf16_bits(x)); }
// two floats -> packed bf16x2 in ONE instruction (cvt.rn.bf16x2.f32; `lo` lands in the low half)
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// write 32 consecutive columns [col0, col0+32) of one row into a K-major SWIZZLE_128B activation tile
__device__ __forceinline__ void store_chunk_sw128(uint8_t* buf, int row, int chunk32, const float (&v)[32]) {
    uint8_t* kb = buf + (chunk32 >> 1) * KB_BYTES + row * 128;
    const int base16 = (chunk32 & 1) * 4;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        uint4 o;
        o.x = cvt_bf16x2(v[t * 8 + 0], v[t * 8 + 1]);
        o.y = cvt_bf16x2(v[t * 8 + 2], v[t * 8 + 3]);
        o.z = cvt_bf16x2(v[t * 8 + 4], v[t * 8 + 5]);
        o.w = cvt_bf16x2(v[t * 8 + 6], v[t * 8 + 7]);
        *reinterpret_cast<uint4*>(kb + (((base16 + t) ^ (row & 7)) << 4)) = o;
    }
}

struct EpiCtx {
    uint8_t* smem;
    float* s_bias;
    float* s_rowx;
    uint64_t* a_ready;
    uint32_t tmem_base;
    int lane, quarter, half, row_in_tile;
};

// arrive on the a_ready barriers of the k-blocks this warp does not write (keeps every barrier at one phase per GEMM)
__device__ __forceinline__ void release_untouched(const EpiCtx& e, int c0, int c1) {
    if (e.lane == 0) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
            const int lo = max(c0, 2 * kb), hi = min(c1, 2 * kb + 2);
            if (lo >= hi) mbar_arrive(&e.a_ready[kb]);
        }
    }
}
// called after chunk c was written: if it is this warp's last chunk inside its k-block, publish the k-block
__device__ __forceinline__ void release_after_chunk(const EpiCtx& e, int c, int c1) {
    const int kb = c >> 1;
    if (c == min(c1, 2 * kb + 2) - 1) {
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (e.lane == 0) mbar_arrive(&e.a_ready[kb]);
    }
}

template <bool POLICY, bool TRAIN>
__global__ void __launch_bounds__(kThreads, 1) fused_mlp_kernel(const __grid_constant__ Maps maps, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* act[2] = {smem + OFF_ACT_A, smem + OFF_ACT_B};
    float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
    float* s_rowx = reinterpret_cast<float*>(smem + OFF_ROWX);
    uint64_t* wfull = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
    uint64_t* wempty = wfull + NWST;
    uint64_t* x_full = wempty + NWST;
    uint64_t* x_free = x_full + 1;
    uint64_t* a_ready = x_free + 1;   // [4]
    uint64_t* acc_full = a_ready + 4; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&maps.x);
        for (int i = 0; i < NWST; ++i) {
            mbar_init(&wfull[i], 1);
            mbar_init(&wempty[i], 1);
        }
        mbar_init(x_full, 1);
        mbar_init(x_free, 1);
        for (int i = 0; i < 4; ++i) mbar_init(&a_ready[i], 8);
        mbar_init(&acc_full[0], 1);
        mbar_init(&acc_full[1], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    // biases (and the value head's weights) into shared memory: slot l < L hidden biases, slot MAXL the head
    for (int i = threadIdx.x; i < (MAXL + 1) * 256; i += kThreads) {
        const int l = i >> 8, c = i & 255;
        float b = 0.f;
        if (l < p.L) {
            if (c < p.H[l]) b = __ldg(p.bias[l] + c);
        } else if (l == MAXL) {
            if (POLICY) {
                if (c < p.n_actions && p.bias[MAXL] != nullptr) b = __ldg(p.bias[MAXL] + c);
            } else {
                if (c < p.H[p.L - 1]) b = __ldg(p.w_head + c);
            }
        }
        s_bias[i] = b;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t ws = 0, wpar = 0, xpar = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(x_free, xpar ^ 1);
                mbar_expect_tx(x_full, p.in_kb * KB_BYTES);
                for (int kb = 0; kb < p.in_kb; ++kb)
                    tma_load_2d(&maps.x, x_full, act[1] + kb * KB_BYTES, kb * KBLK, tile * TILE_M);
                xpar ^= 1;
                for (int ph = 0; ph < p.n_ph; ++ph) {
                    const PhaseDesc& d = p.ph[ph];
                    for (int kb = 0; kb < d.n_kb; ++kb) {
                        mbar_wait(&wempty[ws], wpar ^ 1);
                        mbar_expect_tx(&wfull[ws], (uint32_t)d.N * 128u);
                        tma_load_2d(&maps.w[d.wmap], &wfull[ws], smem + OFF_WRING + ws * WST_BYTES, kb * KBLK, 0);
                        if (++ws == NWST) {
                            ws = 0;
                            wpar ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer + TMA stores =====================
        if (lane == 0) {
            uint32_t ws = 0, wpar = 0, xpar = 0, g = 0, acount = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int ph = 0; ph < p.n_ph; ++ph, ++g) {
                    const PhaseDesc& d = p.ph[ph];
                    const uint32_t acc = g & 1;
                    const uint32_t d_tmem = tmem_base + acc * 256;
                    const uint32_t idesc = umma_idesc_bf16(TILE_M, d.N, 0, 0);
                    const uint32_t apar = acount & 1;
                    if (ph == 0) {
                        mbar_wait(x_full, xpar);
                        xpar ^= 1;
                        tc_fence_after();
                    }
                    uint32_t waited = 0;
                    for (int kb = 0; kb < d.n_kb; ++kb) {
                        if (ph > 0) {
                            mbar_wait(&a_ready[kb], apar);
                            waited |= 1u << kb;
                            tc_fence_after();
                            if (TRAIN) {
                                for (int s = 0; s < d.n_store; ++s)
                                    if (d.st[s].buf == d.src && kb < d.st[s].n_kb)
                                        tma_store_2d(&maps.out[d.st[s].map], act[d.src] + kb * KB_BYTES, kb * KBLK,
                                                     tile * TILE_M);
                            }
                        }
                        mbar_wait(&wfull[ws], wpar);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(act[d.src] + kb * KB_BYTES);
                        const uint32_t b_addr = smem_u32(smem + OFF_WRING + ws * WST_BYTES);
#pragma unroll
                        for (int k = 0; k < KBLK / 16; ++k) {
                            const uint64_t ad = umma_smem_desc(a_addr + k * 32, 16, 1024);
                            const uint64_t bd = umma_smem_desc(b_addr + k * 32, 16, 1024);
                            umma_bf16(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit(&wempty[ws]);
                        if (++ws == NWST) {
                            ws = 0;
                            wpar ^= 1;
                        }
                    }
                    if (ph > 0) {
                        if (TRAIN && d.n_store > 0) {
                            // tiles not covered above: other buffer, or more k-blocks than this GEMM contracts over
                            for (int s = 0; s < d.n_store; ++s) {
                                const StoreDesc& sd = d.st[s];
                                for (int kb = 0; kb < sd.n_kb; ++kb) {
                                    if (sd.buf == d.src && kb < d.n_kb) continue;
                                    if (!(waited & (1u << kb))) {
                                        mbar_wait(&a_ready[kb], apar);
                                        waited |= 1u << kb;
                                    }
                                    tma_store_2d(&maps.out[sd.map], act[sd.buf] + kb * KB_BYTES, kb * KBLK,
                                                 tile * TILE_M);
                                }
                            }
                            bulk_commit();
                        }
                        ++acount;
                    }
                    if (TRAIN) bulk_wait_read_all();   // the epilogue of this GEMM may overwrite either buffer
                    umma_commit(&acc_full[acc]);
                }
                // tail: the last epilogue's tiles, then hand buffer B back to the producer for the next x tile
                {
                    const uint32_t apar = acount & 1;
                    for (int kb = 0; kb < 4; ++kb) mbar_wait(&a_ready[kb], apar);
                    ++acount;
                    if (TRAIN) {
                        for (int s = 0; s < p.n_tail_store; ++s)
                            for (int kb = 0; kb < p.tail[s].n_kb; ++kb)
                                tma_store_2d(&maps.out[p.tail[s].map], act[p.tail[s].buf] + kb * KB_BYTES, kb * KBLK,
                                             tile * TILE_M);
                        bulk_commit();
                        bulk_wait_read_all();
                    }
                    mbar_arrive(x_free);
                }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        EpiCtx e;
        e.smem = smem;
        e.s_bias = s_bias;
        e.s_rowx = s_rowx;
        e.a_ready = a_ready;
        e.tmem_base = tmem_base;
        e.lane = lane;
        e.quarter = warp & 3;
        e.half = (warp - 2) >> 2;
        e.row_in_tile = e.quarter * 32 + lane;

        uint32_t relu[MAXL][4];
        float dbacc[MAXL + 1][4];
        float dwacc[4];
#pragma unroll
        for (int l = 0; l < MAXL; ++l)
#pragma unroll
            for (int j = 0; j < 4; ++j) relu[l][j] = 0u;
#pragma unroll
        for (int l = 0; l <= MAXL; ++l)
#pragma unroll
            for (int j = 0; j < 4; ++j) dbacc[l][j] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) dwacc[j] = 0.f;
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, mrows = 0.f;   // metric partial sums of this thread's rows

        // The chunk loops below are deliberately NOT unrolled and take the layer as a run-time index (register arrays
        // are read/written through predicated selects): the first version specialised every (layer, chunk) pair and
        // grew to 30k instructions (480 KB), and ncu showed the epilogue warps stalled on instruction fetch
        // (stall_no_inst = half of all samples).
        auto relu_get = [&](int l, int j) {
            uint32_t v = 0u;
#pragma unroll
            for (int a = 0; a < MAXL; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) v = (a == l && b == j) ? relu[a][b] : v;
            return v;
        };
        auto relu_set = [&](int l, int j, uint32_t v) {
#pragma unroll
            for (int a = 0; a < MAXL; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) relu[a][b] = (a == l && b == j) ? v : relu[a][b];
        };
        auto db_add = [&](int l, int j, float v) {
#pragma unroll
            for (int a = 0; a <= MAXL; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dbacc[a][b] += (a == l && b == j) ? v : 0.f;
        };
        auto dw_add = [&](int j, float v) {
#pragma unroll
            for (int b = 0; b < 4; ++b) dwacc[b] += (b == j) ? v : 0.f;
        };

        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int64_t row = (int64_t)tile * TILE_M + e.row_in_tile;
            const bool row_ok = row < p.M;
            for (int ph = 0; ph < p.n_ph; ++ph, ++g) {
                const PhaseDesc& d = p.ph[ph];
                const uint32_t acc = g & 1;
                mbar_wait(&acc_full[acc], (g >> 1) & 1);
                tc_fence_after();
                const uint32_t trow = tmem_base + ((uint32_t)(e.quarter * 32) << 16) + acc * 256;
                uint8_t* dst = act[d.src ^ 1];
                uint8_t* srcb = act[d.src];
                const int li = d.layer;

                if (d.kind == PH_FWD || d.kind == PH_FWD_VALUE) {
                    const int nc = d.N >> 5;
                    const int c0 = e.half * (nc >> 1), c1 = c0 + (nc >> 1);
                    const bool tail = (d.kind == PH_FWD_VALUE);
                    if (!tail) release_untouched(e, c0, c1);
                    float dot = 0.f;
#pragma unroll 1
                    for (int c = c0; c < c1; ++c) {
                        float v[32];
                        tmem_ld32(trow + c * 32, v);
                        const float* sb = s_bias + li * 256 + c * 32;
                        uint32_t bits = 0u;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float xv = fmaxf(v[i] + sb[i], 0.f);
                            bits |= (xv > 0.f ? 1u : 0u) << i;
                            v[i] = xv;
                        }
                        if (TRAIN) relu_set(li, c - c0, bits);
                        if (tail) {
                            const float* wv = s_bias + MAXL * 256 + c * 32;
#pragma unroll
                            for (int i = 0; i < 32; ++i) dot = fmaf(bf16_round(v[i]), wv[i], dot);
                        }
                        if (!tail) {
                            store_chunk_sw128(dst, e.row_in_tile, c, v);
                            release_after_chunk(e, c, c1);
                        } else if (TRAIN) {
                            store_chunk_sw128(dst, e.row_in_tile, c, v);
                        }
                    }
                    if (tail) {
                        // ---- value head: v = H_L . w + b (value_estimator.py:27), MSE loss and its gradient ----
                        e.s_rowx[e.half * 128 + e.row_in_tile] = dot;
                        epi_bar_sync();
                        const float bhead = p.bias[MAXL] != nullptr ? __ldg(p.bias[MAXL]) : 0.f;
                        const float val = e.s_rowx[e.row_in_tile] + e.s_rowx[128 + e.row_in_tile] + bhead;
                        if (e.half == 0 && row_ok && p.values_out != nullptr) p.values_out[row] = val;
                        if (!TRAIN) {
                            tc_fence_before();
                            __syncwarp();
                            release_untouched(e, 0, 0);   // nothing written: hand every k-block barrier back
                        }
                        if (TRAIN) {
                            float dv = 0.f;
                            if (row_ok) {
                                const float err = val - __ldg(p.targets + row);
                                dv = 2.0f * p.inv_batch * err;                     // d(MSE * mb/B)/dv, ppo_learner.py:176
                                if (e.half == 0) {
                                    m0 += err * err;
                                    mrows += 1.f;
                                    m1 += dv;                                      // head bias gradient
                                }
                            }
                            release_untouched(e, c0, c1);
#pragma unroll 1
                            for (int c = c0; c < c1; ++c) {
                                float v[32], t[32];
                                tmem_ld32(trow + c * 32, v);
                                const float* sb = s_bias + li * 256 + c * 32;
                                const float* wv = s_bias + MAXL * 256 + c * 32;
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    const float h = bf16_round(fmaxf(v[i] + sb[i], 0.f));
                                    t[i] = dv * h;
                                    v[i] = h > 0.f ? dv * wv[i] : 0.f;
                                }
                                dw_add(c - c0, warp_colsum32(t, e.lane));
#pragma unroll
                                for (int i = 0; i < 32; ++i) t[i] = v[i];
                                db_add(li, c - c0, warp_colsum32(t, e.lane));
                                store_chunk_sw128(srcb, e.row_in_tile, c, v);
                                release_after_chunk(e, c, c1);
                            }
                        }
                    }
                } else if (d.kind == PH_DGRAD) {
                    const int nc = d.N >> 5;
                    const int c0 = e.half * (nc >> 1), c1 = c0 + (nc >> 1);
                    release_untouched(e, c0, c1);
#pragma unroll 1
                    for (int c = c0; c < c1; ++c) {
                        float v[32], t[32];
                        tmem_ld32(trow + c * 32, v);
                        const uint32_t bits = relu_get(li, c - c0);
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            v[i] = ((bits >> i) & 1u) ? v[i] : 0.f;
                            t[i] = v[i];
                        }
                        db_add(li, c - c0, warp_colsum32(t, e.lane));
                        store_chunk_sw128(dst, e.row_in_tile, c, v);
                        release_after_chunk(e, c, c1);
                    }
                } else {
                    // ---- policy head (discrete_policy.py:44-80, ppo_learner.py:153-177) ----
                    // one thread = one row; the logits stay in TMEM and are re-read chunk by chunk in each pass (a TMEM
                    // load is cheaper than the code size of keeping 128 logits in registers).  The upper-half warps have
                    // nothing to do in this phase.
                    const int nact = p.n_actions;
                    const int nch = (nact + 31) >> 5;          // <= 4
                    const int nch_out = p.out_kb * 2;          // chunks of the d(logits) tile (whole k-blocks)
                    if (e.half == 1) {
                        release_untouched(e, 0, 0);
                    } else {
                        release_untouched(e, 0, TRAIN ? nch_out : 0);
                        const float* sb = s_bias + MAXL * 256;
                        float mx = -INFINITY;
                        int argmax = 0;
#pragma unroll 1
                        for (int c = 0; c < nch; ++c) {
                            float v[32];
                            tmem_ld32(trow + c * 32, v);
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const int col = c * 32 + i;
                                const float zz = col < nact ? v[i] + sb[col] : -INFINITY;
                                if (zz > mx) {
                                    mx = zz;
                                    argmax = col;
                                }
                            }
                        }
                        float S = 0.f;
#pragma unroll 1
                        for (int c = 0; c < nch; ++c) {
                            float v[32];
                            tmem_ld32(trow + c * 32, v);
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const int col = c * 32 + i;
                                S += col < nact ? __expf(v[i] + sb[col] - mx) : 0.f;
                            }
                        }
                        const float logS = logf(S);
                        const float kLogMin = -25.328436022934504f;   // ln(1e-11)
                        if (TRAIN) {
                            int a = 0;
                            float old_lp = 0.f, advv = 0.f;
                            if (row_ok) {
                                a = (int)__ldg(p.actions + row);             // acts.long(), discrete_policy.py:71
                                a = min(max(a, 0), nact - 1);
                                old_lp = __ldg(p.old_logp + row);
                                advv = __ldg(p.adv + row);
                            }
                            // pass: entropy, sum_j m_j s_j (log p_j + 1), the action's log-softmax
                            float Hent = 0.f, Gs = 0.f, ls_a = 0.f;
#pragma unroll 1
                            for (int c = 0; c < nch; ++c) {
                                float v[32];
                                tmem_ld32(trow + c * 32, v);
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    const int col = c * 32 + i;
                                    if (col < nact) {
                                        const float ls = (v[i] + sb[col] - mx) - logS;      // log softmax (<= 0)
                                        const float s = __expf(ls);
                                        const float lp = fminf(fmaxf(ls, kLogMin), 0.f);    // log clamp(s, 1e-11, 1), :74-76
                                        const float pj = fminf(fmaxf(s, 1e-11f), 1.0f);
                                        Hent -= pj * lp;                                    // :78
                                        Gs += ls >= kLogMin ? s * (lp + 1.0f) : 0.f;        // clamp passes gradient inside only
                                        ls_a = col == a ? ls : ls_a;
                                    }
                                }
                            }
                            const float lp_a = fminf(fmaxf(ls_a, kLogMin), 0.f);
                            const float s_a = __expf(ls_a);
                            const float p_a = fminf(fmaxf(s_a, 1e-11f), 1.0f);
                            const float log_ratio = lp_a - old_lp;
                            const float ratio = expf(log_ratio);                            // ppo_learner.py:153
                            const float lo = 1.0f - p.clip, hi = 1.0f + p.clip;
                            const float clipped = fminf(fmaxf(ratio, lo), hi);              // :154-156
                            const float s1 = ratio * advv, s2 = clipped * advv;
                            const float in_range = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
                            const float d1 = s1 < s2 ? 1.f : (s1 == s2 ? 0.5f : 0.f);      // torch.min backward, ties 0.5/0.5
                            const float d2 = s1 > s2 ? 1.f : (s1 == s2 ? 0.5f : 0.f);
                            const float okf = row_ok ? 1.f : 0.f;
                            const float d_logp = -p.inv_batch * advv * (d1 + d2 * in_range) * ratio * okf;
                            const float ga = (ls_a >= kLogMin) ? d_logp / p_a : 0.f;        // through log(clamp(s_a))
                            const float cw = p.ent_coef * p.inv_batch * okf;
                            const float G = cw * Gs + ga * s_a;
                            if (row_ok) {
                                m0 += Hent;
                                m1 += (ratio - 1.0f) - log_ratio;                           // :161
                                m2 += fabsf(ratio - 1.0f) > p.clip ? 1.f : 0.f;             // :166
                                m3 += fminf(s1, s2);
                                mrows += 1.f;
                                if (p.logp_out) p.logp_out[row] = lp_a;
                            }
                            // pass: dz_j = s_j (g_j - G) -> bf16 tile (A operand of the first dgrad GEMM) + head bias grads
#pragma unroll 1
                            for (int c = 0; c < nch_out; ++c) {
                                float v[32], t[32];
                                if (c < nch) {
                                    tmem_ld32(trow + c * 32, v);
                                } else {
#pragma unroll
                                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                                }
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    const int col = c * 32 + i;
                                    float o = 0.f;
                                    if (col < nact) {
                                        const float ls = (v[i] + sb[col] - mx) - logS;
                                        const float s = __expf(ls);
                                        const float lp = fminf(fmaxf(ls, kLogMin), 0.f);
                                        float gj = cw * (lp + 1.0f) + (col == a ? ga : 0.f);
                                        gj = ls >= kLogMin ? gj : 0.f;
                                        o = s * (gj - G);
                                    }
                                    v[i] = o;
                                    t[i] = o;
                                }
                                db_add(MAXL, c, warp_colsum32(t, e.lane));
                                store_chunk_sw128(dst, e.row_in_tile, c, v);
                                release_after_chunk(e, c, nch_out);
                            }
                        } else {
                            // ---- sampling (DiscreteFF.get_action, discrete_policy.py:44-62) ----
                            float P = 0.f;
#pragma unroll 1
                            for (int c = 0; c < nch; ++c) {
                                float v[32];
                                tmem_ld32(trow + c * 32, v);
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    const int col = c * 32 + i;
                                    if (col < nact) P += fminf(fmaxf(__expf((v[i] + sb[col] - mx) - logS), 1e-11f), 1.0f);
                                }
                            }
                            int actn = nact - 1;
                            float pa = 0.f;
                            if (p.deterministic) {
                                actn = argmax;
                                pa = fminf(fmaxf(__expf(-logS), 1e-11f), 1.0f);
                            } else {
                                float u = 0.f;
                                if (row_ok) {
                                    if (p.u_inject != nullptr) {
                                        u = __ldg(p.u_inject + row);
                                    } else {
                                        const uint64_t ctr = p.offset + (uint64_t)row;
                                        const uint4 r = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u),
                                                                      make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
                                        u = (float)(r.x >> 8) * (1.0f / 16777216.0f);
                                    }
                                }
                                const float thr = u * P;   // torch.multinomial normalises what it is given
                                float run = 0.f, plast = 0.f;
                                bool found = false;
#pragma unroll 1
                                for (int c = 0; c < nch; ++c) {
                                    float v[32];
                                    tmem_ld32(trow + c * 32, v);
#pragma unroll
                                    for (int i = 0; i < 32; ++i) {
                                        const int col = c * 32 + i;
                                        if (col < nact) {
                                            const float pj = fminf(fmaxf(__expf((v[i] + sb[col] - mx) - logS), 1e-11f), 1.0f);
                                            run += pj;
                                            plast = pj;
                                            if (!found && run > thr) {
                                                found = true;
                                                actn = col;
                                                pa = pj;
                                            }
                                        }
                                    }
                                }
                                if (!found) pa = plast;
                            }
                            if (row_ok) {
                                if (p.actions_out) p.actions_out[row] = (float)actn;   // batched_agent_manager.py:204
                                if (p.actions_i64_out) p.actions_i64_out[row] = (int64_t)actn;
                                if (p.logp_out) p.logp_out[row] = logf(pa);            // :60
                            }
                        }
                    }
                }
                if (!TRAIN) {
                    // inference epilogues that wrote nothing still have to order their TMEM reads before the next MMA
                    tc_fence_before();
                }
            }
        }
        // ---- flush the per-thread accumulators: bias gradients, value-head weight gradient, metrics ----
        if (TRAIN) {
#pragma unroll
            for (int l = 0; l < MAXL; ++l) {
                if (l < p.L && p.gbias[l] != nullptr) {
                    const int nc = p.H[l] >> 5;
                    const int c0 = e.half * (nc >> 1);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < (nc >> 1)) atomicAdd(p.gbias[l] + (c0 + j) * 32 + e.lane, dbacc[l][j]);
                }
            }
            if (POLICY) {
                if (e.half == 0 && p.gbias[MAXL] != nullptr) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int col = c * 32 + e.lane;
                        if (col < p.n_actions) atomicAdd(p.gbias[MAXL] + col, dbacc[MAXL][c]);
                    }
                }
            } else {
                const int nc = p.H[p.L - 1] >> 5;
                const int c0 = e.half * (nc >> 1);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < (nc >> 1)) atomicAdd(p.gw_head + (c0 + j) * 32 + e.lane, dwacc[j]);
            }
            if (p.metrics != nullptr && e.half == 0) {
                const float r0 = warp_sum(m0), r1 = warp_sum(m1), r2 = warp_sum(m2), r3 = warp_sum(m3),
                            rr = warp_sum(mrows);
                if (e.lane == 0 && rr > 0.f) {
                    if (POLICY) {
                        atomicAdd(p.metrics + 0, r0);
                        atomicAdd(p.metrics + 1, r1);
                        atomicAdd(p.metrics + 2, r2);
                        atomicAdd(p.metrics + 3, r3);
                        atomicAdd(p.metrics + 4, rr);
                    } else {
                        atomicAdd(p.metrics + 5, r0);
                        atomicAdd(p.metrics + 6, rr);
                        if (p.gbias[MAXL] != nullptr) atomicAdd(p.gbias[MAXL], r1);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
int check_net(const rlppo_fused_net* net) {
    RLPPO_CHECK_ARG(net != nullptr, "null net descriptor");
    RLPPO_CHECK_ARG(net->n_hidden >= 1 && net->n_hidden <= MAXL, "fused path: 1..%d hidden layers", MAXL);
    RLPPO_CHECK_ARG(net->in_dim >= 1 && net->in_dim <= 256 && net->in_ld % 8 == 0 && net->in_ld >= net->in_dim,
                    "fused path: input width <= 256, row stride a multiple of 8");
    for (int l = 0; l < net->n_hidden; ++l)
        RLPPO_CHECK_ARG(net->hidden[l] >= 64 && net->hidden[l] <= 256 && net->hidden[l] % 64 == 0,
                        "fused path: hidden widths must be 64, 128, 192 or 256 (got %d)", net->hidden[l]);
    return RLPPO_OK;
}

template <bool POLICY, bool TRAIN>
int launch_fused(const rlppo_fused_net* net, const uint16_t* x, int64_t M, Params& p, cudaStream_t s) {
    int rc = check_net(net);
    if (rc) return rc;
    RLPPO_CHECK_ARG(M >= 1 && M < (1ll << 31) - TILE_M, "bad row count");
    const int L = net->n_hidden;
    Maps maps;
    p.M = M;
    p.num_tiles = (int)((M + TILE_M - 1) / TILE_M);
    p.L = L;
    p.in_kb = (net->in_dim + KBLK - 1) / KBLK;
    for (int l = 0; l < MAXL; ++l) p.H[l] = l < L ? net->hidden[l] : 0;
    for (int l = 0; l < L; ++l) {
        p.bias[l] = net->bias[l];
        p.gbias[l] = TRAIN ? net->gbias[l] : nullptr;
        RLPPO_CHECK_ARG(net->bias[l] != nullptr, "missing bias");
    }
    p.bias[MAXL] = net->bias[L];
    p.gbias[MAXL] = TRAIN ? net->gbias[L] : nullptr;

    rc = make_tmap_bf16_2d(&maps.x, x, (uint64_t)M, (uint64_t)net->in_dim, (uint64_t)net->in_ld, TILE_M);
    if (rc) return rc;

    int nw = 0, nout = 0, nph = 0;
    auto add_w = [&](const uint16_t* w, int64_t ld, int rows, int cols, int box_rows) -> int {
        RLPPO_CHECK_ARG(w != nullptr && ld % 8 == 0, "missing weight operand");
        int r = make_tmap_bf16_2d(&maps.w[nw], w, (uint64_t)rows, (uint64_t)cols, (uint64_t)ld, (uint32_t)box_rows);
        if (r) return r;
        ++nw;
        return RLPPO_OK;
    };
    auto add_out = [&](uint16_t* o, int64_t ld, int cols) -> int {
        RLPPO_CHECK_ARG(o != nullptr && ld % 8 == 0 && ld >= cols, "missing activation / gradient output buffer");
        int r = make_tmap_bf16_2d(&maps.out[nout], o, (uint64_t)M, (uint64_t)cols, (uint64_t)ld, TILE_M);
        if (r) return r;
        ++nout;
        return RLPPO_OK;
    };
    const int out_pad8 = POLICY ? (p.n_actions + 7) / 8 * 8 : 0;
    const int head_N = POLICY ? (p.n_actions + 15) / 16 * 16 : 0;
    p.out_kb = POLICY ? (out_pad8 + KBLK - 1) / KBLK : 0;

    // output maps: [0..L) = H_l ; policy: [L] = dz, [L+1 .. 2L] = dH_L .. dH_1 ; value: [L .. 2L) = dH_L .. dH_1
    int map_h[MAXL], map_dz = -1, map_dh[MAXL];
    if (TRAIN) {
        for (int l = 0; l < L; ++l) {
            map_h[l] = nout;
            rc = add_out(net->h[l], net->h_ld[l], net->hidden[l]);
            if (rc) return rc;
        }
        if (POLICY) {
            map_dz = nout;
            rc = add_out(net->dz, net->dz_ld, out_pad8);
            if (rc) return rc;
        }
        for (int l = L - 1; l >= 0; --l) {
            map_dh[l] = nout;
            rc = add_out(net->dh[l], net->dh_ld[l], net->hidden[l]);
            if (rc) return rc;
        }
    }
    auto kb_of = [](int cols) { return (cols + KBLK - 1) / KBLK; };

    // ---- forward phases ----
    int src = 1;   // x arrives in buffer B
    for (int l = 0; l < L; ++l) {
        PhaseDesc& d = p.ph[nph];
        d = PhaseDesc{};
        d.kind = (!POLICY && l == L - 1) ? PH_FWD_VALUE : PH_FWD;
        d.layer = (uint8_t)l;
        d.src = (uint8_t)src;
        const int K = l == 0 ? net->in_dim : net->hidden[l - 1];
        d.n_kb = (uint8_t)kb_of(K);
        d.N = (uint16_t)net->hidden[l];
        d.wmap = (uint8_t)nw;
        rc = add_w(net->wq[l], net->wq_ld[l], net->hidden[l], K, net->hidden[l]);
        if (rc) return rc;
        if (TRAIN && l > 0) {
            d.n_store = 1;
            d.st[0] = StoreDesc{(uint8_t)map_h[l - 1], (uint8_t)src, (uint8_t)kb_of(net->hidden[l - 1]), 0};
        }
        ++nph;
        src ^= 1;
    }
    // after the loop H_L sits in buffer `src` (the last epilogue's destination)
    if (POLICY) {
        PhaseDesc& d = p.ph[nph];
        d = PhaseDesc{};
        d.kind = PH_HEAD;
        d.layer = (uint8_t)L;
        d.src = (uint8_t)src;
        d.n_kb = (uint8_t)kb_of(net->hidden[L - 1]);
        d.N = (uint16_t)head_N;
        d.wmap = (uint8_t)nw;
        rc = add_w(net->wq[L], net->wq_ld[L], out_pad8, net->hidden[L - 1], head_N);
        if (rc) return rc;
        if (TRAIN) {
            d.n_store = 1;
            d.st[0] = StoreDesc{(uint8_t)map_h[L - 1], (uint8_t)src, (uint8_t)kb_of(net->hidden[L - 1]), 0};
        }
        ++nph;
        src ^= 1;   // dz sits in `src`
    }
    p.n_tail_store = 0;
    if (TRAIN) {
        // ---- backward data phases ----
        if (POLICY) {
            PhaseDesc& d = p.ph[nph];
            d = PhaseDesc{};
            d.kind = PH_DGRAD;
            d.layer = (uint8_t)(L - 1);
            d.src = (uint8_t)src;
            d.n_kb = (uint8_t)p.out_kb;
            d.N = (uint16_t)net->hidden[L - 1];
            d.wmap = (uint8_t)nw;
            rc = add_w(net->wt[L], net->wt_ld[L], net->hidden[L - 1], out_pad8, net->hidden[L - 1]);
            if (rc) return rc;
            d.n_store = 1;
            d.st[0] = StoreDesc{(uint8_t)map_dz, (uint8_t)src, (uint8_t)p.out_kb, 0};
            ++nph;
            src ^= 1;   // dH_L sits in `src`
        } else {
            // value tail: H_L went to buffer `src`, dH_L to the other one, which is the next A operand
            src ^= 1;
        }
        for (int l = L - 1; l >= 1; --l) {
            // produce dH_l (hidden index l-1) from dH_{l+1} (hidden index l) with W_{l+1}^T = wt[l]
            PhaseDesc& d = p.ph[nph];
            d = PhaseDesc{};
            d.kind = PH_DGRAD;
            d.layer = (uint8_t)(l - 1);
            d.src = (uint8_t)src;
            d.n_kb = (uint8_t)kb_of(net->hidden[l]);
            d.N = (uint16_t)net->hidden[l - 1];
            d.wmap = (uint8_t)nw;
            rc = add_w(net->wt[l], net->wt_ld[l], net->hidden[l - 1], net->hidden[l], net->hidden[l - 1]);
            if (rc) return rc;
            d.n_store = 1;
            d.st[0] = StoreDesc{(uint8_t)map_dh[l], (uint8_t)src, (uint8_t)kb_of(net->hidden[l]), 0};
            if (!POLICY && l == L - 1) {   // first phase after the value tail also stores H_L from the other buffer
                d.n_store = 2;
                d.st[1] = StoreDesc{(uint8_t)map_h[L - 1], (uint8_t)(src ^ 1), (uint8_t)kb_of(net->hidden[L - 1]), 0};
            }
            ++nph;
            src ^= 1;
        }
        // tail: the gradient tile produced by the last epilogue (dH_1, or dH_L when L == 1)
        p.tail[p.n_tail_store++] = StoreDesc{(uint8_t)map_dh[0], (uint8_t)src, (uint8_t)kb_of(net->hidden[0]), 0};
        if (!POLICY && L == 1)
            p.tail[p.n_tail_store++] = StoreDesc{(uint8_t)map_h[0], (uint8_t)(src ^ 1), (uint8_t)kb_of(net->hidden[0]), 0};
    }
    p.n_ph = nph;

    static bool configured = false;
    auto kfn = fused_mlp_kernel<POLICY, TRAIN>;
    if (!configured) {
        RLPPO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL));
        configured = true;
    }
    const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    kfn<<<grid, kThreads, SMEM_TOTAL, s>>>(maps, p);
    RLPPO_LAUNCH_CHECK();
    return RLPPO_OK;
}

}  // namespace

extern "C" {

int rlppo_policy_train_fused(const rlppo_fused_net* net, const uint16_t* x, int64_t M, int n_actions,
                             const float* actions, const float* old_logp, const float* adv, float inv_batch, float clip,
                             float ent_coef, float* logp_out, float* metrics, void* stream) {
    RLPPO_REQUIRE_DEVICE();
    RLPPO_CHECK_ARG(x && actions && old_logp && adv, "null pointer");
    RLPPO_CHECK_ARG(n_actions >= 1 && n_actions <= 128, "fused path: n_actions must be in [1,128]");
    Params p{};
    p.n_actions = n_actions;
    p.actions = actions; p.old_logp = old_logp; p.adv = adv;
    p.inv_batch = inv_batch; p.clip = clip; p.ent_coef = ent_coef;
    p.logp_out = logp_out; p.metrics = metrics;
    return launch_fused<true, true>(net, x, M, p, static_cast<cudaStream_t>(stream));
}

int rlppo_policy_infer_fused(const rlppo_fused_net* net, const uint16_t* x, int64_t M, int n_actions,
                             const float* u_inject, uint64_t seed, uint64_t offset, int deterministic,
                             float* actions_out, int64_t* actions_i64_out, float* logp_out, void* stream) {
    RLPPO_REQUIRE_DEVICE();
    RLPPO_CHECK_ARG(x != nullptr, "null pointer");
    RLPPO_CHECK_ARG(n_actions >= 1 && n_actions <= 128, "fused path: n_actions must be in [1,128]");
    Params p{};
    p.n_actions = n_actions;
    p.u_inject = u_inject; p.seed = seed; p.offset = offset; p.deterministic = deterministic;
    p.actions_out = actions_out; p.actions_i64_out = actions_i64_out; p.logp_out = logp_out;
    return launch_fused<true, false>(net, x, M, p, static_cast<cudaStream_t>(stream));
}

int rlppo_value_train_fused(const rlppo_fused_net* net, const uint16_t* x, int64_t M, const float* w_head,
                            const float* targets, float inv_batch, float* gw_head, float* values_out, float* metrics,
                            void* stream) {
    RLPPO_REQUIRE_DEVICE();
    RLPPO_CHECK_ARG(x && w_head && targets && gw_head, "null pointer");
    Params p{};
    p.w_head = w_head; p.targets = targets; p.inv_batch = inv_batch; p.gw_head = gw_head;
    p.values_out = values_out; p.metrics = metrics;
    return launch_fused<false, true>(net, x, M, p, static_cast<cudaStream_t>(stream));
}

int rlppo_value_infer_fused(const rlppo_fused_net* net, const uint16_t* x, int64_t M, const float* w_head,
                            float* values_out, void* stream) {
    RLPPO_REQUIRE_DEVICE();
    RLPPO_CHECK_ARG(x && w_head && values_out, "null pointer");
    Params p{};
    p.w_head = w_head; p.values_out = values_out;
    return launch_fused<false, false>(net, x, M, p, static_cast<cudaStream_t>(stream));
}
}
