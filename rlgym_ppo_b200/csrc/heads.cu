// The other two action heads of the reference (selected by policy_type at ppo_learner.py:34-50):
//   MultiDiscreteFF   multi_discrete_policy.py:16-89 + torch_functions.MultiDiscreteRolv (:81-122): 21 logits = eight
//                     categoricals (bins 3,3,3,3,3,2,2,2); log-prob and entropy summed over the eight.
//   ContinuousPolicy  continuous_policy.py:23-120 + MapContinuousToAction (torch_functions.py:15-33): Tanh over 2n outputs,
//                     mean = first half, std = second half mapped onto [var_min, var_max], diagonal Gaussian.
// Their last Linear runs on the tensor-core GEMM (rlppo_linear_fwd[_split], output written as three bf16 parts so the
// logits keep fp32 precision); these kernels are the per-row tails -- sampling (get_action), and for training
// get_backprop_data + the PPO loss block (ppo_learner.py:146-185) + its analytic backward into the logits.  At most 21 /
// 2n <= 64 columns per row: one thread per row, HBM-bound, ~100-200 B per row.
#include <math.h>

#include "common.cuh"

namespace {

using namespace rlppo;

struct ZView {               // split bf16 matrix (rlppo_split layout): value = sum of parts
    const uint16_t* p;
    int64_t ld;
    int parts;
    int64_t pstride;
};
__device__ __forceinline__ float zload(const ZView& z, int64_t row, int col) {
    float s = 0.f;
    for (int q = z.parts - 1; q >= 0; --q) s += bf16_bits_to_f32(z.p[row * z.ld + q * z.pstride + col]);
    return s;
}
struct DzView {
    uint16_t* p;
    int64_t ld;
    int parts;
    int64_t pstride;
};
__device__ __forceinline__ void dzstore(const DzView& d, int64_t row, int col, float x) {
    for (int q = 0; q < d.parts; ++q) {
        const uint16_t b = f32_to_bf16_bits(x);
        d.p[row * d.ld + q * d.pstride + col] = b;
        x -= bf16_bits_to_f32(b);
    }
}

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
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

// The PPO loss block shared by every head (ppo_learner.py:153-177) and its derivative with respect to the log-prob.
struct PpoRow {
    float d_logp, kl, clipc, surr;
};
__device__ __forceinline__ PpoRow ppo_row(float logp, float old_lp, float advv, float clip, float inv_batch) {
    const float log_ratio = logp - old_lp;
    const float ratio = expf(log_ratio);
    const float lo = 1.0f - clip, hi = 1.0f + clip;
    const float clipped = fminf(fmaxf(ratio, lo), hi);
    const float s1 = ratio * advv, s2 = clipped * advv;
    const float in_range = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
    const float d1 = s1 < s2 ? 1.f : (s1 == s2 ? 0.5f : 0.f);      // torch.min backward, ties 0.5 / 0.5
    const float d2 = s1 > s2 ? 1.f : (s1 == s2 ? 0.5f : 0.f);
    PpoRow r;
    r.d_logp = -inv_batch * advv * (d1 + d2 * in_range) * ratio;
    r.kl = (ratio - 1.0f) - log_ratio;
    r.clipc = fabsf(ratio - 1.0f) > clip ? 1.f : 0.f;
    r.surr = fminf(s1, s2);
    return r;
}
__device__ __forceinline__ void metrics_add(float* metrics, float ent, const PpoRow& r, bool ok) {
    const float f = ok ? 1.f : 0.f;
    const float a = warp_sum(ent * f), b = warp_sum(r.kl * f), c = warp_sum(r.clipc * f), d = warp_sum(r.surr * f),
                e = warp_sum(f);
    if ((threadIdx.x & 31) == 0 && metrics != nullptr && e > 0.f) {
        atomicAdd(metrics + 0, a);
        atomicAdd(metrics + 1, b);
        atomicAdd(metrics + 2, c);
        atomicAdd(metrics + 3, d);
        atomicAdd(metrics + 4, e);
    }
}

constexpr int MD_GROUPS = 8;
__constant__ int kBins[MD_GROUPS] = {3, 3, 3, 3, 3, 2, 2, 2};     // multi_discrete_policy.py:21
__constant__ int kStart[MD_GROUPS] = {0, 3, 6, 9, 12, 15, 17, 19};

// ---- MultiDiscrete -----------------------------------------------------------------------------------------------
template <bool TRAIN>
__global__ void head_md_kernel(ZView z, int64_t M, const float* __restrict__ actions, int64_t ld_act,
                               const float* __restrict__ old_logp, const float* __restrict__ adv, float inv_batch,
                               float clip, float ent_coef, DzView dz, int dz_cols, float* __restrict__ logp_out,
                               float* __restrict__ metrics, uint64_t seed, uint64_t offset, int deterministic,
                               const float* __restrict__ u_inject, float* __restrict__ actions_out, int64_t ld_aout) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = row < M;
    float logp = 0.f, ent = 0.f;
    float lsm[21], pr[21], hg[MD_GROUPS];
    int act[MD_GROUPS];
    if (ok) {
        uint4 rnd[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        if (!TRAIN && !deterministic && u_inject == nullptr) {
            const uint64_t ctr = offset + (uint64_t)row;
            const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
            rnd[0] = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u), key);
            rnd[1] = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 1u, 0u), key);
        }
        for (int g = 0; g < MD_GROUPS; ++g) {
            const int s = kStart[g], nb = kBins[g];
            float zz[3], mx = -INFINITY;
            for (int j = 0; j < nb; ++j) {
                zz[j] = zload(z, row, s + j);
                mx = fmaxf(mx, zz[j]);
            }
            float se = 0.f;
            for (int j = 0; j < nb; ++j) se += expf(zz[j] - mx);
            const float lse = mx + logf(se);                   // Categorical(logits=...) normalisation
            float h = 0.f;
            for (int j = 0; j < nb; ++j) {
                lsm[s + j] = zz[j] - lse;
                pr[s + j] = expf(lsm[s + j]);                  // logits_to_probs
                h -= lsm[s + j] * pr[s + j];
            }
            hg[g] = h;
            ent += h;                                          // MultiDiscreteRolv.entropy: sum over the 8 (:121)
            int a;
            if (TRAIN) {
                a = (int)actions[row * ld_act + g];
                a = min(max(a, 0), nb - 1);
            } else if (deterministic) {
                a = 0;
                for (int j = 1; j < nb; ++j) a = zz[j] > zz[a] ? j : a;      // argmax, first of ties (:60-66)
            } else {
                const uint32_t rr[8] = {rnd[0].x, rnd[0].y, rnd[0].z, rnd[0].w, rnd[1].x, rnd[1].y, rnd[1].z, rnd[1].w};
                const float u = u_inject != nullptr ? u_inject[row * MD_GROUPS + g] : u01(rr[g]);
                float run = 0.f;
                a = nb - 1;
                bool found = false;
                for (int j = 0; j < nb; ++j) {
                    run += pr[s + j];
                    if (!found && run > u) {
                        found = true;
                        a = j;
                    }
                }
            }
            act[g] = a;
            logp += lsm[s + a];                                // MultiDiscreteRolv.log_prob: sum over the 8 (:115)
        }
    }
    if (!TRAIN) {
        if (ok) {
            for (int g = 0; g < MD_GROUPS; ++g) actions_out[row * ld_aout + g] = (float)act[g];
            if (logp_out) logp_out[row] = logp;
        }
        return;
    }
    PpoRow r = {0.f, 0.f, 0.f, 0.f};
    if (ok) {
        r = ppo_row(logp, old_logp[row], adv[row], clip, inv_batch);
        // d(ppo_loss)/dz_i = d_logp (onehot_i - p_i) + ent_coef w p_i (lsm_i + H_g)   [entropy.mean() over the minibatch]
        const float ce = ent_coef * inv_batch;
        for (int g = 0; g < MD_GROUPS; ++g) {
            const int s = kStart[g], nb = kBins[g];
            for (int j = 0; j < nb; ++j) {
                const float onehot = j == act[g] ? 1.f : 0.f;
                dzstore(dz, row, s + j, r.d_logp * (onehot - pr[s + j]) + ce * pr[s + j] * (lsm[s + j] + hg[g]));
            }
        }
        for (int c = 21; c < dz_cols; ++c) dzstore(dz, row, c, 0.f);
        if (logp_out) logp_out[row] = logp;
    }
    metrics_add(metrics, ent, r, ok);
}

// ---- Continuous ----------------------------------------------------------------------------------------------------
constexpr int MAX_CONT = 32;     // actions (2n <= 64 outputs)
template <bool TRAIN>
__global__ void head_ct_kernel(ZView z, int64_t M, int n_act, float var_m, float var_b,
                               const float* __restrict__ actions, int64_t ld_act, const float* __restrict__ old_logp,
                               const float* __restrict__ adv, float inv_batch, float clip, float ent_coef, DzView dz,
                               int dz_cols, float* __restrict__ logp_out, float* __restrict__ metrics, uint64_t seed,
                               uint64_t offset, int deterministic, const float* __restrict__ n_inject,
                               float* __restrict__ actions_out, int64_t ld_aout) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = row < M;
    const float kTwoPi = 6.283185307179586f;
    float logp = 0.f, ent = 0.f;
    PpoRow r = {0.f, 0.f, 0.f, 0.f};
    if (ok) {
        // pass 1: log-prob (the reference's four terms as written, continuous_policy.py:52-59) and entropy
        for (int j = 0; j < n_act; ++j) {
            const float tm = tanhf(zload(z, row, j)), ts = tanhf(zload(z, row, n_act + j));     // nn.Tanh (:37)
            const float mean = tm, sd = ts * var_m + var_b;                                      // torch_functions.py:30-33
            float x;
            if (TRAIN) {
                x = actions[row * ld_act + j];
            } else if (deterministic) {
                x = mean;
            } else {
                float nrm;
                if (n_inject != nullptr) {
                    nrm = n_inject[row * n_act + j];
                } else {
                    const uint64_t ctr = offset + (uint64_t)row;
                    const uint4 q = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)(j >> 1), 0u),
                                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
                    // Box-Muller: two normals per counter, one per action of the pair
                    const float u1 = fmaxf(u01(q.x), 1.0f / 16777216.0f), u2 = u01(q.y);
                    const float rad = sqrtf(-2.0f * logf(u1));
                    nrm = (j & 1) ? rad * sinf(kTwoPi * u2) : rad * cosf(kTwoPi * u2);
                }
                x = fminf(fmaxf(mean + sd * nrm, -1.0f), 1.0f);                                  // Normal.sample().clamp (:92)
            }
            if (!TRAIN) actions_out[row * ld_aout + j] = x;
            const float msq = mean * mean, ssq = sd * sd, xsq = x * x;
            const float t1 = -(msq / (2.0f * ssq)), t2 = (mean * x) / ssq, t3 = -(xsq / (2.0f * ssq));
            const float t4 = logf(1.0f / sqrtf(kTwoPi * ssq));
            logp += ((t1 + t2) + t3) + t4;
            ent += 0.5f + 0.5f * logf(kTwoPi) + logf(sd);                                        // Normal.entropy()
        }
        ent /= (float)n_act;           // entropy.mean() over all mb * n elements: the row's share (:117-118)
        if (!TRAIN) {
            if (logp_out) logp_out[row] = (deterministic ? 0.f : logp);
            return;
        }
        r = ppo_row(logp, old_logp[row], adv[row], clip, inv_batch);
        // pass 2: backward.  d_entropy = -ent_coef * mb * w spread over mb * n elements -> -ent_coef * w / n each
        const float ce = -ent_coef * inv_batch / (float)n_act;
        for (int j = 0; j < n_act; ++j) {
            const float tm = tanhf(zload(z, row, j)), ts = tanhf(zload(z, row, n_act + j));
            const float mean = tm, sd = ts * var_m + var_b, ssq = sd * sd;
            const float x = actions[row * ld_act + j];
            const float diff = x - mean;
            const float d_mean = r.d_logp * diff / ssq;
            const float d_std = r.d_logp * (diff * diff / (ssq * sd) - 1.0f / sd) + ce / sd;
            dzstore(dz, row, j, d_mean * (1.0f - tm * tm));
            dzstore(dz, row, n_act + j, d_std * var_m * (1.0f - ts * ts));
        }
        for (int c = 2 * n_act; c < dz_cols; ++c) dzstore(dz, row, c, 0.f);
        if (logp_out) logp_out[row] = logp;
    } else if (!TRAIN) {
        return;
    }
    metrics_add(metrics, ent, r, ok);
}

int check_z(const uint16_t* z, int64_t ldz, int parts, int64_t pstride, int cols) {
    RLPPO_CHECK_ARG(z != nullptr && parts >= 1 && parts <= 3 && (parts == 1 || pstride >= cols) &&
                        ldz >= (int64_t)(parts - 1) * pstride + cols, "bad logits view");
    return RLPPO_OK;
}

}  // namespace

extern "C" {

int rlppo_head_multi_discrete_train(const uint16_t* z, int64_t ldz, int z_parts, int64_t z_pstride, int64_t M,
                                    const float* actions, int64_t ld_act, const float* old_logp, const float* adv,
                                    float inv_batch, float clip, float ent_coef, uint16_t* dz, int64_t lddz, int dz_parts,
                                    int64_t dz_pstride, int dz_cols, float* logp_out, float* metrics, void* stream) {
    RLPPO_REQUIRE_DEVICE();
    int rc = check_z(z, ldz, z_parts, z_pstride, 21);
    if (rc) return rc;
    RLPPO_CHECK_ARG(actions && old_logp && adv && dz && M >= 1 && ld_act >= 8, "bad argument");
    RLPPO_CHECK_ARG(dz_parts >= 1 && dz_parts <= 3 && dz_cols >= 21 && (dz_parts == 1 || dz_pstride >= dz_cols) &&
                        lddz >= (int64_t)(dz_parts - 1) * dz_pstride + dz_cols, "bad d(logits) view");
    const ZView zv{z, ldz, z_parts, z_pstride};
    const DzView dv{dz, lddz, dz_parts, dz_pstride};
    head_md_kernel<true><<<(unsigned)((M + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        zv, M, actions, ld_act, old_logp, adv, inv_batch, clip, ent_coef, dv, dz_cols, logp_out, metrics, 0, 0, 0, nullptr,
        nullptr, 0);
    RLPPO_LAUNCH_CHECK();
    return RLPPO_OK;
}

int rlppo_head_multi_discrete_sample(const uint16_t* z, int64_t ldz, int z_parts, int64_t z_pstride, int64_t M,
                                     const float* u_inject, uint64_t seed, uint64_t offset, int deterministic,
                                     float* actions_out, int64_t ld_aout, float* logp_out, void* stream) {
    RLPPO_REQUIRE_DEVICE();
    int rc = check_z(z, ldz, z_parts, z_pstride, 21);
    if (rc) return rc;
    RLPPO_CHECK_ARG(actions_out && M >= 1 && ld_aout >= 8, "bad argument");
    const ZView zv{z, ldz, z_parts, z_pstride};
    const DzView dv{nullptr, 0, 0, 0};
    head_md_kernel<false><<<(unsigned)((M + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        zv, M, nullptr, 0, nullptr, nullptr, 0.f, 0.f, 0.f, dv, 0, logp_out, nullptr, seed, offset, deterministic, u_inject,
        actions_out, ld_aout);
    RLPPO_LAUNCH_CHECK();
    return RLPPO_OK;
}

int rlppo_head_continuous_train(const uint16_t* z, int64_t ldz, int z_parts, int64_t z_pstride, int64_t M, int n_act,
                                float var_min, float var_max, const float* actions, int64_t ld_act,
                                const float* old_logp, const float* adv, float inv_batch, float clip, float ent_coef,
                                uint16_t* dz, int64_t lddz, int dz_parts, int64_t dz_pstride, int dz_cols, float* logp_out,
                                float* metrics, void* stream) {
    RLPPO_REQUIRE_DEVICE();
    RLPPO_CHECK_ARG(n_act >= 1 && n_act <= MAX_CONT, "continuous head: 1..%d actions", MAX_CONT);
    int rc = check_z(z, ldz, z_parts, z_pstride, 2 * n_act);
    if (rc) return rc;
    RLPPO_CHECK_ARG(actions && old_logp && adv && dz && M >= 1 && ld_act >= n_act, "bad argument");
    RLPPO_CHECK_ARG(dz_parts >= 1 && dz_parts <= 3 && dz_cols >= 2 * n_act && (dz_parts == 1 || dz_pstride >= dz_cols) &&
                        lddz >= (int64_t)(dz_parts - 1) * dz_pstride + dz_cols, "bad d(logits) view");
    const float m = (var_max - var_min) / 2.0f;          // torch_functions.py:27-28
    const ZView zv{z, ldz, z_parts, z_pstride};
    const DzView dv{dz, lddz, dz_parts, dz_pstride};
    head_ct_kernel<true><<<(unsigned)((M + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        zv, M, n_act, m, var_min + m, actions, ld_act, old_logp, adv, inv_batch, clip, ent_coef, dv, dz_cols, logp_out,
        metrics, 0, 0, 0, nullptr, nullptr, 0);
    RLPPO_LAUNCH_CHECK();
    return RLPPO_OK;
}

int rlppo_head_continuous_sample(const uint16_t* z, int64_t ldz, int z_parts, int64_t z_pstride, int64_t M, int n_act,
                                 float var_min, float var_max, const float* n_inject, uint64_t seed, uint64_t offset,
                                 int deterministic, float* actions_out, int64_t ld_aout, float* logp_out, void* stream) {
    RLPPO_REQUIRE_DEVICE();
    RLPPO_CHECK_ARG(n_act >= 1 && n_act <= MAX_CONT, "continuous head: 1..%d actions", MAX_CONT);
    int rc = check_z(z, ldz, z_parts, z_pstride, 2 * n_act);
    if (rc) return rc;
    RLPPO_CHECK_ARG(actions_out && M >= 1 && ld_aout >= n_act, "bad argument");
    const float m = (var_max - var_min) / 2.0f;
    const ZView zv{z, ldz, z_parts, z_pstride};
    const DzView dv{nullptr, 0, 0, 0};
    head_ct_kernel<false><<<(unsigned)((M + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        zv, M, n_act, m, var_min + m, nullptr, 0, nullptr, nullptr, 0.f, 0.f, 0.f, dv, 0, logp_out, nullptr, seed, offset,
        deterministic, n_inject, actions_out, ld_aout);
    RLPPO_LAUNCH_CHECK();
    return RLPPO_OK;
}
}
