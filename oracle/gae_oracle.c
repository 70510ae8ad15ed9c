/*
 * CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/ref_oracle.py header).  Plain-C restatement of
 *   - compute_gae            rlgym_ppo/util/torch_functions.py:36-78   (NumPy>=2 rounding points)
 *   - WelfordRunningStat     rlgym_ppo/util/running_stats.py:37-46     (f64 samples, f32 state)
 *   - RandomState.permutation as called at rlgym_ppo/ppo/experience_buffer.py:98
 *       (NumPy legacy MT19937: numpy/random/src/mt19937 genrand + legacy rk_interval + Fisher-Yates)
 * Built by oracle/Makefile into oracle/_build/liboracle.so.  Never linked into the product library.
 */
#include <stdint.h>
#include <stddef.h>

void oracle_gae_nep50(const float* rew, const float* done, const double* trunc, const float* val,
                      int64_t n, double gamma, double lmbda, int has_std, float std,
                      float* adv, float* vtarget, double* ret) {
    const float gl32 = (float)(gamma * lmbda);           /* python float * np.float32 -> f32 */
    double last_gae = 0.0, last_ret = 0.0;
    for (int64_t t = n - 1; t >= 0; --t) {
        const float nd = 1.0f - done[t];                  /* :59 */
        const double nt = 1.0 - trunc[t];                 /* :60 */
        float nr = rew[t];
        if (has_std) {                                    /* :62-65 */
            nr = rew[t] / std;
            nr = nr < -10.0f ? -10.0f : nr;
            nr = nr > 10.0f ? 10.0f : nr;
        }
        const float gv = (float)(gamma * (double)val[t + 1]);
        const float pred = nr + gv * nd;                  /* :67 */
        const float delta = pred - val[t];                /* :68 */
        last_ret = (double)rew[t] + last_ret * gamma * (double)nd * nt;   /* :69 */
        ret[t] = last_ret;
        last_gae = (double)delta + (double)(gl32 * nd) * nt * last_gae;   /* :72 */
        adv[t] = (float)last_gae;                         /* :76 */
        vtarget[t] = (float)((double)val[t] + last_gae);  /* :77 */
    }
}

/* state: mean[dim] f32, m2[dim] f32, count int64; samples f64 [n, dim] applied one by one */
void oracle_welford_update(float* mean, float* m2, int64_t* count, const double* samples, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        const int64_t cc = *count;
        *count = cc + 1;
        const double delta = samples[i] - (double)mean[0];
        const double delta_n = delta / (double)(*count);
        mean[0] = (float)((double)mean[0] + delta_n);
        m2[0] = (float)((double)m2[0] + delta * delta_n * (double)cc);
    }
}

static uint32_t mt_next(uint32_t* mt, int32_t* pos) {
    if (*pos == 624) {
        for (int k = 0; k < 624; ++k) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        *pos = 0;
    }
    uint32_t y = mt[(*pos)++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

void oracle_mt19937_permutation(uint32_t* key, int32_t* pos, int64_t n, int64_t* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = i;
    for (int64_t i = n - 1; i > 0; --i) {
        uint64_t mx = (uint64_t)i, mask = mx, v;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4;
        mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
        if (mx <= 0xffffffffull) {
            do { v = mt_next(key, pos) & mask; } while (v > mx);
        } else {
            do { v = (((uint64_t)mt_next(key, pos) << 32) | mt_next(key, pos)) & mask; } while (v > mx);
        }
        int64_t tmp = out[i]; out[i] = out[v]; out[v] = tmp;
    }
}
