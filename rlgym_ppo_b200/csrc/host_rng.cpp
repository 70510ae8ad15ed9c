// Host-side NumPy-legacy permutation for the minibatch shuffle (experience_buffer.py:98).
// MT19937 (Matsumoto & Nishimura) + NumPy's legacy rk_interval masked rejection + Fisher-Yates from the top,
// i.e. what np.random.RandomState(seed).permutation(n) executes; the caller round-trips the generator state
// with RandomState.get_state()/set_state() so `ExperienceBuffer.rng` stays a live NumPy object.
// Sequential by construction (each swap depends on the stream position), so it runs on the host, ahead of the
// GPU work that consumes it.
#include <stdint.h>

#include "../../include/rlppo.h"

#include <vector>

#if defined(__x86_64__) && defined(__GNUC__)
#define RLPPO_SIMD_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define RLPPO_SIMD_CLONES
#endif

namespace {
// MT19937 state regeneration + tempering of a whole block; the loops vectorise (AVX2 clone picked at load time)
RLPPO_SIMD_CLONES void mt_refill_block(uint32_t* key) {
    constexpr uint32_t kUpper = 0x80000000u, kLower = 0x7fffffffu, kMatrix = 0x9908b0dfu;
    int k = 0;
    for (; k < 624 - 397; ++k) {
        const uint32_t y = (key[k] & kUpper) | (key[k + 1] & kLower);
        key[k] = key[k + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrix);
    }
    for (; k < 623; ++k) {
        const uint32_t y = (key[k] & kUpper) | (key[k + 1] & kLower);
        key[k] = key[k + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrix);
    }
    const uint32_t y = (key[623] & kUpper) | (key[0] & kLower);
    key[623] = key[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrix);
}
RLPPO_SIMD_CLONES void mt_temper_block(const uint32_t* key, uint32_t* out, int from) {
    for (int k = from; k < 624; ++k) {
        uint32_t y = key[k];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        out[k] = y;
    }
}

struct MT {
    uint32_t* key;
    int32_t pos;
    inline void refill() {
        constexpr uint32_t kUpper = 0x80000000u, kLower = 0x7fffffffu, kMatrix = 0x9908b0dfu;
        int k = 0;
        for (; k < 624 - 397; ++k) {
            const uint32_t y = (key[k] & kUpper) | (key[k + 1] & kLower);
            key[k] = key[k + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrix);
        }
        for (; k < 623; ++k) {
            const uint32_t y = (key[k] & kUpper) | (key[k + 1] & kLower);
            key[k] = key[k + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrix);
        }
        const uint32_t y = (key[623] & kUpper) | (key[0] & kLower);
        key[623] = key[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & kMatrix);
        pos = 0;
    }
    inline uint32_t next() {
        if (pos == 624) refill();
        uint32_t y = key[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
};
}  // namespace

extern "C" int rlppo_host_permutation(uint32_t* h_key, int32_t* h_pos, int64_t n, int64_t* h_out) {
    if (!h_key || !h_pos || !h_out || n < 0 || *h_pos < 0 || *h_pos > 624) return RLPPO_ERR_ARG;
    MT mt{h_key, *h_pos};
    for (int64_t i = 0; i < n; ++i) h_out[i] = i;
    if (n < 2) return RLPPO_OK;
    if (n > 0xffffffffll) {
        // 64-bit draws: the plain loop (never reached by a replay buffer; kept for completeness of the NumPy contract)
        uint64_t mask = 0;
        for (int64_t i = n - 1; i > 0; --i) {
            const uint64_t mx = (uint64_t)i;
            if (mask == 0 || (mask >> 1) >= mx) {
                mask = mx;
                mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4;
                mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
            }
            uint64_t v;
            if (mx <= 0xffffffffull) {
                do { v = mt.next() & mask; } while (v > mx);
            } else {
                do { v = (((uint64_t)mt.next() << 32) | mt.next()) & mask; } while (v > mx);
            }
            const int64_t tmp = h_out[i];
            h_out[i] = h_out[v];
            h_out[v] = tmp;
        }
        *h_pos = mt.pos;
        return RLPPO_OK;
    }
    // Two decoupled passes (the draw sequence does not depend on the array being shuffled):
    //  (1) draw j_i for i = n-1 .. 1 from bulk-tempered MT19937 blocks with the masked rejection rule;
    //  (2) apply the swaps with the target of a later swap prefetched -- the shuffle is a chain of dependent random
    //      accesses into an array larger than L1, which is what made the one-pass loop slow.
    static thread_local std::vector<uint32_t> j_buf;
    if ((int64_t)j_buf.size() < n) j_buf.resize((size_t)n);
    uint32_t* j = j_buf.data();
    uint32_t tempered[624];
    int tpos = 624, tend = 624;          // tempered[tpos..tend) are valid outputs
    auto refill_tempered = [&]() {
        if (mt.pos == 624) {
            mt_refill_block(h_key);
            mt.pos = 0;
        }
        const int base = mt.pos;
        mt_temper_block(h_key, tempered, base);
        tpos = base;
        tend = 624;
    };
    // Masked rejection without an unpredictable branch: within one "level" (mask constant while i stays above
    // mask >> 1) every raw output is written to j[i] and i only moves on when the draw is accepted.
    int64_t i = n - 1;
    while (i > 0) {
        uint32_t mask = (uint32_t)i;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        const int64_t level_lo = (int64_t)(mask >> 1);      // this mask serves i in (level_lo, mask]
        while (i > level_lo) {
            if (tpos == tend) refill_tempered();
            int t = tpos;
            int64_t ii = i;
            const int te = tend;
            while (t < te && ii > level_lo) {
                const uint32_t v = tempered[t++] & mask;
                j[ii] = v;
                ii -= (int64_t)(v <= (uint32_t)ii);
            }
            tpos = t;
            mt.pos = t;
            i = ii;
        }
    }
    constexpr int64_t kAhead = 24;
    for (int64_t i = n - 1; i > 0; --i) {
        if (i > kAhead) __builtin_prefetch(&h_out[j[i - kAhead]], 1, 1);
        const uint32_t v = j[i];
        const int64_t tmp = h_out[i];
        h_out[i] = h_out[v];
        h_out[v] = tmp;
    }
    *h_pos = mt.pos;
    return RLPPO_OK;
}
