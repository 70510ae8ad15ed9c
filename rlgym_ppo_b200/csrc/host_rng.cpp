// Host-side NumPy-legacy permutation for the minibatch shuffle (experience_buffer.py:98).
// MT19937 (Matsumoto & Nishimura) + NumPy's legacy rk_interval masked rejection + Fisher-Yates from the top,
// i.e. what np.random.RandomState(seed).permutation(n) executes; the caller round-trips the generator state
// with RandomState.get_state()/set_state() so `ExperienceBuffer.rng` stays a live NumPy object.
// Sequential by construction (each swap depends on the stream position), so it runs on the host, ahead of the
// GPU work that consumes it.
#include <stdint.h>

#include "../../include/rlppo.h"

namespace {
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
    uint64_t mask = 0;
    for (int64_t i = n - 1; i > 0; --i) {
        const uint64_t mx = (uint64_t)i;
        // smallest 2^k-1 >= mx; recomputed only when the top bit drops
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
