// Host side of the tensor-core resampler (resample_tc.cuh): which rates it takes, the chunk geometry and the Toeplitz
// blocks of the integer taps.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "resample_tc.cuh"
#include "../../../streamz_b200/csrc/tables.hpp"

namespace szb {

// Integer taps of the exactly reproducible resampler: the float taps x 2^15 rounded to nearest, the largest tap of each
// phase then adjusted so that the phase sums to exactly 2^15 (unit DC gain, a constant input comes out unchanged).
inline std::vector<int16_t> resample_taps_q(uint32_t rate) {
    uint32_t L, M;
    resample_ratio(rate, L, M);
    const std::vector<float> c = resample_taps(rate);
    std::vector<int16_t> q(size_t(L) * kResTaps);
    for (uint32_t p = 0; p < L; ++p) {
        long sum = 0;
        int big = 0;
        long v[kResTaps];
        for (int t = 0; t < kResTaps; ++t) {
            v[t] = std::lround(double(c[size_t(p) * kResTaps + t]) * 32768.0);
            sum += v[t];
            if (std::labs(v[t]) > std::labs(v[big])) big = t;
        }
        v[big] += 32768 - sum;
        for (int t = 0; t < kResTaps; ++t) q[size_t(p) * kResTaps + t] = int16_t(v[t]);
    }
    return q;
}

namespace rtc {

struct Plan {
    uint32_t rate = 0;
    bool ok = false;
    Geom geom{};
    std::vector<uint8_t> btiles;     // [n_chunks][2][32][128]
};

// The period must be expressible as (L outputs, M = 160 inputs) with L <= 448: 16 kHz (441), 22.05 kHz (320), 24 kHz (294),
// 48 kHz (147).  Other rates keep the CUDA-core kernel (same arithmetic, same bits).
inline bool make_plan(uint32_t rate, Plan& p) {
    p = Plan{};
    p.rate = rate;
    uint32_t Lr, Mr;
    resample_ratio(rate, Lr, Mr);
    if (kM % Mr != 0) return false;
    const uint32_t mult = kM / Mr, L = Lr * mult;
    if (L > uint32_t(kMaxChunks * kNC) || L < uint32_t(kNC)) return false;
    const std::vector<int16_t> cq = resample_taps_q(rate);        // [Lr][16]
    Geom& g = p.geom;
    g.L = int(L);
    g.n_chunks = int((L + kNC - 1) / kNC);
    p.btiles.assign(size_t(g.n_chunks) * 2 * kBTile, 0);
    for (int n = 0; n < g.n_chunks; ++n) {
        const int q_lo = n * kNC, q_hi = int(std::min<uint32_t>(L, uint32_t(q_lo + kNC)));
        // window of output q covers d = o_q + 1 .. o_q + 16 (row origin is x[M r - 8], taps start at x[M r + o_q - 7])
        const int d_min = int(uint64_t(q_lo) * kM / L) + 1, d_max = int(uint64_t(q_hi - 1) * kM / L) + 16;
        g.ks_first[n] = d_min / 32;
        g.ks_count[n] = d_max / 32 - d_min / 32 + 1;
        if (g.ks_count[n] > 4 || d_max >= kKSteps * 32) return false;
        for (int q = q_lo; q < q_hi; ++q) {
            const uint64_t pos = uint64_t(q) * kM;               // = i0 L + phase, in units of the (L, 160) period
            const int o = int(pos / L);
            const uint32_t ph = uint32_t((pos % L) / mult);      // phase of the reduced ratio
            for (int t = 0; t < kResTaps; ++t) {
                const int d = o + 1 + t, slot = d / 32 - g.ks_first[n];
                const int16_t c = cq[size_t(ph) * kResTaps + t];
                const size_t at = size_t(q - q_lo) * 128 + size_t(slot) * 32 + size_t(d % 32);
                p.btiles[(size_t(2 * n) * kBTile) + at] = uint8_t(int8_t(c >> 8));            // high byte, signed (floor)
                p.btiles[(size_t(2 * n + 1) * kBTile) + at] = uint8_t(c & 0xFF);              // low byte, unsigned
            }
        }
    }
    p.ok = true;
    return true;
}

}  // namespace rtc
}  // namespace szb
