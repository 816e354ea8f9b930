// resample_to_44100 (streamz-rs/src/lib.rs:186-209) for the common rational ratios: "lane = row" polyphase kernel.
//
// Same specification and the same bits as resample_kernel<D> in frontend.cu (DESIGN.md "Resampler"):
//   y[j] = sum_{t<16} c[p][t] * x[i0 - 7 + t],   j M = i0 L + p,   acc = fma(c[p][t], float(x), acc) for t = 0..15 in float32,
//   inputs outside the clip are 0, result clamped to [-32768, 32767] and truncated toward zero (lib.rs:205-208).
//
// Mapping.  The ratio repeats with period (L outputs, M inputs).  The output stream of a clip is cut into rows of MULT
// periods, MULT chosen so that a row is a multiple of 16 bytes; a warp owns 32 consecutive rows and LANE = ROW, so all 32
// lanes work on the same output phase at the same time:
//   * the 16 taps of an output are warp-uniform: one shared table [L][16] per CTA, read with broadcast 128-bit loads
//     (4 shared-memory wavefronts per output for the whole warp instead of 5.7 per output and thread in the generic
//     kernel).  Everything that selects the taps is arithmetic on loop counters with compile-time L and M, so it lives in
//     uniform registers.  (A first version kept the table in the constant bank to feed the FFMA a uniform operand: the
//     28 KB table streamed through once per row thrashes the constant cache and was 5x SLOWER -- measured, abandoned.)
//   * each lane keeps its 16-sample input window in registers: walking the row advances the window by one sample per
//     input step (register renaming after unrolling by 16), so every input sample is read from shared memory once per
//     row instead of once per output tap;
//   * the inputs of one period of the 32 rows are staged in shared memory as i16 (128-bit global loads, odd word pitch:
//     conflict-free for lane = row); outputs go through a small per-lane ring in shared memory (column-major words, so
//     lane l only ever touches bank l) and leave as 128-bit stores, 64 samples per row at a time, straight from the lane
//     that computed them.  Warps are autonomous: no block-wide barrier after the table is loaded.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "frontend.cuh"
#include "tables.hpp"

namespace szb {

constexpr int kRsRows = 32;          // rows per warp tile == lanes
constexpr int kRsBlk = 32;           // outputs per lane and flush (64 bytes = 2 full sectors per row)
constexpr int kRsRing = 2 * kRsBlk;  // outputs per lane in the ring
constexpr int kRsRingPitch = kRsRing + 2;   // halfwords per lane: 33 words, odd -> lane l + word w hits bank (l + w) mod 32

struct RsArgs {
    const int16_t* in;
    const unsigned long long* in_off;
    const unsigned long long* out_off;
    const float* taps;               // [L][16] in output order q of a period: c[(q M) mod L][t] (reduced ratio)
    int16_t* out;
    uint32_t n_clips, tiles_per_clip, rate;
};

// Compile-time geometry: period (L outputs, M inputs; need not be reduced), rows of MULT periods.
template <int L_, int M_, int MULT_>
struct RsCfg {
    static constexpr int L = L_, M = M_, MULT = MULT_, Lb = L_ * MULT_, Mb = M_ * MULT_;
    static constexpr int SB = (M % 80 == 0) ? 80 : 64;                             // input steps per staging block
    static constexpr int span = SB + 16;                                           // staged samples per row and block: x[row + SB blk - 8 ..]
    static constexpr int pitch0 = (span % 2 == 0) ? span : span + 1;               // even pitch ...
    static constexpr int pitch = ((pitch0 / 2) % 2 == 0) ? pitch0 + 2 : pitch0;    // ... with an odd number of words per row
    static constexpr int in_bytes = (kRsRows * pitch * 2 + 15) / 16 * 16;
    static constexpr int ring_bytes = (kRsRingPitch * 2 * kRsRows + 15) / 16 * 16;
    static constexpr int per_warp = in_bytes + ring_bytes;
    static constexpr int taps_bytes = L * kResTaps * 4;
    static constexpr int warps0 = (224 * 1024 - taps_bytes) / per_warp;
    static constexpr int warps = warps0 > 16 ? 16 : warps0;
    static constexpr int chunks = span / 8;                                        // 16-byte chunks per staged row
    // outputs per input step: n(u) = ceil((u + 1) L / M) - ceil(u L / M) = kBase + [rem >= kThr], rem = (u L + M - 1) mod M
    static constexpr int kBase = L / M, kInc = L % M, kThr = M - kInc;
    static_assert(SB % 16 == 0 && Mb % SB == 0 && Lb % 8 == 0 && span % 8 == 0 && warps >= 4 && kBase + 1 <= 6, "unsupported row geometry");
};

__device__ __forceinline__ int16_t rs_quantise(float acc) {
    return int16_t(__float2int_rz(fminf(fmaxf(acc, -32768.f), 32767.f)));   // lib.rs:205-208
}

// NC consecutive outputs of the current input step, as NC interleaved FFMA chains over the register window; their taps
// are rows tp[0 .. NC) of the table (uniform address: one broadcast wavefront per 128-bit load); results go to the lane's
// ring at positions jr .. jr + NC - 1 (mod 128).
template <int NC, int U>
__device__ __forceinline__ void rs_emit(const float4* __restrict__ tp, const float (&xw)[16], uint32_t jr, int16_t* ring) {
    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.f;
#pragma unroll
    for (int t4 = 0; t4 < kResTaps / 4; ++t4) {
        float4 w[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) w[c] = tp[c * (kResTaps / 4) + t4];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fmaf(w[c].x, xw[(U + 4 * t4 + 0) & 15], acc[c]);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fmaf(w[c].y, xw[(U + 4 * t4 + 1) & 15], acc[c]);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fmaf(w[c].z, xw[(U + 4 * t4 + 2) & 15], acc[c]);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fmaf(w[c].w, xw[(U + 4 * t4 + 3) & 15], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) ring[(jr + c) & (kRsRing - 1)] = rs_quantise(acc[c]);
}

// Writes `count` outputs starting at row position j0 (a multiple of 32) from the lane's ring to its row: 4 x 128-bit
// stores when a whole block is inside the clip, scalar stores for the last partial block of a row or clip.
__device__ __noinline__ void rs_flush(const int16_t* ring, int16_t* yrow, uint32_t j0, uint32_t count, uint32_t valid) {
    if (j0 >= valid) return;
    const int16_t* rb = ring + (j0 & (kRsRing - 1));
    if (count == kRsBlk && j0 + kRsBlk <= valid) {
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(rb);
#pragma unroll
        for (int v = 0; v < kRsBlk / 8; ++v) {
            uint4 pk;
            pk.x = rw[4 * v + 0]; pk.y = rw[4 * v + 1]; pk.z = rw[4 * v + 2]; pk.w = rw[4 * v + 3];
            *reinterpret_cast<uint4*>(yrow + j0 + 8 * v) = pk;
        }
    } else {
        const uint32_t n = min(count, valid - j0);
        for (uint32_t i = 0; i < n; ++i) yrow[j0 + i] = rb[i];
    }
}

// Running state of the row walk (all warp-uniform): j = outputs emitted so far in this row, tp = taps row of the next
// output, rem = (u L + M - 1) mod M of the current input step u.
struct RsWalk {
    uint32_t j, rem;
    const float4* tp;
    const float4* tp_end;
};

// One input step: slide the window by one sample, then emit the outputs whose window starts here,
//   {q : floor(q M / L) == u}: kBase of them, one more when rem >= kThr (Bresenham on the ratio, no division).
template <class Cfg, int U>
__device__ __forceinline__ void rs_step(RsWalk& wk, float (&xw)[16], const int16_t*& wp, int16_t* ring, int16_t* yrow, uint32_t valid) {
    xw[(U + 15) & 15] = float(*wp++);                              // newest sample of the window that starts at this step
    const bool extra = wk.rem >= uint32_t(Cfg::kThr);
    wk.rem = extra ? wk.rem - uint32_t(Cfg::kThr) : wk.rem + uint32_t(Cfg::kInc);
    const uint32_t n = uint32_t(Cfg::kBase) + (extra ? 1u : 0u);
    const uint32_t j0 = wk.j;
    if (Cfg::kBase >= 3) {
        uint32_t k = 0;
        for (; k + 3 <= n; k += 3) rs_emit<3, U>(wk.tp + k * (kResTaps / 4), xw, j0 + k, ring);
        if (k + 2 == n) rs_emit<2, U>(wk.tp + k * (kResTaps / 4), xw, j0 + k, ring);
        else if (k + 1 == n) rs_emit<1, U>(wk.tp + k * (kResTaps / 4), xw, j0 + k, ring);
    } else if (Cfg::kBase == 2) {
        if (extra) rs_emit<3, U>(wk.tp, xw, j0, ring); else rs_emit<2, U>(wk.tp, xw, j0, ring);
    } else if (Cfg::kBase == 1) {
        if (extra) rs_emit<2, U>(wk.tp, xw, j0, ring); else rs_emit<1, U>(wk.tp, xw, j0, ring);
    } else {
        if (extra) rs_emit<1, U>(wk.tp, xw, j0, ring);
    }
    wk.tp += n * (kResTaps / 4);
    if (wk.tp == wk.tp_end) wk.tp -= Cfg::L * (kResTaps / 4);      // next period
    wk.j = j0 + n;
    if (((wk.j ^ j0) & uint32_t(kRsBlk)) != 0)                      // a block of 32 outputs just completed (uniform)
        rs_flush(ring, yrow, (wk.j & ~uint32_t(kRsBlk - 1)) - kRsBlk, kRsBlk, valid);
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::warps * 32, 1)
resample_rows_kernel(const __grid_constant__ RsArgs a) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    constexpr uint32_t L = Cfg::L, M = Cfg::M, Lb = Cfg::Lb, Mb = Cfg::Mb, P = Cfg::pitch;
    // the warp index goes through a shuffle so that the compiler knows everything derived from it is warp-uniform (the
    // uniform datapath is only available in code it can prove convergent)
    const uint32_t lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    float4* s_taps = reinterpret_cast<float4*>(rs_smem);
    unsigned char* base = rs_smem + Cfg::taps_bytes + size_t(warp) * Cfg::per_warp;
    int16_t* s_in = reinterpret_cast<int16_t*>(base);
    int16_t* ring = reinterpret_cast<int16_t*>(base + Cfg::in_bytes) + lane * kRsRingPitch;
    for (uint32_t i = threadIdx.x; i < L * kResTaps / 4; i += blockDim.x) s_taps[i] = __ldg(reinterpret_cast<const float4*>(a.taps) + i);
    __syncthreads();
    const uint64_t n_items = uint64_t(a.n_clips) * a.tiles_per_clip;
    const uint64_t n_warps = uint64_t(gridDim.x) * Cfg::warps;

    for (uint64_t item = uint64_t(blockIdx.x) * Cfg::warps + warp; item < n_items; item += n_warps) {
        const uint32_t clip = uint32_t(item / a.tiles_per_clip), tile = uint32_t(item - uint64_t(clip) * a.tiles_per_clip);
        const int64_t n_in = __shfl_sync(0xffffffffu, int64_t(a.in_off[clip + 1] - a.in_off[clip]), 0);
        const uint64_t n_out = uint64_t(n_in) * 44100ull / a.rate;                    // lib.rs:196
        const uint64_t row0 = uint64_t(tile) * kRsRows;
        if (row0 * Lb >= n_out) continue;                                              // uniform
        const int16_t* x = a.in + a.in_off[clip];
        const uint64_t row = row0 + lane;
        const uint64_t jrow = row * Lb;
        int16_t* yrow = a.out + a.out_off[clip] + jrow;
        const uint32_t valid = jrow >= n_out ? 0u : uint32_t(min(uint64_t(Lb), n_out - jrow));
        const bool vec_in = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
        float xw[16];
        RsWalk wk;
        wk.j = 0;
        wk.rem = M - 1;                                                               // (u L + M - 1) mod M at u = 0
        wk.tp = s_taps;
        wk.tp_end = s_taps + L * (kResTaps / 4);

#pragma unroll 1
        for (uint32_t blk = 0; blk < Mb / uint32_t(Cfg::SB); ++blk) {
            // ---- stage x[Mb (row0 + rr) + SB blk - 8 + d], d < SB + 16, of the 32 rows at pitch P (zeros outside the clip) ----
            const int64_t g0 = int64_t(row0) * Mb + int64_t(blk) * Cfg::SB - 8;
            __syncwarp();
            for (uint32_t idx = lane; idx < kRsRows * uint32_t(Cfg::chunks); idx += 32) {
                const uint32_t rr = idx / uint32_t(Cfg::chunks), c = idx - rr * uint32_t(Cfg::chunks);
                const int64_t gi = g0 + int64_t(rr) * Mb + 8 * c;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (gi >= 0 && gi + 8 <= n_in && vec_in) {
                    v = __ldg(reinterpret_cast<const uint4*>(x + gi));
                } else if (gi + 8 > 0 && gi < n_in) {
                    uint32_t wv[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int64_t g = gi + 2 * e;
                        const uint32_t lo = (g >= 0 && g < n_in) ? uint32_t(uint16_t(__ldg(x + g))) : 0u;
                        const uint32_t hi = (g + 1 >= 0 && g + 1 < n_in) ? uint32_t(uint16_t(__ldg(x + g + 1))) : 0u;
                        wv[e] = lo | (hi << 16);
                    }
                    v = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                }
                uint32_t* d = reinterpret_cast<uint32_t*>(s_in + rr * P + 8 * c);    // rows are 4-byte aligned (P even)
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            }
            __syncwarp();

            // ---- lane = row: walk the block, one input step at a time ----
            const int16_t* wp = s_in + lane * P;
            if (blk == 0) {
#pragma unroll
                for (int t = 0; t < 15; ++t) xw[t] = float(wp[1 + t]);               // x[row - 7 .. row + 7]
                xw[15] = 0.f;
            }
            wp += 16;
#pragma unroll 1
            for (uint32_t u0 = 0; u0 < uint32_t(Cfg::SB); u0 += 16) {
                rs_step<Cfg, 0>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 1>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 2>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 3>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 4>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 5>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 6>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 7>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 8>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 9>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 10>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 11>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 12>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 13>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 14>(wk, xw, wp, ring, yrow, valid);
                rs_step<Cfg, 15>(wk, xw, wp, ring, yrow, valid);
            }
        }
        if (Lb % kRsBlk != 0) rs_flush(ring, yrow, Lb / kRsBlk * kRsBlk, Lb % kRsBlk, valid);   // the row's last, partial block
    }
}

// Host side: the taps of a rate in output order, uploaded once per context and rate.
template <class Cfg>
static szb_status launch_rows(szb_ctx* ctx, const int16_t* d_in, const uint64_t* d_in_off, const uint64_t* d_out_off,
                              uint32_t n_clips, uint64_t max_out, uint32_t rate, int16_t* d_out) {
    if (ctx->rows_taps_rate != rate) {
        uint32_t Lr, Mr;
        resample_ratio(rate, Lr, Mr);
        const auto c = resample_taps(rate);                        // [Lr][16]
        std::vector<float> t(size_t(Cfg::L) * kResTaps);
        for (uint32_t q = 0; q < uint32_t(Cfg::L); ++q) {
            const uint32_t ph = uint32_t((uint64_t(q) * Mr) % Lr);
            std::memcpy(&t[size_t(q) * kResTaps], &c[size_t(ph) * kResTaps], kResTaps * sizeof(float));
        }
        SZB_TRY(ctx->rows_taps.reserve(t.size() * sizeof(float)));
        SZB_CUDA(cudaMemcpyAsync(ctx->rows_taps.ptr, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        SZB_CUDA(cudaStreamSynchronize(ctx->stream));              // t is a temporary
        ctx->rows_taps_rate = rate;
    }
    RsArgs a;
    a.in = d_in;
    a.in_off = reinterpret_cast<const unsigned long long*>(d_in_off);
    a.out_off = reinterpret_cast<const unsigned long long*>(d_out_off);
    a.taps = ctx->rows_taps.as<float>();
    a.out = d_out;
    a.n_clips = n_clips;
    const uint64_t rows = (max_out + Cfg::Lb - 1) / Cfg::Lb;
    a.tiles_per_clip = uint32_t((rows + kRsRows - 1) / kRsRows);
    a.rate = rate;
    const uint64_t n_items = uint64_t(n_clips) * a.tiles_per_clip;
    // a warp tile is 32 rows (1.3 - 2.6 s of audio): below one tile per resident warp the generic kernel spreads better
    const uint64_t min_tiles = ctx->rows_min_tiles == ~0ull ? uint64_t(ctx->sm_count) * Cfg::warps : ctx->rows_min_tiles;
    if (n_items < min_tiles) return launch_resample_generic(ctx, d_in, d_in_off, d_out_off, n_clips, max_out, rate, d_out);
    const size_t smem = size_t(Cfg::taps_bytes) + size_t(Cfg::warps) * Cfg::per_warp;
    SZB_CUDA(cudaFuncSetAttribute(resample_rows_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const uint64_t want = (n_items + Cfg::warps - 1) / Cfg::warps;
    const uint32_t grid = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>(want, uint64_t(ctx->sm_count))));
    resample_rows_kernel<Cfg><<<grid, Cfg::warps * 32, smem, ctx->stream>>>(a);
    SZB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return SZB_OK;
}

szb_status launch_resample(szb_ctx* ctx, const int16_t* d_in, const uint64_t* d_in_off, const uint64_t* d_out_off,
                           uint32_t n_clips, uint64_t max_out, uint32_t rate, int16_t* d_out, bool out_aligned16) {
    if (n_clips == 0 || max_out == 0) return SZB_OK;
    // The row kernel stores 128-bit vectors: every clip's output must start on a 16-byte boundary (true for the batch
    // layout, whose resampled clips are padded to 8 samples, and for single clips).  Rows are 40-80 ms of audio and a
    // warp takes 32 of them: batches of short clips keep more lanes busy in the generic kernel.
#define SZB_RS_CASE(RATE, L, M, MULT) \
    case RATE: if (max_out >= 8ull * (L) * (MULT)) return launch_rows<RsCfg<L, M, MULT>>(ctx, d_in, d_in_off, d_out_off, n_clips, max_out, rate, d_out); break
    static const bool no_rows = getenv("SZB_NO_ROWS") != nullptr;   // EXPERIMENT ONLY
    if (out_aligned16 && !no_rows) switch (rate) {     // period (L, M) with 44100 / rate = L / M, M a multiple of 16; rows of MULT periods
        SZB_RS_CASE(8000, 441, 80, 8);
        SZB_RS_CASE(11025, 512, 128, 1);
        SZB_RS_CASE(12000, 294, 80, 4);
        SZB_RS_CASE(16000, 441, 160, 8);
        SZB_RS_CASE(22050, 256, 128, 2);
        SZB_RS_CASE(24000, 147, 80, 8);
        SZB_RS_CASE(32000, 441, 320, 8);
        SZB_RS_CASE(48000, 147, 160, 8);
        default: break;
    }
#undef SZB_RS_CASE
    return launch_resample_generic(ctx, d_in, d_in_off, d_out_off, n_clips, max_out, rate, d_out);
}

}  // namespace szb
