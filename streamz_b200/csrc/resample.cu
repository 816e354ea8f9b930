// resample_to_44100 (streamz-rs/src/lib.rs:186-209) for the common rational ratios: "lane = row" polyphase kernel.
//
// Same specification and the same bits as resample_kernel<D> in frontend.cu (DESIGN.md "Resampler"):
//   y[j] = sum_{t<16} c[p][t] * x[i0 - 7 + t],   j M = i0 L + p,   acc = fma(c[p][t], float(x), acc) for t = 0..15 in float32,
//   inputs outside the clip are 0, result clamped to [-32768, 32767] and truncated toward zero (lib.rs:205-208).
//
// Mapping.  The output stream of a clip is cut into rows of Lb = mult * L samples (mult * M input samples each); the
// phase and the input offset of output q of a row are the same in every row.  A warp owns 32 consecutive rows and LANE =
// ROW, so all 32 lanes work on the same q at the same time:
//   * the 16 taps of an output are warp-uniform -- they live in the kernel parameter block (constant bank), are fetched
//     with uniform loads and feed the FFMA as a uniform operand.  A three-register FFMA issues at half the rate of one
//     with a uniform/constant operand on this machine, and the FFMA chain is what bounds this kernel;
//   * each lane keeps its 16-sample input window in registers: walking the row advances the window by one sample per
//     input step (register renaming after unrolling by 16), so every input sample is read from shared memory once per
//     row instead of once per output tap;
//   * the warp's inputs are staged once in shared memory as i16 with an odd word pitch per row (conflict-free for
//     lane = row); its 32 x Lb outputs are one contiguous, 16-byte aligned span of the output stream, staged in shared
//     memory and written with 128-bit coalesced stores.  Warps are autonomous: no block-wide barrier anywhere.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "frontend.cuh"
#include "tables.hpp"

namespace szb {

constexpr int kRsMaxLb = 448;        // outputs per row (rows of the taps table)
constexpr int kRsRows = 32;          // rows per warp tile == lanes

struct RsTables {                    // 28 672 bytes of kernel parameters (limit 32 764)
    float taps[kRsMaxLb * kResTaps]; // [q][t]: taps of output q of a row, = c[(q M) mod L][t]
};

struct RsArgs {
    const int16_t* in;
    const unsigned long long* in_off;
    const unsigned long long* out_off;
    int16_t* out;
    uint32_t n_clips, tiles_per_clip, rate;
};

// Compile-time geometry of one ratio L / M with rows of MULT periods.
template <int L_, int M_, int MULT_>
struct RsCfg {
    static constexpr int L = L_, M = M_, Lb = L_ * MULT_, Mb = M_ * MULT_;
    static constexpr int span = Mb + kResTaps;                                     // samples a row reads: x[Mb r - 7 .. Mb r + Mb + 8]
    static constexpr int pitch0 = (span % 2 == 0) ? span : span + 1;               // even pitch ...
    static constexpr int pitch = ((pitch0 / 2) % 2 == 0) ? pitch0 + 2 : pitch0;    // ... with an odd number of words per row
    static constexpr int in_bytes = (kRsRows * pitch * 2 + 15) / 16 * 16;
    static constexpr int out_bytes = kRsRows * Lb * 2;                             // multiple of 64
    static constexpr int per_warp = in_bytes + out_bytes;
    static constexpr int warps0 = (226 * 1024) / per_warp;
    static constexpr int warps = warps0 > 8 ? 8 : warps0;
    static_assert(Lb <= kRsMaxLb && Mb >= 16 && warps >= 2, "unsupported row geometry");
};

__device__ __forceinline__ int16_t rs_quantise(float acc) {
    return int16_t(__float2int_rz(fminf(fmaxf(acc, -32768.f), 32767.f)));   // lib.rs:205-208
}

// NC outputs (q .. q + NC - 1) of the current input step, as NC interleaved FFMA chains over the register window.
template <int NC, int U>
__device__ __forceinline__ void rs_emit(const RsTables& T, const float (&xw)[16], uint32_t q, int16_t*& so) {
    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.f;
#pragma unroll
    for (int t = 0; t < kResTaps; ++t) {
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fmaf(T.taps[(q + c) * kResTaps + t], xw[(U + t) & 15], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) so[c] = rs_quantise(acc[c]);
    so += NC;
}

// One input step s: slide the window by one sample, then emit the outputs whose window starts here,
//   {q : floor(q M / L) == s} = [ceil(s L / M), ceil((s + 1) L / M)).
// s and q are ARITHMETIC on the loop counter with compile-time L and M and touch nothing lane-dependent (the staging
// read pointer `wp` and the output pointer `so` advance on their own), so ptxas keeps them in uniform registers, fetches
// the taps with uniform loads (LDCU c[0x0][UR + imm]) and feeds them to the FFMA as uniform operands.
template <class Cfg, int U>
__device__ __forceinline__ void rs_step(const RsTables& T, float (&xw)[16], const int16_t*& wp, uint32_t s, int16_t*& so) {
    if (Cfg::Mb % 16 != 0 && s >= uint32_t(Cfg::Mb)) return;      // uniform
    xw[(U + 15) & 15] = float(*wp++);                              // newest sample (index s + 15) of the window that starts at step s
    uint32_t q = (s * uint32_t(Cfg::L) + uint32_t(Cfg::M - 1)) / uint32_t(Cfg::M);
    const uint32_t q_end = ((s + 1) * uint32_t(Cfg::L) + uint32_t(Cfg::M - 1)) / uint32_t(Cfg::M);
    while (q + 3 <= q_end) { rs_emit<3, U>(T, xw, q, so); q += 3; }
    if (q + 2 == q_end) rs_emit<2, U>(T, xw, q, so);
    else if (q + 1 == q_end) rs_emit<1, U>(T, xw, q, so);
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::warps * 32, 1)
resample_rows_kernel(const __grid_constant__ RsTables T, const __grid_constant__ RsArgs a) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    constexpr uint32_t Lb = Cfg::Lb, Mb = Cfg::Mb, P = Cfg::pitch;
    // the warp index goes through a shuffle so that the compiler knows everything derived from it is warp-uniform (the
    // uniform datapath is only available in code it can prove convergent)
    const uint32_t lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    unsigned char* base = rs_smem + size_t(warp) * Cfg::per_warp;
    int16_t* s_in = reinterpret_cast<int16_t*>(base);
    int16_t* s_out = reinterpret_cast<int16_t*>(base + Cfg::in_bytes);
    const uint64_t n_items = uint64_t(a.n_clips) * a.tiles_per_clip;
    const uint64_t n_warps = uint64_t(gridDim.x) * Cfg::warps;

    for (uint64_t item = uint64_t(blockIdx.x) * Cfg::warps + warp; item < n_items; item += n_warps) {
        const uint32_t clip = uint32_t(item / a.tiles_per_clip), tile = uint32_t(item - uint64_t(clip) * a.tiles_per_clip);
        const int64_t n_in = __shfl_sync(0xffffffffu, int64_t(a.in_off[clip + 1] - a.in_off[clip]), 0);
        const uint64_t n_out = uint64_t(n_in) * 44100ull / a.rate;                    // lib.rs:196
        const uint64_t j_lo = uint64_t(tile) * kRsRows * Lb;
        if (j_lo >= n_out) continue;                                                   // uniform
        const uint64_t j_hi = min(n_out, j_lo + uint64_t(kRsRows) * Lb);
        const int16_t* x = a.in + a.in_off[clip];
        int16_t* y = a.out + a.out_off[clip];

        // ---- stage the inputs of rows [32 tile, 32 tile + 32): row rr holds the Mb + 16 samples x[Mb (r0 + rr) - 7 + d] it
        //      reads, at pitch P (neighbouring rows overlap by 16 samples; keeping each row contiguous keeps the read pointer of
        //      the row walk a plain increment) ----
        const int64_t g0 = int64_t(tile) * kRsRows * Mb - (kResTaps / 2 - 1);
        for (uint32_t rr = 0; rr < kRsRows; ++rr) {
            const int64_t gr = g0 + int64_t(rr) * Mb;
            for (uint32_t d0 = 0; d0 < uint32_t(Cfg::span); d0 += 32) {
                const uint32_t d = d0 + lane;
                const int64_t gi = gr + d;
                if (d < uint32_t(Cfg::span)) s_in[rr * P + d] = (gi >= 0 && gi < n_in) ? __ldg(x + gi) : int16_t(0);
            }
        }
        __syncwarp();

        // ---- lane = row: walk the row, one input step at a time ----
        {
            const int16_t* wp = s_in + lane * P;
            int16_t* so = s_out + lane * Lb;
            float xw[16];
#pragma unroll
            for (int t = 0; t < 15; ++t) xw[t] = float(wp[t]);
            xw[15] = 0.f;
            wp += 15;
#pragma unroll 1
            for (uint32_t s0 = 0; s0 < Mb; s0 += 16) {
                rs_step<Cfg, 0>(T, xw, wp, s0 + 0, so);
                rs_step<Cfg, 1>(T, xw, wp, s0 + 1, so);
                rs_step<Cfg, 2>(T, xw, wp, s0 + 2, so);
                rs_step<Cfg, 3>(T, xw, wp, s0 + 3, so);
                rs_step<Cfg, 4>(T, xw, wp, s0 + 4, so);
                rs_step<Cfg, 5>(T, xw, wp, s0 + 5, so);
                rs_step<Cfg, 6>(T, xw, wp, s0 + 6, so);
                rs_step<Cfg, 7>(T, xw, wp, s0 + 7, so);
                rs_step<Cfg, 8>(T, xw, wp, s0 + 8, so);
                rs_step<Cfg, 9>(T, xw, wp, s0 + 9, so);
                rs_step<Cfg, 10>(T, xw, wp, s0 + 10, so);
                rs_step<Cfg, 11>(T, xw, wp, s0 + 11, so);
                rs_step<Cfg, 12>(T, xw, wp, s0 + 12, so);
                rs_step<Cfg, 13>(T, xw, wp, s0 + 13, so);
                rs_step<Cfg, 14>(T, xw, wp, s0 + 14, so);
                rs_step<Cfg, 15>(T, xw, wp, s0 + 15, so);
            }
        }
        __syncwarp();

        // ---- the tile's outputs are one contiguous span of the stream: 128-bit stores when the clip starts on 16 bytes ----
        {
            const uint32_t n = uint32_t(j_hi - j_lo);
            int16_t* dst = y + j_lo;
            if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                const uint32_t body = n / 8;
                const uint4* sv = reinterpret_cast<const uint4*>(s_out);
                uint4* dv = reinterpret_cast<uint4*>(dst);
                for (uint32_t i0 = 0; i0 < body; i0 += 32) { const uint32_t i = i0 + lane; if (i < body) dv[i] = sv[i]; }
                for (uint32_t i0 = body * 8; i0 < n; i0 += 32) { const uint32_t i = i0 + lane; if (i < n) dst[i] = s_out[i]; }
            } else {
                for (uint32_t i0 = 0; i0 < n; i0 += 32) { const uint32_t i = i0 + lane; if (i < n) dst[i] = s_out[i]; }
            }
        }
        __syncwarp();
    }
}

// Host side: the taps of a rate in output order, cached per thread (one rate is used over and over).
struct RsPlan {
    uint32_t rate = 0;
    RsTables tables;
};

template <class Cfg>
static szb_status launch_rows(szb_ctx* ctx, RsPlan& plan, const int16_t* d_in, const uint64_t* d_in_off, const uint64_t* d_out_off,
                              uint32_t n_clips, uint64_t max_out, uint32_t rate, int16_t* d_out) {
    if (plan.rate != rate) {
        const auto c = resample_taps(rate);                        // [L][16]
        std::memset(&plan.tables, 0, sizeof plan.tables);
        for (uint32_t q = 0; q < uint32_t(Cfg::Lb); ++q) {
            const uint32_t ph = uint32_t((uint64_t(q) * Cfg::M) % Cfg::L);
            std::memcpy(&plan.tables.taps[size_t(q) * kResTaps], &c[size_t(ph) * kResTaps], kResTaps * sizeof(float));
        }
        plan.rate = rate;
    }
    RsArgs a;
    a.in = d_in;
    a.in_off = reinterpret_cast<const unsigned long long*>(d_in_off);
    a.out_off = reinterpret_cast<const unsigned long long*>(d_out_off);
    a.out = d_out;
    a.n_clips = n_clips;
    const uint64_t rows = (max_out + Cfg::Lb - 1) / Cfg::Lb;
    a.tiles_per_clip = uint32_t((rows + kRsRows - 1) / kRsRows);
    a.rate = rate;
    const uint64_t n_items = uint64_t(n_clips) * a.tiles_per_clip;
    const size_t smem = size_t(Cfg::warps) * Cfg::per_warp;
    SZB_CUDA(cudaFuncSetAttribute(resample_rows_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const uint64_t want = (n_items + Cfg::warps - 1) / Cfg::warps;
    const uint32_t grid = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>(want, uint64_t(ctx->sm_count))));
    resample_rows_kernel<Cfg><<<grid, Cfg::warps * 32, smem, ctx->stream>>>(plan.tables, a);
    SZB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return SZB_OK;
}

szb_status launch_resample(szb_ctx* ctx, const int16_t* d_in, const uint64_t* d_in_off, const uint64_t* d_out_off,
                           uint32_t n_clips, uint64_t max_out, uint32_t rate, int16_t* d_out) {
    if (n_clips == 0 || max_out == 0) return SZB_OK;
    static thread_local RsPlan plan;
#define SZB_RS_CASE(RATE, L, M, MULT) \
    case RATE: return launch_rows<RsCfg<L, M, MULT>>(ctx, plan, d_in, d_in_off, d_out_off, n_clips, max_out, rate, d_out)
    // MEASURED ON B200: with the 28 KB taps table streamed through the constant bank every row, the uniform loads miss the
    // small constant cache and the kernel is latency-bound (45 ms against 8.8 ms of the generic kernel on the C2 batch).
    // The row kernel stays compiled and tested (SZB_RESAMPLE_ROWS) but is not the default until the taps come from
    // shared memory.
    static const bool use_rows = [] { const char* e = getenv("SZB_RESAMPLE_ROWS"); return e && e[0] == '1'; }();
    if (use_rows) switch (rate) {      // L / M = 44100 / rate reduced; rows of MULT periods
        SZB_RS_CASE(8000, 441, 80, 1);
        SZB_RS_CASE(11025, 4, 1, 32);
        SZB_RS_CASE(12000, 147, 40, 2);
        SZB_RS_CASE(16000, 441, 160, 1);
        SZB_RS_CASE(22050, 2, 1, 64);
        SZB_RS_CASE(24000, 147, 80, 2);
        SZB_RS_CASE(32000, 441, 320, 1);
        SZB_RS_CASE(48000, 147, 160, 2);
        default: break;
    }
#undef SZB_RS_CASE
    return launch_resample_generic(ctx, d_in, d_in_off, d_out_off, n_clips, max_out, rate, d_out);
}

}  // namespace szb
