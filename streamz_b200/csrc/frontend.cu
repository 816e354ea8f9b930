// Front end of the StreamZ hot path as one fused sm_100a kernel:
//   i16 PCM @ 44.1 kHz -> 800-sample frames, hop 400 -> FFT -> |X|^2 (bins 0..400) -> 26 mel bands -> ln
//   -> DCT-II (20 coeffs) -> delta, delta-delta -> per-window z-score -> [n][60] f32
// Reference: FeatureExtractor::extract -> window_samples_with_plan, streamz-rs/src/lib.rs:261-345
// (no pre-emphasis, rectangular window, unnormalised FFT: lib.rs:292-301).
//
// Mapping (see DESIGN.md "Front-end kernel"): one persistent CTA per SM pulls (clip, window-range) segments from an
// atomic queue and walks each segment in tiles of 32 frames.  LANE = FRAME everywhere: the 32 lanes of a warp hold
// the same element of 32 consecutive frames, so every shared-memory row is [.. ][32 lanes] and every access is
// conflict-free by construction, every twiddle / mel / DCT coefficient is warp-uniform (constant bank), and each
// thread runs whole 20-point DFTs in registers.  Each PCM sample is read from HBM once per tile (+1 hop of overlap
// between tiles), each feature row is written once with 16-byte coalesced stores; MFCCs of the previous tile stay in
// a shared-memory ring so the +-2-frame delta stencil never recomputes or re-reads anything inside a segment.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "common.cuh"
#include "fft_math.cuh"
#include "frontend.cuh"
#include "tables.hpp"

namespace szb {

// ---- constant tables (uploaded per device at context creation) ------------------------------------------------------
__constant__ float2 c_tw400[kR * kR];     // W_400^{n1 k2}
__constant__ float2 c_tw800[201];         // W_800^k, k = 0..200
constexpr int kMelPad = 8;                // mel chunks are zero-padded to multiples of 8 bins (unrolled inner loop)
constexpr int kMelWCap = 48 * 40;         // capacity of the padded weight table
__constant__ float c_melw[kMelWCap];      // per-task mel weights, pre-scaled by 1 / (4 * 32767^2), zero-padded
__constant__ float c_dct[kMfcc * kMels];  // unscaled DCT-II rows
// The 26 filters hold 3..105 non-zero bins each; they are cut into chunks of <= 40 bins ("tasks") that are spread
// over the 20 warps by size (longest-processing-time first) so the mel stage ends at the same time on every warp.
struct MelTask { short k0, len, woff, slot; };   // len is a multiple of kMelPad
constexpr int kMaxMelTasks = 48;
__constant__ MelTask c_mel_tasks[kMaxMelTasks];      // grouped by warp
__constant__ int c_mel_warp_begin[20 + 1];           // tasks of warp w: [begin[w], begin[w+1])
__constant__ int c_mel_part_begin[kMels + 1];        // partial sums (slots) of filter m: [begin[m], begin[m+1])

constexpr int kTile = 32;                 // frames per tile == lanes per warp
constexpr int kWarps = 20;                // one warp per 20-point DFT column
constexpr int kThreads = kWarps * 32;
constexpr int kHopWords = kHop / 2;       // 200 packed (i16,i16) words per hop
constexpr int kHopStride = kHopWords + 1; // 201: odd lane stride -> conflict-free 32-bit reads with lane = frame
constexpr int kPcmRows = kTile + 1;       // 33 hops cover 32 frames
constexpr int kRing = 64;                 // MFCC ring slots (power of two >= kTile + 4)

constexpr size_t kSmemPcm = size_t(kPcmRows) * kHopStride * 4;          // 26 532
constexpr size_t kSmemS = size_t(kHalf) * kTile * 8;                    // 102 400
constexpr size_t kSmemP = size_t(kBins + kMelPad) * kTile * 4;          // 52 352: power rows + zero rows read by padded mel taps
constexpr size_t kSmemE = size_t(kMels + kMaxMelTasks) * kTile * 4;     // 9 472: ln energies + partial sums
constexpr int kRingStride = kRing + 1;                                  // 65: lane = coefficient reads stay conflict-free
constexpr size_t kSmemRing = size_t(kMfcc) * kRingStride * 4;           // 5 200
// The mel weights (7.7 KB, each read once per tile) do not fit the indexed-constant cache next to the twiddles and the
// DCT rows; they are copied into shared memory once per CTA (a warp-uniform LDS is a single-wavefront broadcast).
constexpr size_t kSmemMelW = size_t(kMelWCap) * 4;                      // 7 680
constexpr size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
constexpr size_t kOffPcm = 0;
constexpr size_t kOffS = align16(kOffPcm + kSmemPcm);
constexpr size_t kOffP = align16(kOffS + kSmemS);
constexpr size_t kOffE = align16(kOffP + kSmemP);
constexpr size_t kOffRing = align16(kOffE + kSmemE);
constexpr size_t kOffMelW = align16(kOffRing + kSmemRing);
constexpr size_t kSmemTotal = align16(kOffMelW + kSmemMelW);
static_assert(kSmemTotal <= 227 * 1024, "front-end tile does not fit in shared memory");

// (x[2n], x[2n+1]) packed in one word -> two floats, exactly, without the quarter-rate I2F pipe: flip the sign bits
// (x + 32768 as u16), splice each half under the exponent of 2^23 (byte permute) and subtract 2^23 + 32768.
// Measured: I2F held 9 % of the kernel's stall samples for 2.5 % of its instructions.
__device__ __forceinline__ void s16x2_to_f32(uint32_t v, float& lo, float& hi) {
    const uint32_t u = v ^ 0x80008000u;
    const uint32_t l = __byte_perm(u, 0x4B000000u, 0x7610), h = __byte_perm(u, 0x4B000000u, 0x7632);
    lo = __uint_as_float(l) - 8421376.f;
    hi = __uint_as_float(h) - 8421376.f;
}

// clamp to [-32768, 32767] and truncate toward zero (lib.rs:205-208) in ONE conversion: a float -> s16 cvt saturates at the
// ends of the destination range (two FMNMX fewer per output sample than clamping in float first; same bits)
__device__ __forceinline__ int16_t f32_to_s16_sat_rz(float v) {
    short r;
    asm("cvt.rzi.sat.s16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return r;
}

// Fused mode (kFusedD > 0): `pcm` holds the clips at their ORIGINAL rate and the polyphase FIR of resample_kernel runs inside the
// staging of every tile, so the 44.1 kHz signal exists only as the 33-hop tile in shared memory (same taps, same FMA order,
// same saturating conversion: the tile is bit-identical to what resample_kernel would have written to HBM).  Threads are
// dealt to output rows of Lb samples in teams of `wpt` warps exactly as resample_kernel deals them to a CTA (K = 3 adjacent
// outputs per thread, `lpw` lanes per warp so a window load is one shared-memory wavefront).
struct FusedParams {
    const float* taps;       // [L][16] polyphase taps of the rate
    uint32_t L, M;           // 44100 / rate reduced
    uint32_t Lb, G;          // outputs per row (multiple of lcm(L, 3)), column threads per row
    uint32_t in_per_row;     // Lb * M / L
    uint32_t lpw, wpt, n_teams;
};

template <bool kAligned16, int kFusedD = 0>   // aligned: every segment's first sample is 16-byte aligned (128-bit loads)
__global__ void __launch_bounds__(kThreads, 1)
extract_kernel(const int16_t* __restrict__ pcm, const Segment* __restrict__ segs, uint32_t n_segs,
               unsigned int* __restrict__ queue, float* __restrict__ out, const FusedParams fp = FusedParams{}) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* s_pcm = reinterpret_cast<uint32_t*>(smem + kOffPcm);
    cpx* s_S = reinterpret_cast<cpx*>(smem + kOffS);          // one 64-bit (re, im) pair per element
    float* s_P = reinterpret_cast<float*>(smem + kOffP);
    float* s_E = reinterpret_cast<float*>(smem + kOffE);
    float* s_ring = reinterpret_cast<float*>(smem + kOffRing);
    float* s_melw = reinterpret_cast<float*>(smem + kOffMelW);
    __shared__ uint32_t s_seg;

    // The warp index goes through a shuffle so that ptxas knows it is warp-uniform: twiddles, DCT rows and mel tasks
    // selected by it then travel through uniform loads / uniform registers instead of per-thread LDC, and the branches on it
    // need no divergence scaffolding (measured: LDC held 13 % of the stall samples, BSSY/BSYNC another 5 %).
    // `uwarp` is used ONLY to select constants and to branch, never in a per-thread address: kept apart from `warp`, it
    // stays in a uniform register.
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, uwarp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < kMelPad * kTile; i += kThreads) s_P[kBins * kTile + i] = 0.f;   // rows 401.. stay zero
    for (int i = tid; i < kMelWCap; i += kThreads) s_melw[i] = c_melw[i];
    __syncthreads();
    // fused mode: the thread's place in the resampler's row walk never changes; its three phases (< 4096 each), their window
    // shifts (2 bits each) and its window start are packed into two registers that stay live across the whole kernel
    uint32_t f_pk0 = 0, f_pk1 = 0;
    if (kFusedD > 0) {
        const uint32_t team = uint32_t(warp) / fp.wpt, wl = uint32_t(warp) - team * fp.wpt, g = wl * fp.lpw + uint32_t(lane);
        if (team < fp.n_teams && g < fp.G && (uint32_t(lane) < fp.lpw || wl == fp.wpt - 1)) {
            const uint32_t q0 = uint32_t((uint64_t(3 * g) * fp.M) / fp.L);
            uint32_t p[3], d[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint64_t pos = uint64_t(3 * g + k) * fp.M;
                d[k] = uint32_t(pos / fp.L) - q0;
                p[k] = uint32_t(pos % fp.L);
            }
            f_pk0 = p[0] | (p[1] << 12) | ((d[0] | (d[1] << 2) | (d[2] << 4)) << 24);
            f_pk1 = p[2] | (q0 << 12) | 0x80000000u;      // bit 31: this thread filters
        }
    }
    for (;;) {
        if (tid == 0) s_seg = atomicAdd(queue, 1u);
        __syncthreads();
        const uint32_t si = __shfl_sync(0xffffffffu, s_seg, 0);   // (shuffles: provably warp-uniform control flow below)
        __syncthreads();
        if (si >= n_segs) break;
        Segment sg = segs[si];
        sg.pcm_off = __shfl_sync(0xffffffffu, sg.pcm_off, 0);
        sg.out_row = __shfl_sync(0xffffffffu, sg.out_row, 0);
        sg.n_total = __shfl_sync(0xffffffffu, sg.n_total, 0);
        sg.w_begin = __shfl_sync(0xffffffffu, sg.w_begin, 0);
        sg.w_end = __shfl_sync(0xffffffffu, sg.w_end, 0);
        const uint32_t n_total = sg.n_total;
        const uint32_t f_lo = sg.w_begin >= 2 ? sg.w_begin - 2 : 0;
        const uint32_t f_hi = min(sg.w_end + 2, n_total);
        const int16_t* clip = pcm + sg.pcm_off;
        float* out_clip = out + sg.out_row * kFeat;
        uint32_t emit_next = sg.w_begin;

        // PCM tile = hops [a, a + nf] = 33 * 800 B = 1650 uint4; thread t owns uint4 t, t + 640, t + 1280.
        // They are fetched into registers one tile AHEAD (right after stage A has consumed the staging buffer), so the
        // HBM latency hides behind stages B..7 of the previous tile.
        constexpr int kVecPerHop = kHop / 8;                       // 50 uint4 per hop
        constexpr int kVecPerTile = kPcmRows * kVecPerHop;         // 1650
        constexpr int kPre = (kVecPerTile + kThreads - 1) / kThreads;  // 3
        uint4 pre[kFusedD == 0 ? kPre : 1];
        auto fetch_tile = [&](uint32_t ta) {
            const uint32_t tnf = min(uint32_t(kTile), f_hi - ta);
            const int16_t* src = clip + size_t(ta) * kHop;
#pragma unroll
            for (int r = 0; r < (kFusedD == 0 ? kPre : 0); ++r) {
                const int q = tid + r * kThreads;
                pre[r] = make_uint4(0u, 0u, 0u, 0u);
                if (q < int(tnf + 1) * kVecPerHop) {
                    if (kAligned16) {
                        pre[r] = __ldg(reinterpret_cast<const uint4*>(src) + q);
                    } else {
                        const int16_t* s8 = src + size_t(q) * 8;
                        uint32_t wv[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            wv[e] = uint32_t(uint16_t(__ldg(s8 + 2 * e))) | (uint32_t(uint16_t(__ldg(s8 + 2 * e + 1))) << 16);
                        pre[r] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                    }
                }
            }
        };
        // ---- fused mode: the tile's span of the ORIGINAL-rate clip is requested with cp.async one tile ahead (the staging buffer is
        //      its landing zone), expanded to float into the (idle) FFT buffer, and filtered into the staging buffer.
        constexpr int kFW = kResTaps + (kFusedD > 0 ? kFusedD : 1);     // window of one thread: 16 taps + the spread of its 3 outputs
        int16_t* s_raw = reinterpret_cast<int16_t*>(s_pcm);
        float* s_in = reinterpret_cast<float*>(s_S);
        const int64_t n_in = int64_t(sg.pad);                          // input samples of the clip (fused mode)
        auto tile_rows = [&](uint32_t ta, uint32_t& R_lo, uint32_t& n_rows) {   // output rows of Lb samples that cover the tile's hops
            const uint32_t tnf = min(uint32_t(kTile), f_hi - ta);
            const uint32_t J0 = ta * uint32_t(kHop), J1 = (ta + tnf + 1) * uint32_t(kHop);   // < 2^32: the host admits clips below 10 M windows
            R_lo = J0 / fp.Lb;                                                     // (32-bit: a 64-bit division costs ~30 instructions
            n_rows = (J1 - 1) / fp.Lb - R_lo + 1;                                  //  per thread and tile)
        };
        auto issue_raw = [&](uint32_t ta) {
            uint32_t R_lo, n_rows;
            tile_rows(ta, R_lo, n_rows);
            const int64_t i_lo = int64_t(R_lo) * fp.in_per_row - (kResTaps / 2 - 1), a0 = i_lo & ~int64_t(7);
            const uint32_t n_chunks = (uint32_t(i_lo - a0) + n_rows * fp.in_per_row + kFW + 7) / 8;
            for (uint32_t c = tid; c < n_chunks; c += kThreads) {
                const int64_t gs = a0 + 8 * int64_t(c);
                const bool in = gs >= 0 && gs < n_in;
                const uint32_t bytes = in ? uint32_t(min(int64_t(16), (n_in - gs) * 2)) : 0u;   // the copy zero-fills the rest
                const uint32_t dst = uint32_t(__cvta_generic_to_shared(s_raw + 8 * c));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(in ? clip + gs : clip), "r"(bytes) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (kFusedD == 0) fetch_tile(f_lo); else issue_raw(f_lo);

        for (uint32_t a = f_lo; a < f_hi; a += kTile) {
            const uint32_t nf = min(uint32_t(kTile), f_hi - a);
            if (kFusedD == 0) {
            // ---- 1. prefetched registers -> staging buffer (row stride 201 words keeps lane = frame conflict-free) ----
#pragma unroll
            for (int r = 0; r < (kFusedD == 0 ? kPre : 0); ++r) {
                const int q = tid + r * kThreads;
                if (q < kVecPerTile) {
                    const int h = q / kVecPerHop, c4 = q - h * kVecPerHop;
                    uint32_t* d = s_pcm + h * kHopStride + c4 * 4;
                    d[0] = pre[r].x; d[1] = pre[r].y; d[2] = pre[r].z; d[3] = pre[r].w;
                }
            }
            } else {
            // ---- 1f. raw span -> float -> polyphase FIR -> 44.1 kHz tile (resample_to_44100, lib.rs:186-209, in shared memory) ----
                uint32_t R_lo, n_rows;
                tile_rows(a, R_lo, n_rows);
                const int64_t i_lo = int64_t(R_lo) * fp.in_per_row - (kResTaps / 2 - 1), a0 = i_lo & ~int64_t(7);
                const uint32_t in_shift = uint32_t(i_lo - a0);
                const uint32_t n_chunks = (in_shift + n_rows * fp.in_per_row + kFW + 7) / 8;
                // this thread's taps: its three phases' rows, shifted by exact zeros to the common window start (registers are
                // reloaded per tile: they cannot stay live across the FFT stages)
                const uint32_t team = uint32_t(warp) / fp.wpt, g = (uint32_t(warp) - team * fp.wpt) * fp.lpw + uint32_t(lane);
                const uint32_t q0 = (f_pk1 >> 12) & 0x7FFFFu;
                const bool worker = (f_pk1 >> 31) != 0;
                const uint32_t f_p[3] = { f_pk0 & 0xFFFu, (f_pk0 >> 12) & 0xFFFu, f_pk1 & 0xFFFu };
                const uint32_t f_d[3] = { (f_pk0 >> 24) & 3u, (f_pk0 >> 26) & 3u, (f_pk0 >> 28) & 3u };
                float c[3][kFW];
                if (worker) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float4* cp4 = reinterpret_cast<const float4*>(fp.taps + size_t(f_p[k]) * kResTaps);
                        float row[kResTaps];
#pragma unroll
                        for (int v = 0; v < kResTaps / 4; ++v) {
                            const float4 t4 = __ldg(cp4 + v);
                            row[4 * v] = t4.x; row[4 * v + 1] = t4.y; row[4 * v + 2] = t4.z; row[4 * v + 3] = t4.w;
                        }
#pragma unroll
                        for (int t = 0; t < kFW; ++t) {
                            float v = 0.f;
#pragma unroll
                            for (int dd = 0; dd <= (kFusedD > 0 ? kFusedD : 1); ++dd)
                                if (t - dd >= 0 && t - dd < kResTaps && f_d[k] == uint32_t(dd)) v = row[t - dd];
                            c[k][t] = v;
                        }
                    }
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncthreads();
                for (uint32_t cidx = tid; cidx < n_chunks; cidx += kThreads) {
                    const uint4 rawv = *reinterpret_cast<const uint4*>(s_raw + 8 * cidx);
                    float4 lo, hi;
                    s16x2_to_f32(rawv.x, lo.x, lo.y); s16x2_to_f32(rawv.y, lo.z, lo.w);
                    s16x2_to_f32(rawv.z, hi.x, hi.y); s16x2_to_f32(rawv.w, hi.z, hi.w);
                    reinterpret_cast<float4*>(s_in)[2 * cidx] = lo;
                    reinterpret_cast<float4*>(s_in)[2 * cidx + 1] = hi;
                }
                __syncthreads();                                   // float span complete; the raw span may be overwritten
                if (worker) {
                    const uint32_t J0 = a * uint32_t(kHop);
                    const int32_t span_out = int32_t(nf + 1) * kHop;
                    for (uint32_t r = team; r < n_rows; r += fp.n_teams) {
                        const float* wv = s_in + in_shift + r * fp.in_per_row + q0;
                        float w[kFW];
#pragma unroll
                        for (int t = 0; t < kFW; ++t) w[t] = wv[t];
                        float acc[3] = { 0.f, 0.f, 0.f };
#pragma unroll
                        for (int t = 0; t < kFW; ++t) {
#pragma unroll
                            for (int k = 0; k < 3; ++k) acc[k] = fmaf(c[k][t], w[t], acc[k]);
                        }
                        const int32_t jj0 = int32_t((R_lo + r) * fp.Lb + 3 * g) - int32_t(J0);    // position inside the tile's 33 hops
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const int32_t jj = jj0 + k;
                            if (jj >= 0 && jj < span_out) {
                                const int32_t h = (jj * 5243) >> 21;                           // jj / 400 for jj < 16384
                                reinterpret_cast<int16_t*>(s_pcm)[jj + 2 * h] = f32_to_s16_sat_rz(acc[k]);   // hop rows are 402 halfwords apart
                            }
                        }
                    }
                }
            }
            __syncthreads();

            // ---- 2. stage A: warp = n1, lane = frame.  20-point DFT over n2 of z[n1 + 20 n2], twiddle, transpose ----
            {
                const int n1 = warp;
                const uint32_t* pw = s_pcm + lane * kHopStride + n1;
                cpx z[kR];
#pragma unroll
                for (int n2 = 0; n2 < kR; ++n2) {
                    // z[n] lives in hop `lane` for n < 200 and in hop `lane + 1` (row stride 201 = 200 + 1) above
                    const uint32_t v = pw[kR * n2 + (n2 >= kR / 2 ? 1 : 0)];
                    float lo, hi;
                    s16x2_to_f32(v, lo, hi);
                    z[n2] = cpack(lo, hi);
                }
                dft20(z);
                cpx* dst = s_S + n1 * kTile + lane;
                dst[0] = z[0];
                const float2* tw = c_tw400 + uwarp * kR;
#pragma unroll
                for (int k2 = 1; k2 < kR; ++k2) {
                    const float2 w = tw[k2];
                    dst[k2 * kR * kTile] = cmul_tw(z[k2], w.x, w.y);
                }
            }
            __syncthreads();
            if (a + kTile < f_hi) {                         // staging buffer is free again: start the next tile's loads
                if (kFusedD == 0) fetch_tile(a + kTile); else issue_raw(a + kTile);
            }

            // ---- 3. stage B: warp = k2, lane = frame.  20-point DFT over n1, in place: row k2*20 + k1 = Z[k2 + 20 k1] --
            {
                cpx* col = s_S + warp * kR * kTile + lane;
                cpx z[kR];
#pragma unroll
                for (int n1 = 0; n1 < kR; ++n1) z[n1] = col[n1 * kTile];
                dft20(z);
#pragma unroll
                for (int k1 = 0; k1 < kR; ++k1) col[k1 * kTile] = z[k1];
            }
            __syncthreads();

            // ---- 4. real-input split + power.  warp = k2 handles bins k = k2 + 20 k1 (k <= 200) with their mirrors
            //         400 - k = (20 - k2) + 20 (19 - k1): all row indices are affine in the unrolled k1 -----------------
            {
                const int k2 = warp;
                if (uwarp == 0) {
                    const cpx z = s_S[lane];
                    const float p0 = cre(z) + cim(z), p1 = cre(z) - cim(z);
                    s_P[lane] = 4.f * p0 * p0;
                    s_P[400 * kTile + lane] = 4.f * p1 * p1;
#pragma unroll
                    for (int k1 = 1; k1 <= 10; ++k1) {             // k = 20 k1, mirror row 20 - k1 (k1 = 10: itself)
                        const float2 w = c_tw800[kR * k1];
                        float pk, pmk;
                        split_pair_power(s_S[k1 * kTile + lane], s_S[(kR - k1) * kTile + lane], w.x, w.y, pk, pmk);
                        s_P[(kR * k1) * kTile + lane] = pk;
                        s_P[(kHalf - kR * k1) * kTile + lane] = pmk;
                    }
                } else {
                    const cpx* ra = s_S + (k2 * kR) * kTile + lane;
                    const cpx* rb = s_S + ((kR - k2) * kR + kR - 1) * kTile + lane;
                    const float2* tw = c_tw800 + uwarp;
                    float* pa = s_P + k2 * kTile + lane;
                    float* pb = s_P + (kHalf - k2) * kTile + lane;
#pragma unroll
                    for (int k1 = 0; k1 < 10; ++k1) {
                        const float2 w = tw[kR * k1];
                        float pk, pmk;
                        split_pair_power(ra[k1 * kTile], rb[-k1 * kTile], w.x, w.y, pk, pmk);
                        pa[(kR * k1) * kTile] = pk;
                        pb[-(kR * k1) * kTile] = pmk;
                    }
                }
            }
            __syncthreads();

            // ---- 5a. mel partial sums: size-balanced chunks of the sparse 26 x 401 bank (lib.rs:303-308) ----------------
            for (int t = c_mel_warp_begin[uwarp]; t < c_mel_warp_begin[uwarp + 1]; ++t) {
                const MelTask mt = c_mel_tasks[t];
                const float* wv = s_melw + mt.woff;
                const float* pp = s_P + mt.k0 * kTile + lane;
                float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 1
                for (int i = 0; i < mt.len; i += kMelPad) {
#pragma unroll
                    for (int u = 0; u < kMelPad; u += 2) {
                        acc0 = fmaf(wv[i + u], pp[(i + u) * kTile], acc0);
                        acc1 = fmaf(wv[i + u + 1], pp[(i + u + 1) * kTile], acc1);
                    }
                }
                s_E[(kMels + mt.slot) * kTile + lane] = acc0 + acc1;
            }
            __syncthreads();
            // ---- 5b. ln(max(sum, 1e-12)) (lib.rs:309) --------------------------------------------------------------------
            for (int m = uwarp; m < kMels; m += kWarps) {
                float acc = 0.f;
                for (int q = c_mel_part_begin[m]; q < c_mel_part_begin[m + 1]; ++q) acc += s_E[(kMels + q) * kTile + lane];
                s_E[m * kTile + lane] = logf(fmaxf(acc, 1e-12f));
            }
            __syncthreads();

            // ---- 6. DCT-II, first 20 coefficients: warp = coefficient j, lane = frame (lib.rs:312-315) ----------------
            {
                const int j = warp;
                const float* dj = c_dct + uwarp * kMels;
                float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                for (int m = 0; m < kMels; m += 2) {
                    acc0 = fmaf(dj[m], s_E[m * kTile + lane], acc0);
                    acc1 = fmaf(dj[m + 1], s_E[(m + 1) * kTile + lane], acc1);
                }
                if (uint32_t(lane) < nf) s_ring[j * kRingStride + ((a + lane) & (kRing - 1))] = acc0 + acc1;
            }
            __syncthreads();

            // ---- 7. delta, delta-delta, z-score and store for the windows whose +-2 neighbours are now known.
            //         LANE = WINDOW, warp = coefficient j (as in stage 6): a thread forms (c, d, dd) of its window for one j
            //         from 5 ring words (consecutive lanes -> consecutive words), the 60-value moments cross the warps through
            //         shared memory (two warps add the 20 partial sums), and the normalised rows make one trip through a
            //         [window][60] tile so that the block leaves as contiguous 16-byte stores.  (Round 1 ran one warp per
            //         window with 20 of its 32 lanes active: 165 warp-instructions per window, 14 % of the kernel.)
            //         The power buffer is free from here to the next tile's stage 4 and provides the scratch.
            const bool last = a + nf >= f_hi;
            const uint32_t lim = last ? sg.w_end : min(sg.w_end, a + nf - 2);
            float* s_red = s_P;                               // [2][kMfcc][32] partial sums of x and x^2
            float* s_mom = s_P + 2 * kMfcc * kTile;           // [2][32] sum x, sum x^2 per window
            float* s_rows = s_P + 2 * kMfcc * kTile + 2 * kTile;   // [32][60] normalised rows (16-byte aligned)
            for (uint32_t wb = emit_next; wb < lim; wb += kTile) {
                const uint32_t w = wb + lane;
                const bool live = w < lim;
                const uint32_t w_hi = min(lim, wb + kTile);
                const bool interior = wb >= 2 && w_hi + 2 <= n_total;            // block-uniform: no index clamps needed
                float c0 = 0.f, d1 = 0.f, d2 = 0.f;
                {
                    const float* rj = s_ring + warp * kRingStride;
                    auto C = [rj](int g) { return rj[g & (kRing - 1)]; };
                    if (interior) {                                // same operations on the same values as below: same bits
                        const float m2 = C(int(w) - 2), m1 = C(int(w) - 1), p1 = C(int(w) + 1), p2 = C(int(w) + 2);
                        c0 = C(int(w));
                        d1 = (p1 - m1) * 0.5f;                                     // lib.rs:223
                        d2 = ((p2 - c0) * 0.5f - (c0 - m2) * 0.5f) * 0.5f;         // lib.rs:322
                    } else if (live) {
                        const int hi = int(n_total) - 1;
                        auto cl = [hi](int x) { return min(max(x, 0), hi); };
                        const int ip = cl(int(w) + 1), im = cl(int(w) - 1);
                        c0 = C(int(w));
                        d1 = (C(ip) - C(im)) * 0.5f;                               // lib.rs:223
                        const float dp = (C(cl(ip + 1)) - C(cl(ip - 1))) * 0.5f;   // delta at clamp(w+1)
                        const float dm = (C(cl(im + 1)) - C(cl(im - 1))) * 0.5f;   // delta at clamp(w-1)
                        d2 = (dp - dm) * 0.5f;                                     // lib.rs:322
                    }
                }
                s_red[warp * kTile + lane] = c0 + d1 + d2;
                s_red[(kMfcc + warp) * kTile + lane] = fmaf(c0, c0, fmaf(d1, d1, d2 * d2));
                __syncthreads();
                if (uwarp < 2) {                                   // warp 0: sum x, warp 1: sum x^2 (fixed order)
                    const float* r = s_red + uwarp * kMfcc * kTile + lane;
                    float acc0 = r[0], acc1 = r[kTile];
#pragma unroll
                    for (int j = 2; j < kMfcc; j += 2) { acc0 += r[j * kTile]; acc1 += r[(j + 1) * kTile]; }
                    s_mom[uwarp * kTile + lane] = acc0 + acc1;
                }
                __syncthreads();
                // mean and population variance over the 60 values (lib.rs:328-336).  var = E[x^2] - mean^2: c0 carries most
                // of E[x^2] and of the variance alike (the deltas average out near zero), so the subtraction loses no digits
                // that matter at the 1e-4 parity bar (measured against the float64 oracle: unchanged 1-3e-6);
                // 1 / max(sqrt(var), 1e-6) (lib.rs:337) as min(rsqrt(var), 1e6).
                const float mean = s_mom[lane] * (1.f / float(kFeat));
                const float var = fmaxf(fmaf(-mean, mean, s_mom[kTile + lane] * (1.f / float(kFeat))), 0.f);
                const float inv = fminf(rsqrtf(var), 1e6f);
                float* row = s_rows + lane * kFeat + warp;
                row[0] = (c0 - mean) * inv;                                        // lib.rs:338-340
                row[kMfcc] = (d1 - mean) * inv;
                row[2 * kMfcc] = (d2 - mean) * inv;
                __syncthreads();
                {   // rows wb .. w_hi - 1 are contiguous in the output: (w_hi - wb) * 15 sixteen-byte chunks
                    const int n_vec = int(w_hi - wb) * (kFeat / 4);
                    float4* dst = reinterpret_cast<float4*>(out_clip + size_t(wb) * kFeat);
                    const float4* src = reinterpret_cast<const float4*>(s_rows);
                    if (tid < n_vec) dst[tid] = src[tid];
                }
                // (the next pass of this loop, or the next tile's stage 4, writes the scratch only after further barriers)
                if (wb + kTile < lim) __syncthreads();
            }
            emit_next = max(emit_next, lim);
        }
    }
}

// ---- resampler (resample_to_44100, lib.rs:186-209; polyphase spec in DESIGN.md "Resampler") --------------------------
// Bit-exact contract: acc = fma(c[p][t], float(x[i0 - 7 + t]), acc) for t = 0..15 in float32 (inputs outside the clip
// are 0), then clamp to [-32768, 32767] and truncate toward zero (lib.rs:205-208).
//
// Mapping: outputs are walked in rows of Lb samples (Lb a multiple of lcm(L, 3)); a block works on tiles of kResRows
// rows.  The tile's input span is loaded once (coalesced), converted to float once and staged in shared memory; the
// tile's output is staged in shared memory as i16 and leaves with 16-byte coalesced stores -- every input sample is
// read from HBM once and every output sample written once.  A thread owns THREE adjacent output columns for the whole
// kernel, so their phases and taps never change and live in registers, and the input index advances by the exact
// integer Lb M / L per row -- no division or table lookup in the loop.  The three outputs read overlapping input
// windows: one window of 16 + D samples is fetched from shared memory, and each output applies its 16 taps shifted by
// its own offset d_k <= D inside that window (the padding taps are exact zeros, so the accumulation order and the
// result bits are unchanged): 5.7 shared loads and 17 FMAs per output instead of 16 + 16.
constexpr int kResRows = 28;      // rows per tile: 4 CTAs of ~52 KB per SM at 16 kHz

// Shared-memory layout of a resampler CTA: float input span (origin moved back by up to 7 samples so that it starts on a
// 16-byte boundary of the clip), i16 output tile, raw i16 input span of the NEXT tile (cp.async landing zone).
struct ResLayout { size_t n_in_tile, off_out, off_raw, total; };
__host__ __device__ inline ResLayout res_layout(uint32_t in_per_row, uint32_t Lb, int W) {
    ResLayout l;
    l.n_in_tile = size_t(kResRows) * in_per_row + W;
    l.off_out = ((l.n_in_tile + 16) * 4 + 15) & ~size_t(15);
    l.off_raw = (l.off_out + (size_t(kResRows) * Lb + 16) * 2 + 15) & ~size_t(15);
    l.total = l.off_raw + ((l.n_in_tile + 24) * 2 + 15) / 16 * 16;
    return l;
}

// cp.async of the raw i16 span [i_lo & ~7, ...) of a tile into s_raw, 16 bytes at a time; bytes outside the clip are
// zero-filled by the copy itself (src-size operand).
__device__ __forceinline__ void res_issue(const int16_t* x, int64_t n_in, int64_t i_lo, uint32_t n_in_tile, int16_t* s_raw) {
    const int64_t a0 = i_lo & ~int64_t(7);
    const uint32_t n_chunks = (uint32_t(i_lo - a0) + n_in_tile + 7) / 8;
    for (uint32_t cidx = threadIdx.x; cidx < n_chunks; cidx += blockDim.x) {
        const int64_t gs = a0 + 8 * int64_t(cidx);
        const bool in = gs >= 0 && gs < n_in;
        const uint32_t bytes = in ? uint32_t(min(int64_t(16), (n_in - gs) * 2)) : 0u;
        const uint32_t dst = uint32_t(__cvta_generic_to_shared(s_raw + 8 * cidx));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(in ? x + gs : x), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int K, int D>   // K adjacent outputs per thread, whose windows start at most D input samples apart
__global__ void __launch_bounds__(160, (K <= 3 ? 4 : (K <= 4 ? 3 : 2)))
resample_kernel(const int16_t* __restrict__ in, const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ out_off,
                uint32_t n_clips, const float* __restrict__ taps, uint32_t L, uint32_t M, uint32_t Lb, uint32_t G, uint32_t rate,
                uint32_t lpw, int16_t* __restrict__ out) {
    constexpr int W = kResTaps + D;
    extern __shared__ __align__(16) unsigned char rs_smem[];
    const uint32_t in_per_row = uint32_t(uint64_t(Lb) * M / L);   // exact: L divides Lb
    const ResLayout lay = res_layout(in_per_row, Lb, W);
    const uint32_t n_in_tile = uint32_t(lay.n_in_tile);           // staged input samples per tile
    float* s_in = reinterpret_cast<float*>(rs_smem);
    int16_t* s_out = reinterpret_cast<int16_t*>(rs_smem + lay.off_out);
    int16_t* s_raw = reinterpret_cast<int16_t*>(rs_smem + lay.off_raw);
    // Output columns are dealt to the warps `lpw` lanes at a time (the last warp takes what is left, up to 32): with
    // lpw * K outputs spanning at most 32 input samples, the window loads of a warp touch at most 32 consecutive words --
    // one shared-memory wavefront each.  At 32 lanes (35 words for 16 kHz) every load cost two, and the kernel sat at 82 %
    // of the shared-memory wavefront peak.
    const uint32_t tid = threadIdx.x, n_warps = blockDim.x >> 5;
    const uint32_t g = (tid >> 5) * lpw + (tid & 31u);
    const bool worker = g < G && ((tid & 31u) < lpw || (tid >> 5) == n_warps - 1);
    uint32_t q0 = 0;
    float c[K][W];
    if (worker) {
        q0 = uint32_t((uint64_t(K * g) * M) / L);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint64_t pos = uint64_t(K * g + k) * M;
            const uint32_t p = uint32_t(pos % L), d = uint32_t(pos / L) - q0;
            const float* cp = taps + size_t(p) * kResTaps;
#pragma unroll
            for (int t = 0; t < W; ++t) {
                const int src = t - int(d);
                c[k][t] = (src >= 0 && src < kResTaps) ? __ldg(cp + src) : 0.f;
            }
        }
    }
    for (uint32_t clip = blockIdx.y; clip < n_clips; clip += gridDim.y) {
        const int64_t n_in = int64_t(in_off[clip + 1] - in_off[clip]);
        const uint64_t n_out = uint64_t(n_in) * 44100ull / rate;   // lib.rs:196 (out_off may be padded past this)
        const int16_t* x = in + in_off[clip];
        int16_t* y = out + out_off[clip];
        const uint64_t n_rows = (n_out + Lb - 1) / Lb;
        const bool vec_in = (reinterpret_cast<uintptr_t>(x) & 15) == 0;    // cp.async moves 16-byte chunks
        for (uint64_t row0 = uint64_t(blockIdx.x) * kResRows; row0 < n_rows; row0 += uint64_t(gridDim.x) * kResRows) {
            // ---- stage the tile's input span as float (zeros outside the clip) ----
            const int64_t i_lo = int64_t(row0) * in_per_row - (kResTaps / 2 - 1);
            uint32_t in_shift = 0;                                 // s_in[in_shift + i] holds x[i_lo + i]
            if (vec_in) {
                // The raw span was requested with cp.async while the previous tile was computing (2-byte global loads waited
                // on in the staging loop held 25 % of the stall samples): wait for it, expand it to float, and request the
                // next tile's span before the FMA loop starts.
                const int64_t a0 = i_lo & ~int64_t(7);            // chunk grid anchored at the (16-byte aligned) clip start
                in_shift = uint32_t(i_lo - a0);
                const uint32_t n_chunks = (in_shift + n_in_tile + 7) / 8;
                if (row0 == uint64_t(blockIdx.x) * kResRows) res_issue(x, n_in, i_lo, n_in_tile, s_raw);   // first tile of the clip
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncthreads();
                for (uint32_t cidx = tid; cidx < n_chunks; cidx += blockDim.x) {
                    const uint4 raw = *reinterpret_cast<const uint4*>(s_raw + 8 * cidx);
                    float4 lo, hi;
                    s16x2_to_f32(raw.x, lo.x, lo.y); s16x2_to_f32(raw.y, lo.z, lo.w);
                    s16x2_to_f32(raw.z, hi.x, hi.y); s16x2_to_f32(raw.w, hi.z, hi.w);
                    reinterpret_cast<float4*>(s_in)[2 * cidx] = lo;
                    reinterpret_cast<float4*>(s_in)[2 * cidx + 1] = hi;
                }
                __syncthreads();                                   // s_raw is free again, s_in complete
                const uint64_t row_next = row0 + uint64_t(gridDim.x) * kResRows;
                if (row_next < n_rows) res_issue(x, n_in, int64_t(row_next) * in_per_row - (kResTaps / 2 - 1), n_in_tile, s_raw);
            } else {
                for (uint32_t base = tid; base < n_in_tile; base += 8 * blockDim.x) {
                    int16_t v[8];                                  // 8 independent loads in flight per thread
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t i = base + u * blockDim.x;
                        const int64_t gi = i_lo + i;
                        v[u] = (i < n_in_tile && gi >= 0 && gi < n_in) ? __ldg(x + gi) : int16_t(0);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t i = base + u * blockDim.x;
                        if (i < n_in_tile) s_in[i] = float(v[u]);
                    }
                }
            }
            // outputs of the tile: [j_lo, j_hi); staged at s_out[shift + (j - j_lo)] so that global and shared addresses
            // are congruent modulo 16 bytes
            const uint64_t j_lo = row0 * Lb, j_hi = min(n_out, j_lo + uint64_t(kResRows) * Lb);
            const uint32_t shift = uint32_t((reinterpret_cast<uintptr_t>(y + j_lo) >> 1) & 7);
            if (!vec_in) __syncthreads();
            if (worker) {
                const uint32_t rows_here = uint32_t((j_hi - j_lo + Lb - 1) / Lb);
                for (uint32_t r = 0; r < rows_here; ++r) {
                    const float* wv = s_in + in_shift + r * in_per_row + q0;
                    float w[W];
#pragma unroll
                    for (int t = 0; t < W; ++t) w[t] = wv[t];
                    float acc[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) acc[k] = 0.f;
#pragma unroll
                    for (int t = 0; t < W; ++t) {
#pragma unroll
                        for (int k = 0; k < K; ++k) acc[k] = fmaf(c[k][t], w[t], acc[k]);
                    }
                    int16_t* so = s_out + shift + r * Lb + K * g;
#pragma unroll
                    for (int k = 0; k < K; ++k) so[k] = f32_to_s16_sat_rz(acc[k]);
                }
            }
            __syncthreads();
            // ---- coalesced write-out: scalar head up to the first 16-byte boundary, uint4 body, scalar tail ----
            {
                const uint32_t n = uint32_t(j_hi - j_lo);
                int16_t* dst = y + j_lo;
                const uint32_t head = min(n, (8u - shift) & 7u);
                const uint32_t body = (n - head) / 8;
                if (tid < head) dst[tid] = s_out[shift + tid];
                const uint4* sv = reinterpret_cast<const uint4*>(s_out + shift + head);
                uint4* dv = reinterpret_cast<uint4*>(dst + head);
                for (uint32_t i = tid; i < body; i += blockDim.x) dv[i] = sv[i];
                const uint32_t tail0 = head + body * 8;
                if (tid < n - tail0) dst[tail0 + tid] = s_out[shift + tail0 + tid];
            }
            __syncthreads();
        }
    }
}

// augment (lib.rs:103-116): circular shift, gain, uniform noise, clamp, truncating cast.  The three clip-level draws
// (noise_level, gain, shift) are made on the host; the per-sample noise comes from the counter RNG below so the kernel
// is reproducible and the oracle can restate it.  Arithmetic is done with explicitly rounded mul/add (no FMA
// contraction), as the Rust expression `samples[idx] as f32 * gain + noise * i16::MAX as f32` rounds.
__device__ __forceinline__ unsigned long long aug_splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void augment_kernel(const int16_t* __restrict__ in, uint64_t n, uint64_t shift, float gain, float noise_level,
                               unsigned long long key, int16_t* __restrict__ out) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        uint64_t idx = i + shift;
        if (idx >= n) idx -= n;                                           // (i + shift) % len, shift < len
        const float u = float(uint32_t(aug_splitmix64(key ^ i) >> 40)) * 5.9604644775390625e-08f;   // [0, 1)
        const float noise = __fmul_rn(__fsub_rn(__fmul_rn(2.f, u), 1.f), noise_level);               // U(-nl, nl)
        const float val = __fadd_rn(__fmul_rn(float(in[idx]), gain), __fmul_rn(noise, 32767.f));     // lib.rs:112
        out[i] = int16_t(__float2int_rz(fminf(fmaxf(val, -32768.f), 32767.f)));                      // lib.rs:113
    }
}

// downmix_to_mono (lib.rs:172-183)
__global__ void downmix_kernel(const int16_t* __restrict__ in, uint64_t n_in, uint32_t ch, int16_t* __restrict__ out,
                               uint64_t n_out) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_out; i += uint64_t(gridDim.x) * blockDim.x) {
        int32_t sum = 0;
        for (uint32_t c = 0; c < ch; ++c) {
            const uint64_t s = i * ch + c;
            if (s < n_in) sum += int32_t(in[s]);
        }
        out[i] = int16_t(sum / int32_t(ch));  // C++ integer division truncates toward zero like Rust
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------

szb_status upload_frontend_tables() {
    const auto tw400 = twiddles_400();
    const auto tw800 = twiddles_800_half();
    const auto dense = mel_filterbank_dense();
    const MelCsr csr = mel_filterbank_csr(dense);
    const auto dct = dct2_rows();
    const double scale = 1.0 / (4.0 * 32767.0 * 32767.0);  // undo integer-valued input and the unscaled real split
    SZB_CUDA(cudaMemcpyToSymbol(c_tw400, tw400.data(), tw400.size() * sizeof(float)));
    SZB_CUDA(cudaMemcpyToSymbol(c_tw800, tw800.data(), tw800.size() * sizeof(float)));
    SZB_CUDA(cudaMemcpyToSymbol(c_dct, dct.data(), dct.size() * sizeof(float)));
    {   // mel tasks: chunks of <= 40 bins, longest-processing-time-first over the warps
        struct T { MelTask t; int m; };
        std::vector<T> tasks;
        std::vector<float> melw;
        int part_begin[kMels + 1];
        int slot = 0;
        for (int m = 0; m < kMels; ++m) {
            part_begin[m] = slot;
            const int len = csr.len[m], parts = std::max(1, (len + 39) / 40);
            for (int p = 0; p < parts; ++p) {
                const int b = len * p / parts, e = len * (p + 1) / parts;
                const int padded = (e - b + kMelPad - 1) / kMelPad * kMelPad;   // extra taps are exact zeros
                SZB_REQUIRE(melw.size() + size_t(padded) <= size_t(kMelWCap), "mel weight table overflow");
                MelTask t;
                t.k0 = short(csr.start[m] + b); t.len = short(padded); t.woff = short(melw.size()); t.slot = short(slot++);
                for (int i = 0; i < padded; ++i)
                    melw.push_back(i < e - b ? float(double(csr.w[size_t(csr.off[m]) + b + i]) * scale) : 0.f);
                SZB_REQUIRE(t.k0 + padded <= kBins + kMelPad, "padded mel chunk runs past the zero rows");
                tasks.push_back({ t, m });
            }
        }
        part_begin[kMels] = slot;
        SZB_REQUIRE(slot <= kMaxMelTasks, "mel task table overflow (%d)", slot);
        std::sort(tasks.begin(), tasks.end(), [](const T& x, const T& y) { return x.t.len > y.t.len; });
        std::vector<std::vector<MelTask>> per_warp(kWarps);
        std::vector<int> load(kWarps, 0);
        for (const T& t : tasks) {
            const int w = int(std::min_element(load.begin(), load.end()) - load.begin());
            per_warp[w].push_back(t.t);
            load[w] += t.t.len + 6;   // + fixed per-task overhead
        }
        MelTask flat[kMaxMelTasks] = {};
        int warp_begin[kWarps + 1];
        int n = 0;
        for (int w = 0; w < kWarps; ++w) {
            warp_begin[w] = n;
            for (const MelTask& t : per_warp[w]) flat[n++] = t;
        }
        warp_begin[kWarps] = n;
        melw.resize(kMelWCap, 0.f);
        SZB_CUDA(cudaMemcpyToSymbol(c_melw, melw.data(), melw.size() * sizeof(float)));
        SZB_CUDA(cudaMemcpyToSymbol(c_mel_tasks, flat, sizeof flat));
        SZB_CUDA(cudaMemcpyToSymbol(c_mel_warp_begin, warp_begin, sizeof warp_begin));
        SZB_CUDA(cudaMemcpyToSymbol(c_mel_part_begin, part_begin, sizeof part_begin));
    }
    SZB_CUDA(cudaFuncSetAttribute(extract_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemTotal)));
    SZB_CUDA(cudaFuncSetAttribute(extract_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemTotal)));
    SZB_CUDA(cudaFuncSetAttribute(extract_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemTotal)));
    SZB_CUDA(cudaFuncSetAttribute(extract_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemTotal)));
    return SZB_OK;
}

// Split clips into (clip, window-range) segments: whole clips when there is enough work to fill the machine, else
// ranges of >= 64 windows so short batches still spread over the SMs.
void build_segments(const uint64_t* clip_off44, const uint64_t* win_off, uint32_t clip_begin, uint32_t clip_end, int sm_count,
                    std::vector<Segment>& segs, const uint64_t* clip_n_in) {
    const uint64_t total = win_off[clip_end] - win_off[clip_begin];
    uint64_t target = total / (uint64_t(sm_count) * 4);
    target = std::max<uint64_t>(64, std::min<uint64_t>(target, 4096));
    target = (target + kTile - 1) / kTile * kTile;
    for (uint32_t c = clip_begin; c < clip_end; ++c) {
        const uint64_t n = win_off[c + 1] - win_off[c];
        if (n == 0) continue;
        const uint64_t parts = (n + target + target / 2 - 1) / (target + target / 2);  // allow 1.5x target per segment
        const uint64_t np = std::max<uint64_t>(1, parts);
        for (uint64_t p = 0; p < np; ++p) {
            Segment s;
            s.pcm_off = clip_off44[c];
            s.out_row = win_off[c];
            s.n_total = uint32_t(n);
            s.w_begin = uint32_t(n * p / np);
            s.w_end = uint32_t(n * (p + 1) / np);
            s.pad = clip_n_in ? uint32_t(clip_n_in[c]) : 0u;
            if (s.w_end > s.w_begin) segs.push_back(s);
        }
    }
}

// Uploads a segment table (and zeroes `n_queues` work-queue counters) for one or more extract launches.
szb_status upload_segments(szb_ctx* ctx, const std::vector<Segment>& segs, uint32_t n_queues) {
    SZB_TRY(ctx->segs.reserve(std::max<size_t>(1, segs.size()) * sizeof(Segment)));
    SZB_TRY(ctx->counter.reserve(std::max<uint32_t>(1, n_queues) * sizeof(unsigned int)));
    if (!segs.empty()) {
        // pinned staging slot of its own per call: the copy is asynchronous and the *_dev entry points do not synchronise
        void* hp = nullptr;
        SZB_TRY(ctx->h_stage.acquire(segs.size() * sizeof(Segment), &hp));
        std::memcpy(hp, segs.data(), segs.size() * sizeof(Segment));
        SZB_CUDA(cudaMemcpyAsync(ctx->segs.ptr, hp, segs.size() * sizeof(Segment), cudaMemcpyHostToDevice, ctx->stream));
        SZB_TRY(ctx->h_stage.uploaded(ctx->stream));
    }
    SZB_CUDA(cudaMemsetAsync(ctx->counter.ptr, 0, std::max<uint32_t>(1, n_queues) * sizeof(unsigned int), ctx->stream));
    return SZB_OK;
}

// Geometry of the resampler's row walk for `rate` (shared by resample_kernel's launch and the fused extraction kernel).
struct ResGeom { uint32_t L, M, Lb, G, D, threads, lpw, in_per_row; };
static bool res_geometry(uint32_t rate, ResGeom& q) {
    constexpr uint32_t K = 3;
    resample_ratio(rate, q.L, q.M);
    const uint32_t span = uint32_t((uint64_t(K - 1) * q.M + q.L - 1) / q.L);   // K adjacent outputs span at most this many inputs
    if (span > 5 || q.L > 4096) return false;
    const uint32_t Lp = q.L / std::gcd(q.L, K) * K;                            // lcm(L, K)
    const uint32_t mult = std::max<uint32_t>(1, K * 128 / Lp);
    q.Lb = Lp * mult;
    q.G = q.Lb / K;
    if (q.G > 160) return false;
    q.threads = (q.G + 31) / 32 * 32;
    q.D = std::max<uint32_t>(1, span);
    q.in_per_row = uint32_t(uint64_t(q.Lb) * q.M / q.L);
    // lanes per warp: the most whose K * lpw outputs read at most 32 consecutive input samples, if the row still fits
    q.lpw = uint32_t((32ull * q.L) / (uint64_t(K) * q.M));
    const uint32_t n_warps = q.threads / 32;
    if (q.lpw >= 32 || q.lpw < 24 || (n_warps - 1) * q.lpw + 32 < q.G) q.lpw = 32;
    return true;
}

// Can szb_extract_batch(rate != 44100) run the FIR inside the extraction kernel's staging?  (window spread of a thread's three
// outputs <= 2 samples, i.e. rates up to 44.1 kHz; the tile's raw span fits the staging buffer and its float copy the FFT buffer)
bool fused_resample_supported(uint32_t rate) {
    ResGeom q;
    if (rate == SZB_SAMPLE_RATE || !res_geometry(rate, q) || q.D > 2) return false;
    const uint64_t rows = (uint64_t(kPcmRows) * kHop + q.Lb - 1) / q.Lb + 1;   // output rows a 33-hop tile can touch
    const uint64_t span = rows * q.in_per_row + kResTaps + q.D + 8;            // input samples staged per tile
    return span * 2 + 32 <= kSmemPcm && span * 4 + 64 <= kSmemS && kWarps / (q.threads / 32) >= 1;
}

// Launches the extraction kernel over segments [seg_begin, seg_begin + n_segs) of the uploaded table, queue `queue`.
// fused_rate != 0: d_pcm44 holds the clips at `fused_rate` Hz and the kernel resamples each tile itself (segments carry the
// clips' input offsets and lengths).
szb_status launch_extract(szb_ctx* ctx, const int16_t* d_pcm44, size_t seg_begin, size_t n_segs, uint32_t queue, float* d_feats,
                          bool aligned16, uint32_t fused_rate) {
    if (n_segs == 0) return SZB_OK;
    const int grid = int(std::min<size_t>(n_segs, size_t(ctx->sm_count)));
    FusedParams fp{};
    ResGeom q{};
    if (fused_rate) {
        SZB_REQUIRE(fused_resample_supported(fused_rate) && res_geometry(fused_rate, q) && aligned16, "fused resampling does not support rate %u",
                    fused_rate);
        if (ctx->taps_rate != fused_rate) {
            const auto taps = resample_taps(fused_rate);
            SZB_TRY(ctx->taps.reserve(taps.size() * sizeof(float)));
            SZB_CUDA(cudaMemcpyAsync(ctx->taps.ptr, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
            SZB_CUDA(cudaStreamSynchronize(ctx->stream));  // taps is a temporary
            ctx->taps_rate = fused_rate;
        }
        fp.taps = ctx->taps.as<float>();
        fp.L = q.L; fp.M = q.M; fp.Lb = q.Lb; fp.G = q.G; fp.in_per_row = q.in_per_row; fp.lpw = q.lpw;
        fp.wpt = q.threads / 32;
        fp.n_teams = uint32_t(kWarps) / fp.wpt;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->ktime_on) {
        SZB_CUDA(cudaEventCreate(&e0));
        SZB_CUDA(cudaEventCreate(&e1));
        SZB_CUDA(cudaEventRecord(e0, ctx->stream));
    }
    const Segment* sp = ctx->segs.as<Segment>() + seg_begin;
    unsigned int* qp = ctx->counter.as<unsigned int>() + queue;
    if (fused_rate && q.D == 1)
        extract_kernel<true, 1><<<grid, kThreads, kSmemTotal, ctx->stream>>>(d_pcm44, sp, uint32_t(n_segs), qp, d_feats, fp);
    else if (fused_rate)
        extract_kernel<true, 2><<<grid, kThreads, kSmemTotal, ctx->stream>>>(d_pcm44, sp, uint32_t(n_segs), qp, d_feats, fp);
    else if (aligned16)
        extract_kernel<true><<<grid, kThreads, kSmemTotal, ctx->stream>>>(d_pcm44, sp, uint32_t(n_segs), qp, d_feats);
    else
        extract_kernel<false><<<grid, kThreads, kSmemTotal, ctx->stream>>>(d_pcm44, sp, uint32_t(n_segs), qp, d_feats);
    SZB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    if (ctx->ktime_on) {
        SZB_CUDA(cudaEventRecord(e1, ctx->stream));
        ctx->ktime_pending.emplace_back(e0, e1);
    }
    return SZB_OK;
}

template <int K, int D>
static szb_status launch_resample_t(szb_ctx* ctx, dim3 grid, uint32_t threads, size_t smem, const int16_t* d_in, const uint64_t* d_in_off,
                                    const uint64_t* d_out_off, uint32_t n_clips, uint32_t L, uint32_t M, uint32_t Lb, uint32_t G,
                                    uint32_t rate, int16_t* d_out) {
    SZB_CUDA(cudaFuncSetAttribute(resample_kernel<K, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    // lanes per warp: the most whose K * lpw outputs read at most 32 consecutive input samples, if the row still fits
    uint32_t lpw = uint32_t((32ull * L) / (uint64_t(K) * M));
    const uint32_t n_warps = threads / 32;
    if (lpw >= 32 || lpw < 24 || (n_warps - 1) * lpw + 32 < G) lpw = 32;
    resample_kernel<K, D><<<grid, threads, smem, ctx->stream>>>(d_in, d_in_off, d_out_off, n_clips, ctx->taps.as<float>(), L, M, Lb, G,
                                                                 rate, lpw, d_out);
    SZB_CUDA(cudaGetLastError());
    return SZB_OK;
}

template <int K>
static szb_status launch_resample_k(szb_ctx* ctx, const int16_t* d_in, const uint64_t* d_in_off, const uint64_t* d_out_off,
                                    uint32_t n_clips, uint64_t max_out, uint32_t rate, int16_t* d_out) {
    static_assert(K == 3, "res_geometry is written for three outputs per thread");
    ResGeom q;
    SZB_REQUIRE(res_geometry(rate, q), "resample: unsupported rate %u", rate);
    const uint32_t L = q.L, M = q.M, Lb = q.Lb, G = q.G, threads = q.threads, D = q.D;
    if (ctx->taps_rate != rate) {
        const auto taps = resample_taps(rate);
        SZB_TRY(ctx->taps.reserve(taps.size() * sizeof(float)));
        SZB_CUDA(cudaMemcpyAsync(ctx->taps.ptr, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        SZB_CUDA(cudaStreamSynchronize(ctx->stream));  // taps is a temporary
        ctx->taps_rate = rate;
    }
    const size_t smem = res_layout(q.in_per_row, Lb, int(kResTaps + D)).total;
    SZB_REQUIRE(smem <= 200 * 1024, "resample: rate %u needs %zu bytes of shared memory", rate, smem);
    const uint64_t tiles = ((max_out + Lb - 1) / Lb + kResRows - 1) / kResRows;
    const uint32_t gy = std::min<uint32_t>(n_clips, 65535);
    const uint64_t want_x = std::max<uint64_t>(1, (uint64_t(ctx->sm_count) * 16 + gy - 1) / gy);
    const uint32_t gx = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>(want_x, tiles)));
    dim3 grid(gx, gy);
    if (D == 1) SZB_TRY((launch_resample_t<K, 1>(ctx, grid, threads, smem, d_in, d_in_off, d_out_off, n_clips, L, M, Lb, G, rate, d_out)));
    else if (D == 2) SZB_TRY((launch_resample_t<K, 2>(ctx, grid, threads, smem, d_in, d_in_off, d_out_off, n_clips, L, M, Lb, G, rate, d_out)));
    else if (D == 4) SZB_TRY((launch_resample_t<K, 4>(ctx, grid, threads, smem, d_in, d_in_off, d_out_off, n_clips, L, M, Lb, G, rate, d_out)));
    else if (D == 3) SZB_TRY((launch_resample_t<K, 3>(ctx, grid, threads, smem, d_in, d_in_off, d_out_off, n_clips, L, M, Lb, G, rate, d_out)));
    else SZB_TRY((launch_resample_t<K, 5>(ctx, grid, threads, smem, d_in, d_in_off, d_out_off, n_clips, L, M, Lb, G, rate, d_out)));
    ctx->launches += 1;
    return SZB_OK;
}

// MEASURED ALTERNATIVES (git history, DESIGN.md "Resampler"), none faster than this kernel's 8.8 ms on C2: K = 6 outputs
// per thread (12.1 ms: 168 registers halve the occupancy); the window fetched as aligned 128-bit vectors with the taps
// shifted to match (9.1 ms with spills at 96 registers, 10.1 ms at 128 registers and 3 CTAs per SM); a lane = row kernel with
// warp-uniform taps (constant bank: 45 ms, the 28 KB table thrashes the constant cache; shared memory: 9.2 ms, 16 operand
// words per output through the LSU).
szb_status launch_resample(szb_ctx* ctx, const int16_t* d_in, const uint64_t* d_in_off, const uint64_t* d_out_off,
                           uint32_t n_clips, uint64_t max_out, uint32_t rate, int16_t* d_out) {
    if (n_clips == 0 || max_out == 0) return SZB_OK;
    return launch_resample_k<3>(ctx, d_in, d_in_off, d_out_off, n_clips, max_out, rate, d_out);
}

szb_status launch_augment(szb_ctx* ctx, const int16_t* d_in, uint64_t n, uint64_t shift, float gain, float noise_level, uint64_t key,
                          int16_t* d_out) {
    if (n == 0) return SZB_OK;
    const int threads = 256;
    const uint64_t blocks = std::min<uint64_t>((n + threads - 1) / threads, uint64_t(ctx->sm_count) * 16);
    augment_kernel<<<uint32_t(blocks), threads, 0, ctx->stream>>>(d_in, n, shift, gain, noise_level, key, d_out);
    SZB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return SZB_OK;
}

szb_status launch_downmix(szb_ctx* ctx, const int16_t* d_in, uint64_t n_in, uint32_t ch, int16_t* d_out, uint64_t n_out) {
    if (n_out == 0) return SZB_OK;
    const int threads = 256;
    const uint64_t blocks = std::min<uint64_t>((n_out + threads - 1) / threads, uint64_t(ctx->sm_count) * 16);
    downmix_kernel<<<uint32_t(blocks), threads, 0, ctx->stream>>>(d_in, n_in, ch, d_out, n_out);
    SZB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return SZB_OK;
}

}  // namespace szb
