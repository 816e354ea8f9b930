// EXPERIMENT (not part of libstreamz_b200): integer polyphase resampler on the tensor cores (tcgen05.mma kind::i8, s32
// accumulators in TMEM).  Bit-exact against a scalar evaluation of the integer specification below on ragged, short, odd
// and full-scale clips (resample_tc_test.cu), but MEASURED at 10.5 ms for the C2 batch against 7.0 ms of the float
// CUDA-core kernel of the product: with one 256-thread CTA per SM (193 KB of shared memory) the epilogue and the 2-byte
// write-out are latency-bound.  A 512-thread, role-split version is projected at 5-6 ms -- not enough to justify
// replacing the float-FMA specification of resample_to_44100.  Kept as the starting point for that work.
//
// Specification (DESIGN.md "Resampler"; the same arithmetic as the CUDA-core kernel in frontend.cu and the oracle):
//   y[j] = clamp( ( sum_{t<16} cq[p][t] * x[i0 - 7 + t]  +  2^14 ) >> 15 ),   j M = i0 L + p,
// with 16-bit integer taps cq (the float taps x 2^15, rounded, each phase adjusted to sum to exactly 2^15).  Every product
// and sum is an integer, so the result does not depend on the order of accumulation: it can be computed EXACTLY by 8-bit
// tensor-core MMAs.  x = 256 xh + xl and cq = 256 ch + cl (xh, ch signed bytes; xl, cl unsigned bytes) give
//   sum cq x = 65536 sum(xh ch) + 256 (sum(xh cl) + sum(xl ch)) + sum(xl cl),
// four MMAs per K-step into three accumulators, combined in the epilogue as 2 hh + ((256 mid + ll + 2^14) >> 15).
//
// Mapping.  The ratio repeats with period (L outputs, M = 160 inputs).  MMA row = one period of a clip: its M + 32 input
// samples x[M r - 8 ..] are the K dimension (6 K-steps of 32), its L outputs the N dimension, cut into chunks of 32 whose
// 16-tap windows touch at most 3 K-steps.  The B operand is the banded Toeplitz block of the taps for (chunk, K-step),
// built once on the host; A is the byte-split PCM of 128 consecutive periods.  A dense 32 x 32 block carries ~16 % non-zero
// taps -- the tensor core does ~6x the necessary MACs and is still far from being the bottleneck (0.5 ms for 4.4 G outputs).
#pragma once
#include <cstdint>

#include "../../../streamz_b200/csrc/common.cuh"
#include "../../../streamz_b200/csrc/gemm_tc.cuh"

namespace szb {
namespace rtc {

constexpr int kM = 160;                  // inputs per period (all supported ratios are expressed with M = 160)
constexpr int kKSteps = kM / 32 + 1;     // 6 K-steps of 32 bytes: M + 32 samples per row
constexpr int kRows = 128;               // periods per tile = MMA M
constexpr int kNC = 32;                  // outputs per chunk = MMA N
constexpr int kMaxChunks = 14;           // L <= 448
constexpr int kThreads = 256;
constexpr int kAPlane = 2 * kRows * 128; // two 128-byte K panels per plane
constexpr int kBTile = kNC * 128;        // one (chunk, plane) Toeplitz block: 32 rows x (up to 4 K-step slots of 32 B)
constexpr int kStPitch = 17;             // staging row pitch in words (32 outputs = 16 words + 1): conflict-free for lane = row

struct Geom {
    int L, n_chunks;
    int ks_first[kMaxChunks], ks_count[kMaxChunks];   // K-steps [first, first + count) chunk n reads (count <= 3)
};

struct Args {
    const int16_t* in;
    const unsigned long long* in_off;
    const unsigned long long* out_off;
    const uint8_t* btiles;               // [n_chunks][2 planes (hi s8, lo u8)][32 rows][128 B], plain row-major
    int16_t* out;
    uint32_t n_clips, tiles_per_clip, rate;
};

constexpr size_t kSmemA = 2 * size_t(kAPlane);                          // 65 536
constexpr size_t smem_bytes(int n_chunks) { return kSmemA + size_t(n_chunks) * 2 * kBTile + 2 * size_t(kRows) * kStPitch * 4; }

__host__ __device__ constexpr uint32_t make_idesc_i8(int m, int n, int a_signed, int b_signed) {
    return (2u << 4) | (uint32_t(a_signed) << 7) | (uint32_t(b_signed) << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

// 8 consecutive samples of a clip starting at s (zeros outside [0, n_in)); 128-bit load when the clip allows it
__device__ __forceinline__ uint4 load8(const int16_t* x, int64_t n_in, int64_t s, bool vec) {
    if (s >= 0 && s + 8 <= n_in && vec) return __ldg(reinterpret_cast<const uint4*>(x + s));
    uint32_t w[4] = { 0u, 0u, 0u, 0u };
    if (s + 8 > 0 && s < n_in) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int64_t g = s + e;
            const uint32_t v = (g >= 0 && g < n_in) ? uint32_t(uint16_t(__ldg(x + g))) : 0u;
            w[e >> 1] |= v << (16 * (e & 1));
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// 8 samples -> 8 high bytes (signed plane) and 8 low bytes (unsigned plane), written to row r at K offset d (multiple of 8)
__device__ __forceinline__ void store_split(unsigned char* a_hi, unsigned char* a_lo, int r, int d, uint4 v) {
    const uint32_t off = uint32_t(d >> 7) * uint32_t(kRows * 128) + tc::sw128_off(r, (d & 127) >> 4) + uint32_t(d & 15);
    const uint2 hi = make_uint2(__byte_perm(v.x, v.y, 0x7531), __byte_perm(v.z, v.w, 0x7531));
    const uint2 lo = make_uint2(__byte_perm(v.x, v.y, 0x6420), __byte_perm(v.z, v.w, 0x6420));
    *reinterpret_cast<uint2*>(a_hi + off) = hi;
    *reinterpret_cast<uint2*>(a_lo + off) = lo;
}

__global__ void __launch_bounds__(kThreads, 1) resample_tc_kernel(const __grid_constant__ Geom g, const __grid_constant__ Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t s_full[2];
    __shared__ uint32_t s_tmem;
    unsigned char* a_hi = smem;
    unsigned char* a_lo = smem + kAPlane;
    unsigned char* b_base = smem + kSmemA;
    uint32_t* s_stage = reinterpret_cast<uint32_t*>(smem + kSmemA + size_t(g.n_chunks) * 2 * kBTile);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = g.L, NCH = g.n_chunks;

    // ---- once per CTA: TMEM, barriers, the Toeplitz blocks (swizzled on the way in) ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&s_tmem)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        tc::mbar_init(&s_full[0], 1);
        tc::mbar_init(&s_full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < NCH * 2 * kNC * 8; i += kThreads) {            // 16-byte chunks
        const int tile = i / (kNC * 8), rc = i - tile * (kNC * 8), r = rc >> 3, c = rc & 7;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.btiles + size_t(tile) * kBTile + r * 128 + c * 16));
        *reinterpret_cast<uint4*>(b_base + size_t(tile) * kBTile + tc::sw128_off(r, c)) = v;
    }
    for (int i = tid; i < int(kSmemA / 16); i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);   // K padding stays 0
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    uint32_t use[2] = { 0u, 0u };                                         // completed uses of each TMEM buffer (barrier parity)

    constexpr int kChunks8 = (kRows * kM + 32) / 8;                       // 16-byte input chunks per tile
    constexpr int kPre = (kChunks8 + kThreads - 1) / kThreads;
    const uint64_t n_items = uint64_t(a.n_clips) * a.tiles_per_clip;

    for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t clip = uint32_t(item / a.tiles_per_clip), tile = uint32_t(item - uint64_t(clip) * a.tiles_per_clip);
        const int64_t n_in = int64_t(a.in_off[clip + 1] - a.in_off[clip]);
        const uint64_t n_out = uint64_t(n_in) * 44100ull / a.rate;                       // lib.rs:196
        const uint64_t row0 = uint64_t(tile) * kRows;
        if (row0 * uint64_t(L) >= n_out) continue;                                       // uniform
        const int16_t* x = a.in + a.in_off[clip];
        int16_t* y = a.out + a.out_off[clip];
        const bool vec = (reinterpret_cast<uintptr_t>(x) & 15) == 0;

        // ---- A operand: byte-split PCM of the tile's 128 periods (row r: x[M (row0 + r) - 8 + d], d < M + 32) ----
        const int64_t s0 = int64_t(row0) * kM - 8;
        {
            uint4 v[kPre];                                                 // all loads in flight before the first use
#pragma unroll
            for (int i = 0; i < kPre; ++i) {
                const int ci = tid + i * kThreads;
                v[i] = ci < kChunks8 ? load8(x, n_in, s0 + 8 * int64_t(ci), vec) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int i = 0; i < kPre; ++i) {
                const int ci = tid + i * kThreads;
                if (ci < kChunks8) {
                    const int u = 8 * ci, r = u / kM, d = u - r * kM;
                    if (r < kRows) store_split(a_hi, a_lo, r, d, v[i]);
                    if (d < 32 && r >= 1) store_split(a_hi, a_lo, r - 1, d + kM, v[i]);  // the halo of the previous row
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();

        auto issue = [&](int n) {                                          // MMAs of chunk n into TMEM buffer n & 1
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d0 = tmem + uint32_t(n & 1) * 128u;
            const uint64_t dbh = tc::make_desc_k_sw128(tc::smem_u32(b_base + size_t(2 * n) * kBTile));
            const uint64_t dbl = tc::make_desc_k_sw128(tc::smem_u32(b_base + size_t(2 * n + 1) * kBTile));
            for (int j = 0; j < g.ks_count[n]; ++j) {
                const int ks = g.ks_first[n] + j;
                const uint32_t a_off = uint32_t(ks >> 2) * uint32_t(kRows * 128);
                const uint64_t dah = tc::make_desc_k_sw128(tc::smem_u32(a_hi + a_off)) + uint64_t(((ks & 3) * 32) >> 4);
                const uint64_t dal = tc::make_desc_k_sw128(tc::smem_u32(a_lo + a_off)) + uint64_t(((ks & 3) * 32) >> 4);
                const uint64_t bj = uint64_t((j * 32) >> 4);
                const uint32_t acc = j > 0;
                umma_i8(d0 + 0, dah, dbh + bj, make_idesc_i8(kRows, kNC, 1, 1), acc);    // xh ch
                umma_i8(d0 + 32, dah, dbl + bj, make_idesc_i8(kRows, kNC, 1, 0), acc);   // xh cl
                umma_i8(d0 + 32, dal, dbh + bj, make_idesc_i8(kRows, kNC, 0, 1), 1u);    // + xl ch
                umma_i8(d0 + 64, dal, dbl + bj, make_idesc_i8(kRows, kNC, 0, 0), acc);   // xl cl
            }
            tc::umma_commit(&s_full[n & 1]);
        };
        auto flush = [&](int n) {                                          // staged outputs of chunk n -> global
            const int q = kNC * n + lane;
            if (q >= L) return;
            const uint64_t j0 = row0 * uint64_t(L) + uint64_t(q);          // output index of row 0 of the tile
            if (j0 >= n_out) return;
            const int r_lim = int(min(uint64_t(kRows), (n_out - j0 + uint64_t(L) - 1) / uint64_t(L)));   // rows with j < n_out
            const uint32_t* st = s_stage + size_t(n & 1) * kRows * kStPitch + (lane >> 1);
            int16_t* yq = y + j0;
            const uint32_t sh = (lane & 1) * 16;
#pragma unroll 4
            for (int r = warp; r < r_lim; r += kThreads / 32) yq[size_t(r) * L] = int16_t((st[r * kStPitch] >> sh) & 0xFFFFu);
        };

        if (tid == 0) issue(0);
        for (int n = 0; n < NCH; ++n) {
            if (tid == 0 && n + 1 < NCH) issue(n + 1);
            if (n > 0) flush(n - 1);
            tc::mbar_wait(&s_full[n & 1], use[n & 1] & 1u);
            use[n & 1] += 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {   // epilogue: warp w reads TMEM lanes 32 (w % 4) .., columns 16 (w / 4) .. of the three accumulators
                const int row = 32 * (warp & 3) + lane, half = warp >> 2;
                const uint32_t base = tmem + (uint32_t(32 * (warp & 3)) << 16) + uint32_t(n & 1) * 128u + uint32_t(16 * half);
                uint32_t hh[16], mid[16], ll[16];
#define SZB_LD16(arr, addr)                                                                                                         \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
                 : "=r"(arr[0]), "=r"(arr[1]), "=r"(arr[2]), "=r"(arr[3]), "=r"(arr[4]), "=r"(arr[5]), "=r"(arr[6]), "=r"(arr[7]),    \
                   "=r"(arr[8]), "=r"(arr[9]), "=r"(arr[10]), "=r"(arr[11]), "=r"(arr[12]), "=r"(arr[13]), "=r"(arr[14]), "=r"(arr[15]) \
                 : "r"(addr))
                SZB_LD16(hh, base);
                SZB_LD16(mid, base + 32u);
                SZB_LD16(ll, base + 64u);
#undef SZB_LD16
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t* st = s_stage + size_t(n & 1) * kRows * kStPitch + row * kStPitch + 8 * half;
#pragma unroll
                for (int k = 0; k < 16; k += 2) {
                    int v0 = 2 * int(hh[k]) + ((int(mid[k]) * 256 + int(ll[k]) + 16384) >> 15);
                    int v1 = 2 * int(hh[k + 1]) + ((int(mid[k + 1]) * 256 + int(ll[k + 1]) + 16384) >> 15);
                    v0 = min(max(v0, -32768), 32767);
                    v1 = min(max(v1, -32768), 32767);
                    st[k >> 1] = (uint32_t(v0) & 0xFFFFu) | (uint32_t(v1) << 16);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();           // staging n complete; TMEM buffer n & 1 and (after the last chunk) the A tile are free
        }
        flush(NCH - 1);
        __syncthreads();               // staging and A are rewritten by the next tile
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
}

}  // namespace rtc
}  // namespace szb
