// PROTOTYPE (outside the library): the 800-point DFT of the front end on the 5th-generation tensor cores.
//
// Question: is a tcgen05 DFT faster than the CUDA-core FFT stages of extract_kernel
// (stages 2-4: ~640 of its 940 warp-instructions per window, ~10 ms of its 14.6 ms on BASELINE configs[1])?
//
// Formulation: 800 = 25 x 32, two GEMM stages with SHARED small operand matrices and w-major rows, fp16 operands split
// hi + lo (three products per stage: hi*hi + lo*hi + hi*lo, FP32 accumulation in TMEM):
//   stage 1   rows (w, n1), n1 < 25;  K = n2 < 32;  Y[w,n1,k2] = sum_n2 x[400 w + n1 + 25 n2] W32^(n2 k2),  k2 = 0..16
//             (real input: 32 real outputs per row).  n2 < 16 lies in hop w, n2 >= 16 in hop w + 1 (400 = 25 * 16), so the
//             operand is staged ONCE PER HOP as T[hop * 25 + n1][16] and the second K-half of a row is the same array
//             25 rows further down -- a descriptor offset, no second copy.  M = 128 covers 5 windows (125 rows).
//   E1        Z = Y * W800^(n1 k2) (CUDA cores), split into fp16 hi / lo, written as the stage-2 operand
//   stage 2   rows (w, k2), 17 per window;  K = (n1, re/im) = 50 -> 64;  N = (k1, re/im) = 50 -> 64: DFT-25 over n1,
//             X[k2 + 32 k1]; bins above 400 are the mirrors of the residues 17..31 (real input)
//   E2        P = re^2 + im^2 into shared memory [bin][window] for the mel stage
// The samples enter as x / 64 = a + b / 64 with a = x >> 6 in [-512, 511] and b = x & 63, both exact in fp16.
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tc_dft tc_dft.cu && ./tc_dft [n_clips]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kW = 14;                  // windows per iteration
constexpr int kHops = kW + 1;           // 15 hops of 400 samples
constexpr int kN1 = 25, kK2 = 17;
constexpr int kThreads = 384;           // three warpgroups: one stage-1 tile each
constexpr int kBins = 401;
constexpr int kClipSamples = 441000, kClipWindows = 1101;
constexpr int kBlocksPerClip = (kClipWindows + kW - 1) / kW;   // 79
constexpr int kPseudoMel = 26;

// ---- shared-memory map (bytes) ----------------------------------------------------------------------------------------
constexpr int kTRows = 408;                                   // 15 * 25 = 375 used, tiles read up to row 250 + 128 + 25
constexpr int kTArr = kTRows * 16;                            // one (piece, k-chunk) array of T
constexpr int kA2Rows = 256;                                  // 14 * 17 = 238 used
constexpr int kA2Arr = kA2Rows * 16 + 16;                     // + 16: consecutive k-chunk arrays shift by 4 banks
constexpr int kOffX = 0;                                      // region X: T -> A2 -> P, one after the other
constexpr int kSizeX = 2 * 8 * kA2Arr;                        // 65 792
constexpr int kOffRaw = kOffX + kSizeX;                       // raw PCM of the iteration: 15 hops * 800 B
constexpr int kSizeRaw = kHops * 800 + 64;
constexpr int kOffB1 = (kOffRaw + kSizeRaw + 127) / 128 * 128;   // [piece 2][half 2][chunk 2][32 rows][16 B]
constexpr int kSizeB1 = 8 * 512;
constexpr int kOffB2 = kOffB1 + kSizeB1;                      // [piece 2][k-chunk 8][64 rows][16 B]
constexpr int kSizeB2 = 16 * 1024;
constexpr int kOffTw = kOffB2 + kSizeB2;                      // [16][25] float2: W800^(n1 k2), k2 = 1..16
constexpr int kSizeTw = 16 * kN1 * 8;
constexpr int kSmem = kOffTw + kSizeTw;
constexpr int kPS = 17;                 // row stride of P[bin][window]: odd, so bins (lanes) spread over the banks
static_assert(4 * kTArr <= kSizeX && kBins * kPS * 4 <= kSizeX, "region X too small");
static_assert(kSmem <= 113 * 1024, "two CTAs per SM");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
// K-major, no swizzle ("interleave"): 8 x 16 B core matrices; rows of a core matrix 16 B apart, 8-row groups SBO apart,
// the two 16-byte k-chunks of one MMA LBO apart (cute::UMMA::SmemDescriptor, version 1, layout type 0)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return uint64_t((addr >> 4) & 0x3FFF) | (uint64_t((lbo >> 4) & 0x3FFF) << 16) | (uint64_t((sbo >> 4) & 0x3FFF) << 32) | (uint64_t(1) << 46);
}
// kind::f16: D = F32 (bits 4-5 = 1), A = B = F16 (0), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) { return (1u << 4) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

struct Tables {            // operand matrices and twiddles, built on the host, copied into shared memory by every CTA
    uint4 b1[kSizeB1 / 16];
    uint4 b2[kSizeB2 / 16];
    uint4 tw[kSizeTw / 16];
};

__global__ void __launch_bounds__(kThreads, 2)
tc_dft_kernel(const int16_t* __restrict__ pcm, const Tables* __restrict__ tables, uint32_t n_clips, float* __restrict__ out,
              float* __restrict__ dbg_power /* [kW][401] of item 0, or null */, unsigned long long* __restrict__ phase_clk /* [8] or null */) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t s_bar[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, wg = warp >> 2;
    unsigned char* X = smem + kOffX;
    int16_t* raw = reinterpret_cast<int16_t*>(smem + kOffRaw);
    const float2* tw = reinterpret_cast<const float2*>(smem + kOffTw);

    for (int i = tid; i < int(sizeof(Tables) / 16); i += kThreads)
        reinterpret_cast<uint4*>(smem + kOffB1)[i] = reinterpret_cast<const uint4*>(tables)[i];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t d1 = tmem, d2 = tmem + 96;                 // 3 x 32 columns, then 2 x 64 columns
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    constexpr uint32_t idesc1 = make_idesc_f16(128, 32), idesc2 = make_idesc_f16(128, 64);

    const uint64_t n_items = uint64_t(n_clips) * kBlocksPerClip;
    auto issue_raw = [&](uint64_t item) {                     // cp.async of the 15 hops of an item: 750 sixteen-byte chunks
        const uint64_t clip = item / kBlocksPerClip, blk = item % kBlocksPerClip;
        const int16_t* src = pcm + clip * kClipSamples + blk * (kW * 400);
        for (int c = tid; c < kHops * 50; c += kThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(raw + 8 * c)), "l"(src + 8 * c) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    uint32_t it = 0;
    unsigned long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define PHASE(i) do { if (tid == 0) { const unsigned long long tn = clock64(); ph[i] += tn - tprev; tprev = tn; } } while (0)
    if (blockIdx.x < n_items) issue_raw(blockIdx.x);
    for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                       // raw complete; region X free (previous consumer done)
        PHASE(0);
        // ---- 1. staging: thread <-> (hop, n1): 16 samples n1 + 25 j -> fp16 hi / lo, two 16-byte k-chunks each ----------
        if (tid < kHops * kN1) {
            const int16_t* s = raw + (tid / kN1) * 400 + (tid % kN1);
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint32_t v = uint32_t(uint16_t(s[25 * (2 * q)])) | (uint32_t(uint16_t(s[25 * (2 * q + 1)])) << 16);
                const uint32_t t = v ^ 0x80008000u;                                  // x + 32768 per half
                uint32_t hb = ((t >> 6) & 0x03FF03FFu) | 0x64006400u;                // fp16(1024 + (x + 32768 >> 6))
                uint32_t lb = (t & 0x003F003Fu) | 0x4C004C00u;                       // fp16(16 + (x & 63) / 64)
                const __half2 h2 = __hsub2(*reinterpret_cast<__half2*>(&hb), __floats2half2_rn(1536.f, 1536.f));
                const __half2 l2 = __hsub2(*reinterpret_cast<__half2*>(&lb), __floats2half2_rn(16.f, 16.f));
                hi[q] = *reinterpret_cast<const uint32_t*>(&h2);
                lo[q] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            unsigned char* row = X + tid * 16;                 // T[piece][chunk][row]: arrays kTArr apart
            *reinterpret_cast<uint4*>(row + 0 * kTArr) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(row + 1 * kTArr) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            *reinterpret_cast<uint4*>(row + 2 * kTArr) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4*>(row + 3 * kTArr) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        } else {                                               // rows 375..407 are read by the last tile's idle lanes: keep them finite
            for (int r = kHops * kN1 + (tid - kHops * kN1); r < kTRows; r += kThreads - kHops * kN1)
#pragma unroll
                for (int a = 0; a < 4; ++a) *reinterpret_cast<uint4*>(X + a * kTArr + r * 16) = make_uint4(0, 0, 0, 0);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        PHASE(1);
        // ---- 2. stage-1 MMAs: 3 tiles x 2 K-halves x 3 products -------------------------------------------------------
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t xb = smem_u32(X), b1 = smem_u32(smem + kOffB1);
#pragma unroll
            for (int t = 0; t < 3; ++t)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t row_off = uint32_t(125 * t + 25 * h) * 16;
                    const uint64_t a_hi = make_desc(xb + 0 * kTArr + row_off, kTArr, 128), a_lo = make_desc(xb + 2 * kTArr + row_off, kTArr, 128);
                    const uint64_t b_hi = make_desc(b1 + (0 * 2 + h) * 1024, 512, 128), b_lo = make_desc(b1 + (1 * 2 + h) * 1024, 512, 128);
                    umma_f16(d1 + 32 * t, a_lo, b_hi, idesc1, h);            // small terms first
                    umma_f16(d1 + 32 * t, a_hi, b_lo, idesc1, 1u);
                    umma_f16(d1 + 32 * t, a_hi, b_hi, idesc1, 1u);
                }
            umma_commit(&s_bar[0]);
        }
        mbar_wait(&s_bar[0], par);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        PHASE(2);
        // ---- 3. E1: twiddle, split, transpose into the stage-2 operand A2[piece][k-chunk n1 / 4][row w * 17 + k2] -------
        {
            // the K padding (k = 50..63) meets zero rows of B2 but must be finite: region X held T and P before
            for (int i = tid; i < 2 * kA2Rows * 2; i += kThreads) {
                const int piece = i / (2 * kA2Rows), r = (i >> 1) % kA2Rows, c = 6 + (i & 1);
                *reinterpret_cast<uint4*>(X + (piece * 8 + c) * kA2Arr + r * 16) = make_uint4(0, 0, 0, 0);
            }
            __syncthreads();                                   // (chunk 6 also receives n1 = 24 below)
            const int l = tid & 127;
            const int rows_here = wg < 2 ? 125 : (kW - 10) * kN1;
            float y[32];
            tmem_ld32(d1 + 32 * wg + lane_base, y);            // warp-collective: outside the row predicate
            if (l < rows_here) {
                const int w_local = 5 * wg + l / kN1, n1 = l % kN1;
                unsigned char* dst_hi = X + (0 * 8 + (n1 >> 2)) * kA2Arr + (w_local * kK2) * 16 + (n1 & 3) * 4;
                unsigned char* dst_lo = dst_hi + 8 * kA2Arr;
                auto put = [&](int k2, float zr, float zi) {
                    const __half2 h = __floats2half2_rn(zr, zi);
                    const float2 hf = __half22float2(h);
                    const __half2 lo2 = __floats2half2_rn(zr - hf.x, zi - hf.y);
                    *reinterpret_cast<__half2*>(dst_hi + k2 * 16) = h;
                    *reinterpret_cast<__half2*>(dst_lo + k2 * 16) = lo2;
                };
                put(0, y[0], 0.f);
#pragma unroll
                for (int k2 = 1; k2 < 16; ++k2) {
                    const float2 w = tw[(k2 - 1) * kN1 + n1];
                    put(k2, y[2 * k2] * w.x - y[2 * k2 + 1] * w.y, y[2 * k2] * w.y + y[2 * k2 + 1] * w.x);
                }
                const float2 w16 = tw[15 * kN1 + n1];
                put(16, y[1] * w16.x, y[1] * w16.y);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        PHASE(3);
        // ---- 4. stage-2 MMAs: 2 tiles x 4 K-steps x 3 products; meanwhile the next item's samples are requested ------------
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t xb = smem_u32(X), b2 = smem_u32(smem + kOffB2);
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const uint64_t a_hi = make_desc(xb + (0 * 8 + 2 * s) * kA2Arr + u * 128 * 16, kA2Arr, 128);
                    const uint64_t a_lo = make_desc(xb + (1 * 8 + 2 * s) * kA2Arr + u * 128 * 16, kA2Arr, 128);
                    const uint64_t b_hi = make_desc(b2 + (0 * 8 + 2 * s) * 1024, 1024, 128), b_lo = make_desc(b2 + (1 * 8 + 2 * s) * 1024, 1024, 128);
                    umma_f16(d2 + 64 * u, a_lo, b_hi, idesc2, s);
                    umma_f16(d2 + 64 * u, a_hi, b_lo, idesc2, 1u);
                    umma_f16(d2 + 64 * u, a_hi, b_hi, idesc2, 1u);
                }
            umma_commit(&s_bar[1]);
        }
        if (item + gridDim.x < n_items) issue_raw(item + gridDim.x);
        mbar_wait(&s_bar[1], par);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        PHASE(4);
        // ---- 5. E2: power spectrum into P[bin][16 windows] (region X again) ------------------------------------------------
        float* P = reinterpret_cast<float*>(X);
        if (wg < 2) {
            const int R = 128 * wg + (tid & 127);
            float x0[32], x1[32];
            tmem_ld32(d2 + 64 * wg + lane_base, x0);
            tmem_ld32(d2 + 64 * wg + 32 + lane_base, x1);
            if (R < kW * kK2) {
                const int w_local = R / kK2, k2 = R % kK2;
                const bool edge = k2 == 0 || k2 == 16;         // residues whose mirrors are their own duplicates
#pragma unroll
                for (int k1 = 0; k1 < kN1; ++k1) {
                    const float re = k1 < 16 ? x0[2 * k1] : x1[2 * k1 - 32], im = k1 < 16 ? x0[2 * k1 + 1] : x1[2 * k1 + 1 - 32];
                    const int k = k2 + 32 * k1;
                    if (k <= 400) P[k * kPS + w_local] = re * re + im * im;
                    else if (!edge) P[(800 - k) * kPS + w_local] = re * re + im * im;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        PHASE(5);
        // ---- 6. stand-in for the mel stage: 26 band sums per window ------------------------------------------------------------
        const uint64_t clip = item / kBlocksPerClip, blk = item % kBlocksPerClip;
        if (tid < kW * kPseudoMel) {
            const int w_local = tid / kPseudoMel, m = tid % kPseudoMel;
            float acc = 0.f;
#pragma unroll 5
            for (int b = 0; b < 15; ++b) acc += P[(m * 15 + b + (m == 25 ? 11 : 0)) * kPS + w_local];
            const uint64_t w = blk * kW + w_local;
            if (w < kClipWindows) out[(clip * kClipWindows + w) * kPseudoMel + m] = acc;
        }
        PHASE(6);
        if (dbg_power && item == 0)
            for (int i = tid; i < kW * kBins; i += kThreads) dbg_power[i] = P[(i % kBins) * kPS + i / kBins];
    }
    if (phase_clk && blockIdx.x == 0 && tid == 0) { for (int i = 0; i < 7; ++i) phase_clk[i] = ph[i]; phase_clk[7] = it; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
}

__global__ void fill_pcm(int16_t* p, size_t n, uint32_t seed) {   // speech-like enough for timing: two tones + noise
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        uint32_t h = uint32_t(i) * 2654435761u ^ seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        const float t = float(i % 441000) / 44100.f;
        const float v = 9000.f * __sinf(6.2831853f * 220.f * t) + 4000.f * __sinf(6.2831853f * 1730.f * t + 1.f) + float(int(h & 2047) - 1024);
        p[i] = int16_t(v);
    }
}

static uint16_t f2h(float f) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float h2f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }

int main(int argc, char** argv) {
    const uint32_t n_clips = argc > 1 ? uint32_t(atoi(argv[1])) : 10000;
    // ---- operand tables --------------------------------------------------------------------------------------------------
    Tables* tb = new Tables();
    uint16_t* b1 = reinterpret_cast<uint16_t*>(tb->b1);
    uint16_t* b2 = reinterpret_cast<uint16_t*>(tb->b2);
    float* twf = reinterpret_cast<float*>(tb->tw);
    memset(tb, 0, sizeof(Tables));
    for (int col = 0; col < 32; ++col) {                      // stage-1 output column: 0 Re Y0, 1 Re Y16, 2k Re Yk, 2k+1 Im Yk
        const int k2 = col == 0 ? 0 : (col == 1 ? 16 : col / 2);
        const bool imag = col >= 2 && (col & 1);
        for (int n2 = 0; n2 < 32; ++n2) {
            const double a = -2.0 * M_PI * double((n2 * k2) % 32) / 32.0;
            const float w = float(imag ? sin(a) : cos(a));
            const uint16_t wh = f2h(w), wl = f2h(w - h2f(wh));
            const int half = n2 / 16, chunk = (n2 % 16) / 8, e = n2 % 8;
            b1[((0 * 2 + half) * 2 + chunk) * 256 + col * 8 + e] = wh;
            b1[((1 * 2 + half) * 2 + chunk) * 256 + col * 8 + e] = wl;
        }
    }
    for (int col = 0; col < 50; ++col) {                      // stage-2 output column 2 k1 + c'
        const int k1 = col / 2, cp = col & 1;
        for (int k = 0; k < 50; ++k) {                        // K index 2 n1 + c
            const int n1 = k / 2, c = k & 1;
            const double a = -2.0 * M_PI * double((n1 * k1) % 25) / 25.0;
            const double cr = cos(a), si = sin(a);            // omega^(n1 k1) = cr + i si
            const float w = float(cp == 0 ? (c == 0 ? cr : -si) : (c == 0 ? si : cr));
            const uint16_t wh = f2h(w), wl = f2h(w - h2f(wh));
            b2[(0 * 8 + k / 8) * 512 + col * 8 + k % 8] = wh;
            b2[(1 * 8 + k / 8) * 512 + col * 8 + k % 8] = wl;
        }
    }
    for (int k2 = 1; k2 <= 16; ++k2)
        for (int n1 = 0; n1 < 25; ++n1) {
            const double a = -2.0 * M_PI * double((n1 * k2) % 800) / 800.0;
            twf[((k2 - 1) * 25 + n1) * 2 + 0] = float(cos(a));
            twf[((k2 - 1) * 25 + n1) * 2 + 1] = float(sin(a));
        }
    Tables* d_tb; CK(cudaMalloc(&d_tb, sizeof(Tables))); CK(cudaMemcpy(d_tb, tb, sizeof(Tables), cudaMemcpyHostToDevice));
    // ---- data ------------------------------------------------------------------------------------------------------------
    const size_t n_samples = size_t(n_clips) * kClipSamples + 16 * 400;
    int16_t* d_pcm; CK(cudaMalloc(&d_pcm, n_samples * 2));
    fill_pcm<<<148 * 8, 256>>>(d_pcm, n_samples, 12345u);
    float* d_out; CK(cudaMalloc(&d_out, size_t(n_clips) * kClipWindows * kPseudoMel * 4));
    float* d_dbg; CK(cudaMalloc(&d_dbg, kW * kBins * 4));
    CK(cudaDeviceSynchronize());
    CK(cudaFuncSetAttribute(tc_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    CK(cudaFuncSetAttribute(tc_dft_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    unsigned long long* d_ph; CK(cudaMalloc(&d_ph, 64)); CK(cudaMemset(d_ph, 0, 64));
    int sms = 0, per_sm = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tc_dft_kernel, kThreads, kSmem));
    printf("smem %d B per CTA, %d CTA(s) per SM, %d SMs\n", kSmem, per_sm, sms);
    { cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0); printf(" regs/SM %d, regs/block %d, smem/SM %zu, smem/block optin %zu, reserved/block %zu, max threads/SM %d, max blocks/SM %d\n",
        pr.regsPerMultiprocessor, pr.regsPerBlock, pr.sharedMemPerMultiprocessor, pr.sharedMemPerBlockOptin, pr.reservedSharedMemPerBlock, pr.maxThreadsPerMultiProcessor, pr.maxBlocksPerMultiProcessor); }
    for (int kb = 0; kb <= 116; kb += 8) { int o = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, tc_dft_kernel, kThreads, kb * 1024); printf(" occ(%d KB)=%d", kb, o); }
    { cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, tc_dft_kernel); printf("\n regs %d, static smem %zu, local %zu, maxDyn %d, carveout %d\n", fa.numRegs, fa.sharedSizeBytes, fa.localSizeBytes, fa.maxDynamicSharedSizeBytes, fa.preferredShmemCarveout); }
    const int force = argc > 2 ? atoi(argv[2]) : 0;            // CTAs per SM to launch regardless of the occupancy estimate
    const int grid = sms * (force > 0 ? force : (per_sm > 0 ? per_sm : 1));
    printf("grid %d\n", grid);
    // ---- correctness: power spectra of the first 14 windows against a float64 DFT ----------------------------------------
    tc_dft_kernel<<<1, kThreads, kSmem>>>(d_pcm, d_tb, 1, d_out, d_dbg, nullptr);
    CK(cudaDeviceSynchronize());
    std::vector<float> P(kW * kBins);
    std::vector<int16_t> x(kHops * 400);
    CK(cudaMemcpy(P.data(), d_dbg, P.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(x.data(), d_pcm, x.size() * 2, cudaMemcpyDeviceToHost));
    double worst = 0.0, worst_rel_bin = 0.0;
    for (int w = 0; w < kW; ++w) {
        std::vector<double> ref(kBins);
        double pmax = 0.0;
        for (int k = 0; k < kBins; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < 800; ++n) { const double a = -2.0 * M_PI * double((n * k) % 800) / 800.0, v = x[400 * w + n] / 64.0; re += v * cos(a); im += v * sin(a); }
            ref[k] = re * re + im * im; pmax = fmax(pmax, ref[k]);
        }
        for (int k = 0; k < kBins; ++k) {
            worst = fmax(worst, fabs(P[w * kBins + k] - ref[k]) / pmax);
            if (ref[k] > 1e-6 * pmax) worst_rel_bin = fmax(worst_rel_bin, fabs(P[w * kBins + k] - ref[k]) / ref[k]);
        }
    }
    printf("power spectrum vs float64 DFT (14 windows x 401 bins): max |err| / max bin = %.3e, max relative error on bins above 1e-6 of the peak = %.3e -> %s\n",
           worst, worst_rel_bin, worst < 1e-5 ? "OK" : "FAIL");
    // ---- timing ----------------------------------------------------------------------------------------------------------
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        tc_dft_kernel<<<grid, kThreads, kSmem>>>(d_pcm, d_tb, n_clips, d_out, nullptr, d_ph);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double wins = double(n_clips) * kClipWindows;
        printf("run %d: %u clips, %.0f windows: %.3f ms  (%.1f ns per window per SM-slot, %.2f G windows/s)\n", rep, n_clips, wins, ms,
               ms * 1e6 / wins * sms, wins / ms / 1e6);
    }
    CK(cudaGetLastError());
    unsigned long long ph[8]; CK(cudaMemcpy(ph, d_ph, 64, cudaMemcpyDeviceToHost));
    const char* names[7] = {"wait raw + barrier", "staging (PCM -> fp16 hi/lo)", "stage-1 MMAs (issue + wait)", "E1 (twiddle, split, transpose)",
                            "stage-2 MMAs (issue + wait)", "E2 (power -> smem)", "stand-in mel + store"};
    printf("CTA 0, %llu iterations of %d windows; clocks per iteration by phase (thread 0's view):\n", ph[7], kW);
    for (int i = 0; i < 7; ++i) printf("  %-34s %8.0f\n", names[i], double(ph[i]) / double(ph[7] ? ph[7] : 1));
    return worst < 1e-5 ? 0 : 1;
}
