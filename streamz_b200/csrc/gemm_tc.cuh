// Tensor-core GEMM for the SimpleNeuralNet layers (streamz-rs/src/lib.rs:880-891, 1013-1045) on sm_100a:
// tcgen05.mma (kind::tf32) with the accumulator in TMEM, operands staged in shared memory in the canonical
// K-major SWIZZLE_128B layout, completion tracked with tcgen05.commit -> mbarrier, epilogue via tcgen05.ld.
//
//   C[M, N] (+)= A[M, K] * B[N, K]^T          A and B are FP32, K contiguous ("TN": both operands K-major)
//
// Every GEMM of the forward and backward pass is brought into this one form by keeping transposed copies of the
// weights and by letting the producing epilogue also write the transposed activation (see mlp.cu), so a single,
// well-tested operand layout serves all eight products of a training step.
//
// Kernels in this file: gemm_tc_async_kernel (the default: warp-specialised cp.async ring, both operands from shared
// memory), gemm_tc_ta_kernel (opt-in: A operand in tensor memory), gemm_tc_kernel (fallback for rows that are not 16-byte
// aligned).  All three share the operand form, the 3xTF32 split and -- the first two -- the staged epilogue.
//
// Precision: PASSES = 1 is plain TF32 (10-bit mantissa inputs, FP32 accumulate).  PASSES = 3 is the split
// "3xTF32" scheme: x = hi + lo with hi = tf32(x), lo = tf32(x - hi), and C += A_lo B_hi + A_hi B_lo + A_hi B_hi,
// which restores ~FP32 accuracy (error ~2^-21 relative per product) for parity with the reference's FP32 arithmetic.
#pragma once
#include <cstdint>

#include "common.cuh"

namespace szb {
namespace tc {

constexpr int BM = 128;        // UMMA M (one CTA, cta_group::1)
constexpr int BK = 32;         // K per stage: 32 floats = 128 bytes = one SWIZZLE_128B row
constexpr int UK = 8;          // K per tcgen05.mma for tf32 (32 bytes)
constexpr int kStages = 2;
constexpr int kThreadsTc = 128;

enum EpiTc : int { TC_BIAS = 0, TC_BIAS_RELU = 1, TC_BIAS_TANH = 2, TC_MUL_DTANH = 3, TC_MUL_DRELU = 4, TC_ATOMIC = 5, TC_SOFTMAX_CE = 6 };

struct GemmArgs {
    const float* A; int lda;      // [M][lda]
    const float* B; int ldb;      // [N][ldb]
    float* C; int ldc;            // [M][ldc] (may be null when only CT is wanted)
    float* CT; int ldct;          // optional transposed output [N][ldct]
    const float* bias;            // [N]   (TC_BIAS*)
    const float* aux; int ldaux;  // TRANSPOSED auxiliary activation [N][ldaux] (TC_MUL_*): aux[n][m] pairs with C[m][n]
    int M, N, K;
    int k_chunk;                  // K range per blockIdx.z (multiple of BK); == K rounded up when no split
    // TC_SOFTMAX_CE (layer 3 of a training step, N <= BN): per-row class labels or one shared target vector, the rows that
    // survived input dropout, and the [n_used, loss] block of the gradient vector
    const uint32_t* labels; const float* target_vec; const uint8_t* valid; float* tail;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (spin > (1u << 24)) __trap();   // never hang the GPU on a protocol bug
    }
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 128 bytes, groups of 8
// rows 1024 bytes apart (SBO = 64 x 16 B), LBO = 1 (ignored for swizzled K-major), version 1 (Blackwell), layout 2.
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr >> 4) & 0x3FFF);
    d |= uint64_t(1) << 16;
    d |= uint64_t(64) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}

// cute::UMMA::InstrDescriptor for kind::tf32: D = F32 (bits 4-5 = 1), A = B = TF32 (2), both K-major, N >> 3, M >> 4.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Byte offset of the 16-byte chunk `c` (0..7) of row `r` inside a [rows][128 B] SWIZZLE_128B tile (1024-B aligned).
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return uint32_t(r) * 128u + (uint32_t(c ^ (r & 7)) << 4); }

// Nearest TF32 value (10-bit mantissa, round half up in magnitude): x - tf32_hi(x) is then exact in FP32 and at most
// 2^-11 |x|, so the three-product split carries ~2^-21 relative error per term.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x00001000u) & 0xFFFFE000u); }
// What kind::tf32 makes of an FP32 word: the 13 low mantissa bits are not read.  The warp-specialised kernel therefore
// leaves the raw tile in place as the hi operand and only writes lo = x - trunc(x) (exact; |lo| < 2^-10 |x|, itself read
// truncated: x = hi + lo to 2^-20 relative) -- one 16-byte shared-memory store per chunk less than rounding hi in place.
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <int BN, int PASSES>
struct SmemLayout {
    static constexpr int kATile = BM * BK * 4;                 // 16 KB
    static constexpr int kBTile = BN * BK * 4;
    static constexpr int kParts = PASSES == 3 ? 2 : 1;         // hi (+ lo)
    static constexpr int kStageBytes = kParts * (kATile + kBTile);
    static constexpr int kTotal = kStages * kStageBytes + 1024;  // + alignment slack
};

// Epilogue shared by both kernels: thread t <-> row m0 + t (TMEM lane t); columns in chunks of 16.
template <int BN, int EPI>
__device__ __forceinline__ void tc_epilogue(const GemmArgs& g, uint32_t tmem_d, int m0, int n0, bool have_acc) {
    const int tid = threadIdx.x & 127, warp = tid >> 5;      // TMEM lane quarter = warp % 4
    const int n_kb = have_acc ? 1 : 0;
    const int m = m0 + tid;
    const uint32_t lane_addr = tmem_d + (uint32_t(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
        if (n0 + c0 >= g.N) break;
        uint32_t r[16];
        if (n_kb > 0) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(lane_addr + uint32_t(c0)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = 0u;
        }
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int n = n0 + c0 + j;
            float x = __uint_as_float(r[j]);
            if (m < g.M && n < g.N) {
                if (EPI == TC_BIAS) x += g.bias[n];
                if (EPI == TC_BIAS_RELU) { x += g.bias[n]; x = x > 0.f ? x : 0.f; }                       // lib.rs:882
                if (EPI == TC_BIAS_TANH) x = tanhf(x + g.bias[n]);                                         // lib.rs:883
                if (EPI == TC_MUL_DTANH) { const float h = g.aux[size_t(n) * g.ldaux + m]; x *= (1.f - h * h); }   // lib.rs:1034
                if (EPI == TC_MUL_DRELU) x = g.aux[size_t(n) * g.ldaux + m] > 0.f ? x : 0.f;               // lib.rs:1040
            }
            v[j] = x;
        }
        if (m < g.M) {
            if (EPI == TC_ATOMIC) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (n0 + c0 + j < g.N) atomicAdd(&g.C[size_t(m) * g.ldc + n0 + c0 + j], v[j]);
            } else if (g.C) {
                float* crow = g.C + size_t(m) * g.ldc + n0 + c0;
                if (n0 + c0 + 16 <= g.N && (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(crow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + c0 + j < g.N) crow[j] = v[j];
                }
            }
        }
        if (EPI != TC_ATOMIC && g.CT && m < g.M) {   // lanes hold consecutive m: coalesced column stores
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (n0 + c0 + j < g.N) g.CT[size_t(n0 + c0 + j) * g.ldct + m] = v[j];
        }
    }
}

// PDL (programmatic dependent launch): every kernel of a training step releases its successor at entry and waits for its
// predecessor right before its first global-memory access, so launch latency, TMEM allocation and barrier set-up of kernel
// n + 1 overlap the tail of kernel n.  Both are no-ops for a launch without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Epilogue of the warp-specialised kernel.  A warp owns 32 accumulator rows (its TMEM lane quarter) x COLS columns and
// walks them in chunks of 32 columns:
//   row domain (lane = row m, the tcgen05.ld 32x32b layout): bias / activation / derivative factor -- the factor comes from
//     the TRANSPOSED activation aux[n][m], so the 32 lanes read 128 contiguous bytes -- and the transposed output CT[n][m]
//     (again 128 contiguous bytes per store);
//   column domain: the chunk makes one trip through a warp-private 32 x 36-float shared-memory tile and leaves as 16-byte
//     stores (or 16-byte vector reductions for the split-K gradient GEMMs) in which 8 lanes cover 128 contiguous bytes of a
//     row of C.  Before, a lane stored 16 bytes of its own row: 32 rows, 32 half-written sectors per instruction, which made
//     the epilogue as long as the K loop (7-13 us per GEMM of the batch-4096 step).
constexpr int kEpiStride = 36;                                  // floats per staged row: 16-byte aligned, conflict-free
constexpr int kEpiWarpBytes = 32 * kEpiStride * 4;              // 4608 B per warp
template <int COLS, int EPI>
__device__ __forceinline__ void tc_epilogue_staged(const GemmArgs& g, uint32_t tmem_d, int m0, int n0, bool have_acc, float* stage) {
    static_assert(COLS % 32 == 0, "a warp walks its columns in chunks of 32");
    const int lane = threadIdx.x & 31, q = (threadIdx.x >> 5) & 3;
    const int mrow0 = m0 + q * 32;
    const int m = mrow0 + lane;
    const uint32_t lane_addr = tmem_d + (uint32_t(q * 32) << 16);
    const bool c_vec = g.C && (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
#pragma unroll 1
    for (int c0 = 0; c0 < COLS; c0 += 32) {
        if (n0 + c0 >= g.N) break;
        uint32_t r[32];
        if (have_acc) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(lane_addr + uint32_t(c0)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        // the 32 factors / biases of the chunk are requested together, before anything depends on them: as one
        // load -> use -> store chain per element the loop paid a full L2 round trip 64 times per thread (13 us per GEMM)
        const bool row_ok = m < g.M;
        const bool full = (m0 + BM <= g.M) && (n0 + c0 + 32 <= g.N);      // CTA-/warp-uniform: no per-element predicates
        float a[32];
        if (EPI == TC_MUL_DTANH || EPI == TC_MUL_DRELU) {
            // L1-bypassing loads: with ~200 KB of the SM's 256 KB carved out as shared memory, the remaining L1 cannot hold a
            // line for each of the CTA's 256 outstanding 128-byte requests, and allocating loads throttle on it
            const float* ap = g.aux + size_t(n0 + c0) * g.ldaux + m;
            if (full) {
#pragma unroll
                for (int j = 0; j < 32; ++j) a[j] = __ldcg(ap + size_t(j) * g.ldaux);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) a[j] = (row_ok && n0 + c0 + j < g.N) ? __ldcg(ap + size_t(j) * g.ldaux) : 0.f;
            }
        } else if (EPI != TC_ATOMIC) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = n0 + c0 + j;
                a[j] = n < g.N ? __ldg(g.bias + n) : 0.f;
            }
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(r[j]);
            if (EPI == TC_BIAS) x += a[j];
            if (EPI == TC_BIAS_RELU) { x += a[j]; x = x > 0.f ? x : 0.f; }      // lib.rs:882
            if (EPI == TC_BIAS_TANH) x = tanhf(x + a[j]);                        // lib.rs:883
            if (EPI == TC_MUL_DTANH) x *= (1.f - a[j] * a[j]);                   // lib.rs:1034
            if (EPI == TC_MUL_DRELU) x = a[j] > 0.f ? x : 0.f;                   // lib.rs:1040
            v[j] = x;
        }
        if (EPI != TC_ATOMIC && g.CT && row_ok) {
            float* tp = g.CT + size_t(n0 + c0) * g.ldct + m;
            if (full) {
#pragma unroll
                for (int j = 0; j < 32; ++j) tp[size_t(j) * g.ldct] = v[j];
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (n0 + c0 + j < g.N) tp[size_t(j) * g.ldct] = v[j];
            }
        }
        if (g.C) {
            float* mine = stage + lane * kEpiStride;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(mine + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            __syncwarp();
            const int cc = (lane & 7) * 4, n = n0 + c0 + cc;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + (lane >> 3), mm = mrow0 + rr;
                const float4 w = *reinterpret_cast<const float4*>(stage + rr * kEpiStride + cc);
                if (mm < g.M && n < g.N) {
                    float* dst = g.C + size_t(mm) * g.ldc + n;
                    if (c_vec && n + 4 <= g.N) {
                        if (EPI == TC_ATOMIC) red_add_v4(dst, w);
                        else *reinterpret_cast<float4*>(dst) = w;
                    } else {
                        const float e[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (n + t < g.N) {
                                if (EPI == TC_ATOMIC) atomicAdd(dst + t, e[t]);
                                else dst[t] = e[t];
                            }
                    }
                }
            }
            __syncwarp();
        }
    }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// One CTA = one 128 x BN tile of C (x one K split).  128 threads: all of them stage operands; one issues the MMAs;
// in the epilogue thread t owns accumulator row t (TMEM lane t).
template <int BN, int PASSES, int EPI>
__global__ void __launch_bounds__(kThreadsTc) gemm_tc_kernel(const GemmArgs g) {
    static_assert(BN == 64 || BN == 128 || BN == 256, "BN must be 64, 128 or 256");
    using SL = SmemLayout<BN, PASSES>;
    extern __shared__ unsigned char tc_smem_raw[];
    __shared__ uint64_t s_bar[kStages];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5;
    pdl_launch_dependents();
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kb0 = blockIdx.z * g.k_chunk, kb1 = min(g.K, kb0 + g.k_chunk);
    const int n_kb = (kb1 - kb0 + BK - 1) / BK;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = s_tmem;
    constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
    pdl_wait();

    const bool a_vec = (g.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
    const bool b_vec = (g.ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0);

    // A [rows][32]-float K-slab is 8 sixteen-byte chunks per row.  Thread t owns chunk (t & 7) of rows (t >> 3) + 16 u:
    // 8 consecutive threads read one 128-byte row segment.  The slab of k-block kb + 1 is fetched into registers while
    // the tensor core works on k-block kb, so global/L2 latency is paid once per tile, not once per k-block.
    constexpr int kRa = BM * 8 / kThreadsTc, kRb = BN * 8 / kThreadsTc;
    const int lr = tid >> 3, lc = tid & 7;
    auto gload = [&](const float* src, int ld, int row0, int rows_end, int k0, int kend, bool vec, int u) -> float4 {
        const int gr = row0 + lr + 16 * u, gk = k0 + lc * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < rows_end && gk < kend) {
            const float* p = src + size_t(gr) * ld + gk;
            if (vec && gk + 4 <= kend) {
                v = __ldg(reinterpret_cast<const float4*>(p));
            } else {
                v.x = __ldg(p);
                if (gk + 1 < kend) v.y = __ldg(p + 1);
                if (gk + 2 < kend) v.z = __ldg(p + 2);
                if (gk + 3 < kend) v.w = __ldg(p + 3);
            }
        }
        return v;
    };
    auto sstore = [&](unsigned char* hi, unsigned char* lo, int u, float4 v) {
        const uint32_t off = sw128_off(lr + 16 * u, lc);
        if (PASSES == 3) {
            const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            *reinterpret_cast<float4*>(hi + off) = h;
            *reinterpret_cast<float4*>(lo + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        } else {
            *reinterpret_cast<float4*>(hi + off) = v;
        }
    };
    float4 ra[kRa], rb[kRb];
    auto fetch = [&](int kb) {
        const int k0 = kb0 + kb * BK;
#pragma unroll
        for (int u = 0; u < kRa; ++u) ra[u] = gload(g.A, g.lda, m0, g.M, k0, kb1, a_vec, u);
#pragma unroll
        for (int u = 0; u < kRb; ++u) rb[u] = gload(g.B, g.ldb, n0, g.N, k0, kb1, b_vec, u);
    };
    if (n_kb > 0) fetch(0);

    for (int kb = 0; kb < n_kb; ++kb) {
        const int s = kb % kStages;
        if (kb >= kStages) mbar_wait(&s_bar[s], uint32_t((kb / kStages - 1) & 1));   // MMAs that read this stage are done
        unsigned char* st = smem + s * SL::kStageBytes;
        unsigned char* a_hi = st;
        unsigned char* b_hi = st + SL::kATile;
        unsigned char* a_lo = st + SL::kATile + SL::kBTile;
        unsigned char* b_lo = a_lo + SL::kATile;
#pragma unroll
        for (int u = 0; u < kRa; ++u) sstore(a_hi, a_lo, u, ra[u]);
#pragma unroll
        for (int u = 0; u < kRb; ++u) sstore(b_hi, b_lo, u, rb[u]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t da_hi = make_desc_k_sw128(smem_u32(a_hi)), db_hi = make_desc_k_sw128(smem_u32(b_hi));
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
                const uint64_t adv = uint64_t((k * UK * 4) >> 4);      // 32 bytes per k-step inside the swizzle atom
                const uint32_t acc0 = (kb > 0 || k > 0) ? 1u : 0u;
                if (PASSES == 3) {
                    const uint64_t da_lo = make_desc_k_sw128(smem_u32(a_lo)), db_lo = make_desc_k_sw128(smem_u32(b_lo));
                    umma_tf32(tmem_d, da_lo + adv, db_hi + adv, idesc, acc0);   // small terms first
                    umma_tf32(tmem_d, da_hi + adv, db_lo + adv, idesc, 1u);
                    umma_tf32(tmem_d, da_hi + adv, db_hi + adv, idesc, 1u);
                } else {
                    umma_tf32(tmem_d, da_hi + adv, db_hi + adv, idesc, acc0);
                }
            }
            umma_commit(&s_bar[s]);   // arrives when every MMA issued so far has completed
        }
        if (kb + 1 < n_kb) fetch(kb + 1);
    }
    // the last commit covers all MMAs of the tile
    if (n_kb > 0) {
        const int last = n_kb - 1;
        mbar_wait(&s_bar[last % kStages], uint32_t((last / kStages) & 1));
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    tc_epilogue<BN, EPI>(g, tmem_d, m0, n0, n_kb > 0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(BN) : "memory");
}


// ---- cp.async-pipelined variant (the one used whenever rows are 16-byte aligned) ---------------------------------------
// Raw FP32 K-slabs flow global -> shared with cp.async (LDGSTS, zero-filled past the matrix edge) through a ring of
// kAStages stages, issued two k-blocks ahead of the tensor core; each thread then turns ITS OWN sixteen 16-byte chunks
// into the lo tile (3xTF32; the raw tile itself serves as hi, see tf32_trunc), so no barrier is needed between the copy and the split.
// The copies run kDist k-blocks ahead of the tensor core through a ring of kDist + 2 stages (the slot refilled at
// iteration kb held k-block kb - 2, whose MMAs are the ones the producers wait for anyway before reusing a lo buffer).
// MEASURED ON B200: deepening the distance from 2 to 5 k-blocks does not move the training step (180 vs 183 us at batch
// 4096): the step is bound by the chain of 14 dependent launches, not by operand latency inside a GEMM.  Kept at 2.
template <int BN, int PASSES>
struct SmemLayoutAsync {
    static constexpr int kATile = BM * BK * 4;
    static constexpr int kBTile = BN * BK * 4;
    static constexpr int kStageBytes = kATile + kBTile;
    static constexpr int kLoBytes = PASSES == 3 ? 2 * kStageBytes : 0;
    static constexpr int kDist = 2;
    static constexpr int kAStages = kDist + 2;
    static constexpr int kTotal = kAStages * kStageBytes + kLoBytes;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

#ifdef SZB_GEMM_TRACE
__device__ unsigned long long* g_gemm_trace = nullptr;   // tools/micro/gemm_step_bench.cu: per-CTA phase stamps (ns)
__device__ __forceinline__ void trace_stamp(int slot) {
    if (g_gemm_trace && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        g_gemm_trace[size_t(cta) * 8 + slot] = t;
        // slot 7: SM clocks between stamps 1 and 4 (set-up done .. last MMA complete), to read the stamps in clocks as well
        if (slot == 1) g_gemm_trace[size_t(cta) * 8 + 7] = (unsigned long long)clock64();
        if (slot == 4) g_gemm_trace[size_t(cta) * 8 + 7] = (unsigned long long)clock64() - g_gemm_trace[size_t(cta) * 8 + 7];
    }
}
#define SZB_TRACE(slot) trace_stamp(slot)
#else
#define SZB_TRACE(slot)
#endif

constexpr int kProducerThreads = 256;                    // warps 0-7: stage + split operands, then run the epilogue
constexpr int kThreadsAsync = kProducerThreads + 32;      // warp 8: one elected lane issues the tcgen05.mma stream

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp-specialised pipeline, no CTA-wide barrier inside the K loop:
//   producers (256 threads): wait free[(kb+2)%4] -> cp.async k-block kb+2 -> wait own copies of k-block kb -> split into
//                            the lo tile (raw = hi) -> fence.proxy.async -> arrive on full[kb%4]
//   issuer (1 thread):       wait full[kb%4] -> tcgen05.mma x 4 (x3 passes) -> tcgen05.commit -> free[kb%4]
template <int BN, int PASSES, int EPI>
__global__ void __launch_bounds__(kThreadsAsync) gemm_tc_async_kernel(const GemmArgs g) {
    using SL = SmemLayoutAsync<BN, PASSES>;
    constexpr int kAStages = SL::kAStages, kDist = SL::kDist;
    extern __shared__ __align__(1024) unsigned char tc_smem[];   // SWIZZLE_128B tiles need 1024-byte alignment
    __shared__ uint64_t s_full[kAStages], s_free[kAStages];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5;
    SZB_TRACE(0);
    pdl_launch_dependents();
    const uint32_t smem_base = smem_u32(tc_smem);
    if ((smem_base & 1023u) != 0) __trap();
    const uint32_t lo_base = smem_base + SL::kAStages * SL::kStageBytes;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kb0 = blockIdx.z * g.k_chunk, kb1 = min(g.K, kb0 + g.k_chunk);
    const int n_kb = (kb1 - kb0 + BK - 1) / BK;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < kAStages; ++s) {
            mbar_init(&s_full[s], kProducerThreads);
            mbar_init(&s_free[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = s_tmem;
    constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
    pdl_wait();          // everything above is CTA-local; from here on the kernel reads what its predecessor wrote
    SZB_TRACE(1);

    if (warp < kProducerThreads / 32) {
        // ------------------------------------------------ producers ------------------------------------------------
        constexpr int kRows = kProducerThreads / 8;                 // rows covered per pass (32)
        constexpr int kRa = BM / kRows, kRb = BN / kRows;           // 4 + 4 sixteen-byte chunks per thread and k-block
        const int lr = tid >> 3, lc = tid & 7;
        auto issue = [&](int kb) {    // cp.async of k-block kb into stage kb % kAStages
            const uint32_t st = smem_base + (kb % kAStages) * SL::kStageBytes;
            const int gk = kb0 + kb * BK + lc * 4;
            const int kleft = kb1 - gk;
            const uint32_t kbytes = kleft >= 4 ? 16u : (kleft > 0 ? uint32_t(kleft) * 4u : 0u);
#pragma unroll
            for (int u = 0; u < kRa; ++u) {
                const int r = lr + kRows * u, gr = m0 + r;
                const bool ok = gr < g.M && kbytes > 0;
                cp_async16(st + sw128_off(r, lc), ok ? static_cast<const void*>(g.A + size_t(gr) * g.lda + gk) : static_cast<const void*>(g.A),
                           ok ? kbytes : 0u);
            }
#pragma unroll
            for (int u = 0; u < kRb; ++u) {
                const int r = lr + kRows * u, gr = n0 + r;
                const bool ok = gr < g.N && kbytes > 0;
                cp_async16(st + SL::kATile + sw128_off(r, lc),
                           ok ? static_cast<const void*>(g.B + size_t(gr) * g.ldb + gk) : static_cast<const void*>(g.B), ok ? kbytes : 0u);
            }
        };
        for (int j = 0; j < kDist; ++j) {        // prologue: kDist k-blocks in flight
            if (j < n_kb) issue(j);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int kb = 0; kb < n_kb; ++kb) {
            // MMAs of k-block kb - 2 done => raw stage (kb + kDist) % kAStages and lo buffer kb % 2 are free again
            if (kb >= 2) mbar_wait(&s_free[(kb - 2) % kAStages], uint32_t(((kb - 2) / kAStages) & 1));
            if (kb + kDist < n_kb) issue(kb + kDist);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(kDist) : "memory");   // this thread's chunks of k-block kb have landed
            const uint32_t st = smem_base + (kb % kAStages) * SL::kStageBytes;
            const uint32_t lo = lo_base + (kb & 1) * SL::kStageBytes;
            if (PASSES == 3) {
#pragma unroll
                for (int u = 0; u < kRa + kRb; ++u) {
                    const uint32_t off = (u < kRa ? 0u : uint32_t(SL::kATile)) + sw128_off(lr + kRows * (u < kRa ? u : u - kRa), lc);
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(st + off));
                    const float4 h = make_float4(tf32_trunc(v.x), tf32_trunc(v.y), tf32_trunc(v.z), tf32_trunc(v.w));
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo + off), "f"(v.x - h.x), "f"(v.y - h.y), "f"(v.z - h.z),
                                 "f"(v.w - h.w)
                                 : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my generic-proxy writes -> visible to the tensor core
            mbar_arrive(&s_full[kb % kAStages]);
            if (kb == 0) SZB_TRACE(2);
        }
        SZB_TRACE(3);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (tid == kProducerThreads) {
        // ------------------------------------------------ MMA issuer -----------------------------------------------
        for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(&s_full[kb % kAStages], uint32_t((kb / kAStages) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t st = smem_base + (kb % kAStages) * SL::kStageBytes;
            const uint32_t lo = lo_base + (kb & 1) * SL::kStageBytes;
            const uint64_t da_hi = make_desc_k_sw128(st), db_hi = make_desc_k_sw128(st + SL::kATile);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
                const uint64_t adv = uint64_t((k * UK * 4) >> 4);
                const uint32_t acc0 = (kb > 0 || k > 0) ? 1u : 0u;
                if (PASSES == 3) {
                    const uint64_t da_lo = make_desc_k_sw128(lo), db_lo = make_desc_k_sw128(lo + SL::kATile);
                    umma_tf32(tmem_d, da_lo + adv, db_hi + adv, idesc, acc0);
                    umma_tf32(tmem_d, da_hi + adv, db_lo + adv, idesc, 1u);
                    umma_tf32(tmem_d, da_hi + adv, db_hi + adv, idesc, 1u);
                } else {
                    umma_tf32(tmem_d, da_hi + adv, db_hi + adv, idesc, acc0);
                }
            }
            umma_commit(&s_free[kb % kAStages]);   // arrives when every MMA issued so far has completed
        }
    }
    if (warp < kProducerThreads / 32) {
        if (n_kb > 0) {
            const int last = n_kb - 1;
            mbar_wait(&s_free[last % kAStages], uint32_t((last / kAStages) & 1));   // the last commit covers the whole tile
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        SZB_TRACE(4);
        // two warpgroups split the columns: warps 0-3 take [0, BN/2), warps 4-7 take [BN/2, BN).  Every MMA has completed
        // and every cp.async has landed (the MMAs consumed them), so the operand ring is free to stage the output tile.
        tc_epilogue_staged<BN / 2, EPI>(g, tmem_d + uint32_t((warp >> 2) * (BN / 2)), m0, n0 + (warp >> 2) * (BN / 2), n_kb > 0,
                                        reinterpret_cast<float*>(tc_smem + warp * kEpiWarpBytes));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    SZB_TRACE(5);
    __syncthreads();
    SZB_TRACE(6);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(BN) : "memory");
}

// ---- A operand in tensor memory (opt-in: szb_ctx::gemm_ta / SZB_GEMM_TA=1; validated layout: tools/micro/tmem_a_gemm.cu) ----
// The warp-specialised kernel above is bound by shared-memory bandwidth: per 32-wide k-block a 128 x 128 tile moves 96 KB of
// staging traffic and the tensor core reads 96 KB of operands, half of them the A slices that each of the three 3xTF32
// products fetches again.  Here the raw A slab is staged once (coalesced cp.async, as before), then thread (row m, column
// half h) reads ITS 16 floats of the k-block back (conflict-free: the 128-byte swizzle spreads 8 rows over 8 chunk
// positions), and writes them -- and their lo parts -- into TMEM (lane = row, column = k, the cute::UMMA::tmem_frg layout of
// an M = 128 A fragment) with tcgen05.st; the MMAs take A from TMEM and only B from shared memory.  A's share of the
// shared-memory traffic falls from 96 KB per k-block (fill, split read, lo write, three MMA reads of hi/lo) to 32 KB.
// (First version, measured: rows fetched straight from global memory into registers -- correct, but a 16-byte-per-lane
// row gather costs 32 L1 wavefronts per instruction, as much pipe time as the shared-memory traffic it replaced.)
#ifndef SZB_TA_DIST
#define SZB_TA_DIST 2
#endif
#ifndef SZB_TA_NBUF
#define SZB_TA_NBUF 2
#endif
template <int BN, int PASSES>
struct SmemLayoutTa {
    static constexpr int kATile = BM * BK * 4;                       // raw A slab: staged (coalesced cp.async), read once
    static constexpr int kBTile = BN * BK * 4;
    static constexpr int kStageBytes = kATile + kBTile;
    static constexpr int kDist = SZB_TA_DIST;                        // k-blocks of cp.async in flight ahead of the tensor core
    static constexpr int kNBuf = SZB_TA_NBUF;                        // lo tiles of B / TMEM buffers of A: k-blocks the producers may run ahead of the MMAs
    static constexpr int kStages = kDist + kNBuf;
    static constexpr int kLoBytes = PASSES == 3 ? kNBuf * kBTile : 0;    // lo tiles of B only: A's lo part goes to TMEM
    static constexpr int kRing = kStages * kStageBytes + kLoBytes;
    static constexpr int kTotal = kRing > 116 * 1024 ? kRing : 116 * 1024;   // > half an SM: one CTA per SM, as the grids assume
    static constexpr int kACols = 32 * (PASSES == 3 ? 2 : 1);            // TMEM columns of one A buffer: [hi 32 | lo 32]
    static constexpr int kTmemCols = (BN + kNBuf * kACols) <= 128 ? 128 : ((BN + kNBuf * kACols) <= 256 ? 256 : 512);
    static_assert(BN + kNBuf * kACols <= 512, "tensor memory has 512 columns");
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
          "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
        : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int BN, int PASSES, int EPI>
__device__ __forceinline__ void gemm_tc_ta_body(const GemmArgs& g, const int bx, const int by, const int bz) {
    using SL = SmemLayoutTa<BN, PASSES>;
    constexpr int kStages = SL::kStages, kDist = SL::kDist, kNBuf = SL::kNBuf;
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ uint64_t s_full[kStages], s_free[kStages];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    SZB_TRACE(0);
    const uint32_t smem_base = smem_u32(tc_smem);
    if ((smem_base & 1023u) != 0) __trap();
    const uint32_t lo_base = smem_base + kStages * SL::kStageBytes;
    const int m0 = by * BM, n0 = bx * BN;
    const int kb0 = bz * g.k_chunk, kb1 = min(g.K, kb0 + g.k_chunk);
    const int n_kb = (kb1 - kb0 + BK - 1) / BK;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(SL::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&s_full[s], kProducerThreads);
            mbar_init(&s_free[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = s_tmem;                      // columns [0, BN): accumulator
    const uint32_t tmem_a = s_tmem + BN;                 // then two A buffers of kACols columns
    constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
    // Dependents are released only now, with this CTA's tensor memory allocated: the ring is small enough for CTAs of the
    // NEXT kernel to share the SM, and if they could take the TMEM columns first, a CTA of this grid would wait in
    // tcgen05.alloc for columns held by CTAs that wait (griddepcontrol.wait) for this grid to finish.
    pdl_launch_dependents();
    pdl_wait();
    SZB_TRACE(1);

    if (warp < kProducerThreads / 32) {
        // ------------------------------------------------ producers ------------------------------------------------
        // A and B raw slabs: cp.async ring as in gemm_tc_async_kernel; thread <-> chunk (tid & 7) of rows (tid >> 3) + 32 u
        constexpr int kRows = kProducerThreads / 8;
        constexpr int kRa = BM / kRows, kRb = BN / kRows;
        const int lr = tid >> 3, lc = tid & 7;
        auto issue = [&](int kb) {
            const uint32_t st = smem_base + (kb % kStages) * SL::kStageBytes;
            const int gk = kb0 + kb * BK + lc * 4;
            const int kleft = kb1 - gk;
            const uint32_t kbytes = kleft >= 4 ? 16u : (kleft > 0 ? uint32_t(kleft) * 4u : 0u);
#pragma unroll
            for (int u = 0; u < kRa; ++u) {
                const int r = lr + kRows * u, gr = m0 + r;
                const bool ok = gr < g.M && kbytes > 0;
                cp_async16(st + sw128_off(r, lc), ok ? static_cast<const void*>(g.A + size_t(gr) * g.lda + gk) : static_cast<const void*>(g.A),
                           ok ? kbytes : 0u);
            }
#pragma unroll
            for (int u = 0; u < kRb; ++u) {
                const int r = lr + kRows * u, gr = n0 + r;
                const bool ok = gr < g.N && kbytes > 0;
                cp_async16(st + SL::kATile + sw128_off(r, lc),
                           ok ? static_cast<const void*>(g.B + size_t(gr) * g.ldb + gk) : static_cast<const void*>(g.B), ok ? kbytes : 0u);
            }
        };
        // A -> TMEM: thread <-> row 32 (warp & 3) + lane of the tile (its TMEM lane), chunks [4 h, 4 h + 4) of the k-block
        const int q = warp & 3, h = warp >> 2;
        const uint32_t a_lane = uint32_t(q * 32) << 16;
        for (int j = 0; j < kDist; ++j) {
#ifndef SZB_X_NOLOAD
            if (j < n_kb) issue(j);
#endif
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int kb = 0; kb < n_kb; ++kb) {
            // MMAs of k-block kb - kNBuf done => ring stage (kb + kDist) % kStages, lo buffer and TMEM A buffer kb % kNBuf are free
            if (kb >= kNBuf) {
                mbar_wait(&s_free[(kb - kNBuf) % kStages], uint32_t(((kb - kNBuf) / kStages) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
#ifndef SZB_X_NOLOAD
            if (kb + kDist < n_kb) issue(kb + kDist);
#endif
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(kDist) : "memory");   // this thread's chunks of k-block kb have landed
            const uint32_t st = smem_base + (kb % kStages) * SL::kStageBytes;
#ifndef SZB_X_NOSPLIT
            // ---- B lo split first (own chunks only, no barrier needed), so the barrier below has less to wait for
            if (PASSES == 3) {
                const uint32_t lo = lo_base + (kb % kNBuf) * SL::kBTile;
#pragma unroll
                for (int u = 0; u < kRb; ++u) {
                    const uint32_t off = sw128_off(lr + kRows * u, lc);
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(st + SL::kATile + off));
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo + off), "f"(v.x - tf32_trunc(v.x)),
                                 "f"(v.y - tf32_trunc(v.y)), "f"(v.z - tf32_trunc(v.z)), "f"(v.w - tf32_trunc(v.w))
                                 : "memory");
                }
            }
            // ---- A -> TMEM: a row's chunks were copied by other threads: producers-only barrier, then read the own row back
            asm volatile("bar.sync 1, %0;" ::"n"(kProducerThreads) : "memory");
            {
                float hi[16], lo[16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                                 : "r"(st + sw128_off(q * 32 + lane, 4 * h + c)));
                    hi[4 * c + 0] = v.x; hi[4 * c + 1] = v.y; hi[4 * c + 2] = v.z; hi[4 * c + 3] = v.w;
                }
                const uint32_t abuf = tmem_a + a_lane + uint32_t((kb % kNBuf) * SL::kACols + h * 16);
                tmem_st16(abuf, hi);                     // kind::tf32 reads the upper 19 bits: the raw word is the hi operand
                if (PASSES == 3) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) lo[j] = hi[j] - tf32_trunc(hi[j]);
                    tmem_st16(abuf + 32, lo);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#endif
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&s_full[kb % kStages]);
            if (kb == 0) SZB_TRACE(2);
        }
        SZB_TRACE(3);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (tid == kProducerThreads) {
        // ------------------------------------------------ MMA issuer -----------------------------------------------
        for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(&s_full[kb % kStages], uint32_t((kb / kStages) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t db_hi = make_desc_k_sw128(smem_base + (kb % kStages) * SL::kStageBytes + SL::kATile);
            const uint64_t db_lo = make_desc_k_sw128(lo_base + (kb % kNBuf) * SL::kBTile);
            const uint32_t a_hi = tmem_a + uint32_t((kb % kNBuf) * SL::kACols), a_lo = a_hi + 32;
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
                const uint64_t adv = uint64_t((k * UK * 4) >> 4);
                const uint32_t acol = uint32_t(k * UK);
                const uint32_t acc0 = (kb > 0 || k > 0) ? 1u : 0u;
#ifdef SZB_X_NOMMA
                if (kb > 0) continue;      // experiment: one k-block of MMAs only (the accumulator is defined), then bare commits
#endif
                if (PASSES == 3) {
                    umma_tf32_ts(tmem_d, a_lo + acol, db_hi + adv, idesc, acc0);
                    umma_tf32_ts(tmem_d, a_hi + acol, db_lo + adv, idesc, 1u);
                    umma_tf32_ts(tmem_d, a_hi + acol, db_hi + adv, idesc, 1u);
                } else {
                    umma_tf32_ts(tmem_d, a_hi + acol, db_hi + adv, idesc, acc0);
                }
            }
            umma_commit(&s_free[kb % kStages]);
        }
    }
    if (warp < kProducerThreads / 32) {
        if (n_kb > 0) {
            const int last = n_kb - 1;
            mbar_wait(&s_free[last % kStages], uint32_t((last / kStages) & 1));
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        SZB_TRACE(4);
        tc_epilogue_staged<BN / 2, EPI>(g, tmem_d + uint32_t((warp >> 2) * (BN / 2)), m0, n0 + (warp >> 2) * (BN / 2), n_kb > 0,
                                        reinterpret_cast<float*>(tc_smem + warp * kEpiWarpBytes));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    SZB_TRACE(5);
    __syncthreads();
    SZB_TRACE(6);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(SL::kTmemCols) : "memory");
}

template <int BN, int PASSES, int EPI>
__global__ void __launch_bounds__(kThreadsAsync) gemm_tc_ta_kernel(const GemmArgs g) {
    gemm_tc_ta_body<BN, PASSES, EPI>(g, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Several independent products in ONE launch (the three weight-gradient GEMMs of a training step: each depends only on the
// activations and deltas, none on another).  As three launches each of them was cut into ~148 short K ranges to fill the SMs
// -- dW3 into 49 ranges of three k-blocks -- and paid set-up, pipeline fill and a full tile of vector reductions per range;
// one launch shares the 148 SMs among all tiles, so a CTA walks a K range several times longer and the reduction traffic
// falls by the same factor.  CTA b serves problem p with first[p] <= b < first[p + 1]; inside a problem the CTAs are ordered
// (column tile, row tile, K range) like the grid of the single-problem kernel.
constexpr int kMaxGroup = 3;
struct GroupArgs {
    GemmArgs g[kMaxGroup];
    int first[kMaxGroup + 1];
    int tiles_n[kMaxGroup], tiles_m[kMaxGroup];
    int count;
};
template <int BN, int PASSES, int EPI>
__global__ void __launch_bounds__(kThreadsAsync) gemm_tc_ta_group_kernel(const __grid_constant__ GroupArgs ga) {
    const int b = blockIdx.x;
    int p = 0;
#pragma unroll
    for (int i = 1; i < kMaxGroup; ++i)
        if (i < ga.count && b >= ga.first[i]) p = i;
    const int local = b - ga.first[p];
    const int tn = ga.tiles_n[p], tm = ga.tiles_m[p];
    gemm_tc_ta_body<BN, PASSES, EPI>(ga.g[p], local % tn, (local / tn) % tm, local / (tn * tm));
}

template <int BN, int PASSES, int EPI>
szb_status launch_gemm_tc(szb_ctx* ctx, GemmArgs g, int split_k) {
    if (g.M <= 0 || g.N <= 0) return SZB_OK;
    using SL = SmemLayout<BN, PASSES>;
    using SLA = SmemLayoutAsync<BN, PASSES>;
    const int kb_total = (g.K + BK - 1) / BK;
    split_k = EPI == TC_ATOMIC ? std::max(1, std::min(split_k, kb_total)) : 1;
    g.k_chunk = ((kb_total + split_k - 1) / split_k) * BK;
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, (g.K + g.k_chunk - 1) / g.k_chunk);
    const bool aligned = g.lda % 4 == 0 && g.ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(g.B) & 15) == 0;
    if (aligned && ctx->gemm_ta) {
        using SLT = SmemLayoutTa<BN, PASSES>;
        static bool attr_set[64] = {};
        if (!attr_set[ctx->device & 63]) {
            SZB_CUDA(cudaFuncSetAttribute(gemm_tc_ta_kernel<BN, PASSES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLT::kTotal));
            attr_set[ctx->device & 63] = true;
        }
        SZB_CUDA(launch_pdl(ctx, gemm_tc_ta_kernel<BN, PASSES, EPI>, grid, dim3(kThreadsAsync), size_t(SLT::kTotal), g));
    } else if (aligned) {
        static bool attr_set[64] = {};      // per template instantiation and device
        if (!attr_set[ctx->device & 63]) {
            SZB_CUDA(cudaFuncSetAttribute(gemm_tc_async_kernel<BN, PASSES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLA::kTotal));
            attr_set[ctx->device & 63] = true;
        }
        SZB_CUDA(launch_pdl(ctx, gemm_tc_async_kernel<BN, PASSES, EPI>, grid, dim3(kThreadsAsync), size_t(SLA::kTotal), g));
    } else {
        static bool attr_set[64] = {};
        if (!attr_set[ctx->device & 63]) {
            SZB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, PASSES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::kTotal));
            attr_set[ctx->device & 63] = true;
        }
        SZB_CUDA(launch_pdl(ctx, gemm_tc_kernel<BN, PASSES, EPI>, grid, dim3(kThreadsTc), size_t(SL::kTotal), g));
    }
    SZB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return SZB_OK;
}

inline bool gemm_operands_aligned(const GemmArgs& g) {
    return g.lda % 4 == 0 && g.ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0;
}

// The weight-gradient GEMMs of a step (split-K, vector reductions into the gradient vector) as one launch of
// gemm_tc_ta_group_kernel.  The K ranges are sized so that all tiles of all problems together fill the SMs once.
// *done = false (nothing launched) when an operand is not 16-byte aligned or the TMEM-A kernel is switched off.
template <int PASSES>
szb_status launch_gemm_tc_group(szb_ctx* ctx, const GemmArgs* gs, int count, bool* done) {
    constexpr int BN = 128;
    *done = false;
    if (count < 1 || count > kMaxGroup || !ctx->gemm_ta) return SZB_OK;
    GroupArgs ga{};
    int total_tiles = 0;
    for (int p = 0; p < count; ++p) {
        if (gs[p].M <= 0 || gs[p].N <= 0 || gs[p].K <= 0 || !gemm_operands_aligned(gs[p])) return SZB_OK;
        ga.g[p] = gs[p];
        ga.tiles_n[p] = (gs[p].N + BN - 1) / BN;
        ga.tiles_m[p] = (gs[p].M + BM - 1) / BM;
        total_tiles += ga.tiles_n[p] * ga.tiles_m[p];
    }
    const int split = std::max(1, ctx->sm_count / std::max(1, total_tiles));
    int ctas = 0;
    for (int p = 0; p < count; ++p) {
        const int kb_total = (gs[p].K + BK - 1) / BK;
        const int sp = std::max(1, std::min(split, kb_total));
        ga.g[p].k_chunk = ((kb_total + sp - 1) / sp) * BK;
        const int nz = (gs[p].K + ga.g[p].k_chunk - 1) / ga.g[p].k_chunk;      // every K range is non-empty
        ga.first[p] = ctas;
        ctas += ga.tiles_n[p] * ga.tiles_m[p] * nz;
    }
    for (int p = count; p <= kMaxGroup; ++p) ga.first[p] = ctas;
    ga.count = count;
    using SLT = SmemLayoutTa<BN, PASSES>;
    static bool attr_set[64] = {};
    if (!attr_set[ctx->device & 63]) {
        SZB_CUDA(cudaFuncSetAttribute(gemm_tc_ta_group_kernel<BN, PASSES, TC_ATOMIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLT::kTotal));
        attr_set[ctx->device & 63] = true;
    }
    SZB_CUDA(launch_pdl(ctx, gemm_tc_ta_group_kernel<BN, PASSES, TC_ATOMIC>, dim3(ctas), dim3(kThreadsAsync), size_t(SLT::kTotal), ga));
    SZB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    *done = true;
    return SZB_OK;
}

}  // namespace tc
}  // namespace szb
