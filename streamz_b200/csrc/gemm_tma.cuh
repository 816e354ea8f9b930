// Third generation of the dense-layer GEMM (same operand form and epilogues as gemm_tc.cuh: C[M,N] (+)= A[M,K] * B[N,K]^T, both
// operands K-contiguous FP32, 3xTF32 split, accumulator and the A operand in tensor memory):
//
//   * operands travel global -> shared memory by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B tensor maps, completion on an
//     mbarrier with expect_tx; rows / columns past the matrix edge are zero-filled by the unit) -- one thread issues two bulk
//     copies per k-block where 256 threads issued eight 16-byte cp.async each and then waited on their own copy groups;
//   * the work between "tile landed" and "MMAs may start" is split over two producer groups that run concurrently:
//     warps 0-7 move the A slab into tensor memory (own row read back from the swizzled tile, lo = x - trunc(x), two
//     tcgen05.st), warps 8-15 write the lo tile of B.  In gemm_tc_ta_kernel the same eight warps did both in sequence,
//     separated by a block barrier (the copies of a row were issued by other threads); with TMA the mbarrier makes the whole
//     tile visible, so the barrier is gone too;
//   * one mbarrier arrival per warp (after __syncwarp) instead of one per thread: 16 arrivals per k-block instead of 256;
//   * all sixteen producer warps run the epilogue (a TMEM lane quarter x a quarter of the columns each).
//
// Why (measured on B200 with tools/micro/gemm_step_bench.cu experiments, profiles/r02_gemm_kloop_experiments.txt): the K
// loop of gemm_tc_ta_kernel was PRODUCER-bound -- fwd2 (16 k-blocks) needs 10.2 us with the MMAs removed, 7.9 us with only
// the MMAs, 5.9 us with only the copies, 11.1 us as a whole; deeper cp.async look-ahead or more lo / TMEM buffers change
// nothing.  tools/micro/mma_rate.cu: a 128 x 128 x 8 TF32 MMA issues at its 64-clock floor, a 128 x 64 x 8 one at 55 clocks
// (floor 32), and a commit + wait per 12 MMAs costs ~20 clocks per MMA.
#pragma once
#include <cuda.h>

#include <cstring>
#include <mutex>
#include <vector>

#include "gemm_tc.cuh"

namespace szb {
namespace tc {

#ifndef SZB_TMA_NBUF
#define SZB_TMA_NBUF 2
#endif
#ifndef SZB_TMA_STAGES
#define SZB_TMA_STAGES 4
#endif

constexpr int kTmaProducerWarps = 16;                         // 0-7: A -> tensor memory, 8-15: lo tile of B
constexpr int kTmaThreads = (kTmaProducerWarps + 2) * 32;     // + warp 16 (MMA issuer) + warp 17 (TMA issuer)

template <int BN, int PASSES>
struct SmemLayoutTma {
    static constexpr int kATile = BM * BK * 4;
    static constexpr int kBTile = BN * BK * 4;
    static constexpr int kStageBytes = kATile + kBTile;
    static constexpr int kNBuf = SZB_TMA_NBUF;                   // lo tiles of B / TMEM buffers of A
    static constexpr int kStages = SZB_TMA_STAGES;               // raw tiles in flight or in use
    static constexpr int kLoBytes = PASSES == 3 ? kNBuf * kBTile : 0;
    static constexpr int kRing = kStages * kStageBytes + kLoBytes;
    static constexpr int kTotal = kRing > 116 * 1024 ? kRing : 116 * 1024;   // > half an SM: one CTA per SM (TMEM, PDL: see gemm_tc.cuh)
    static constexpr int kACols = 32 * (PASSES == 3 ? 2 : 1);
    static constexpr int kTmemCols = (BN + kNBuf * kACols) <= 128 ? 128 : ((BN + kNBuf * kACols) <= 256 ? 256 : 512);
    static_assert(kStages >= kNBuf, "a raw tile stays until its MMAs are done");
    static_assert(BN + kNBuf * kACols <= 512, "tensor memory has 512 columns");
    static_assert(kTotal >= kTmaProducerWarps * kEpiWarpBytes, "the epilogue stages its tiles in the ring");
};

// Wait of a PRODUCER warp: the issuer of the MMAs shares its scheduler with four producer warps, and a warp that spins on
// mbarrier.try_wait takes issue slots from it.  SZB_TMA_BACKOFF (ns) > 0: sleep between polls.  MEASURED: 50 and 200 ns change
// nothing (chain of the step's eight GEMMs 75.9 / 75.8 us against 76.0 us): the polls are not what slows the MMA stream; 0 stays.
#ifndef SZB_TMA_BACKOFF
#define SZB_TMA_BACKOFF 0
#endif
__device__ __forceinline__ void mbar_wait_producer(uint64_t* bar, uint32_t parity) {
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
        if (!done && SZB_TMA_BACKOFF > 0) __nanosleep(SZB_TMA_BACKOFF);
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// box (c0 .. c0 + 31 along K, r0 .. r0 + rows - 1) of a [rows][K] FP32 matrix -> a [rows][128 B] SWIZZLE_128B tile
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int r0, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(r0), "r"(smem_u32(bar))
                 : "memory");
}

// Layer-3 epilogue of a TRAINING step when one tile holds a whole row of logits (n_out <= 128 = BN, the CTA's n0 is 0): softmax
// (lib.rs:1023-1026), delta3 = p - t (lib.rs:1028), the loss of the row (lib.rs:611-615) and the count of rows that survived
// input dropout (lib.rs:607-609, 1047) straight from the accumulator -- what softmax_train_kernel does in a launch of its own.
// All sixteen producer warps take part: a thread owns 32 columns (warp >> 2) of one row (TMEM lane 32 (warp & 3) + lane), the
// four partial maxima and sums of a row meet in shared memory (two barriers of the 512 producer threads), and every thread
// adds them in the same order.  delta3 leaves row-major through the warp's shared-memory tile (C: the A operand of the next
// GEMM) and transposed with lanes along the rows (CT: the B operand of the weight-gradient GEMM).
// (A first version gave a whole row to one thread of warps 0-3: 300 exponentials per thread on one warp per scheduler made
// the step 30 us slower than the separate kernel.)
constexpr int kSoftmaxRedOffset = 80 * 1024;              // [2][4][128] floats behind the sixteen staging tiles (16 x 4608 B)
__device__ __forceinline__ void tma_epilogue_softmax(const GemmArgs& g, uint32_t tmem_d, int m0, bool have_acc, float* stage, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q = warp & 3, cg = warp >> 2;
    const int row = q * 32 + lane, mrow0 = m0 + q * 32, m = m0 + row;
    const int c0 = cg * 32, C = g.N;
    const bool row_ok = m < g.M;
    const bool c_vec = g.C && (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
    uint32_t r[32];
    if (have_acc) {
        tmem_ld32(tmem_d + (uint32_t(q * 32) << 16) + uint32_t(c0), r);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
    }
    float v[32];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        v[j] = c0 + j < C ? __uint_as_float(r[j]) + __ldg(g.bias + c0 + j) : -INFINITY;
        mx = fmaxf(mx, v[j]);                                                    // lib.rs:1023
    }
    red[cg * 128 + row] = mx;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    mx = fmaxf(fmaxf(red[row], red[128 + row]), fmaxf(red[256 + row], red[384 + row]));
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        v[j] = c0 + j < C ? expf(v[j] - mx) : 0.f;                               // lib.rs:1024
        sum += v[j];
    }
    red[512 + cg * 128 + row] = sum;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    sum = ((red[512 + row] + red[640 + row]) + red[768 + row]) + red[896 + row]; // the same order in all four threads of a row
    const bool ok = row_ok && (g.valid ? g.valid[m] != 0 : true);
    const uint32_t label = (row_ok && g.labels) ? g.labels[m] : 0xffffffffu;
    float loss = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int n = c0 + j;
        float d = 0.f;
        if (n < C) {
            const float p = v[j] / sum;                                          // lib.rs:1026
            const float t = g.target_vec ? __ldg(g.target_vec + n) : (uint32_t(n) == label ? 1.f : 0.f);
            d = ok ? p - t : 0.f;                                                // lib.rs:1028; skipped windows contribute nothing
            if (ok && !g.target_vec && uint32_t(n) == label) loss = -logf(fmaxf(p, 1e-12f));
        }
        v[j] = d;
    }
    if (c0 < C) {
        if (g.CT && row_ok) {
            float* tp = g.CT + size_t(c0) * g.ldct + m;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c0 + j < C) tp[size_t(j) * g.ldct] = v[j];
        }
        if (g.C) {
            float* mine = stage + lane * kEpiStride;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(mine + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            __syncwarp();
            const int cc = (lane & 7) * 4, n = c0 + cc;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + (lane >> 3), mm = mrow0 + rr;
                const float4 w = *reinterpret_cast<const float4*>(stage + rr * kEpiStride + cc);
                if (mm < g.M && n < C) {
                    float* dst = g.C + size_t(mm) * g.ldc + n;
                    if (c_vec && n + 4 <= C) {
                        *reinterpret_cast<float4*>(dst) = w;
                    } else {
                        const float e[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (n + t < C) dst[t] = e[t];
                    }
                }
            }
            __syncwarp();
        }
    }
    float cnt = (ok && cg == 0) ? 1.f : 0.f;                                      // one of the four threads of a row counts it
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        loss += __shfl_xor_sync(0xffffffffu, loss, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0 && g.tail) {
        if (cnt > 0.f) atomicAdd(g.tail, cnt);
        if (loss != 0.f) atomicAdd(g.tail + 1, loss);
    }
}

template <int BN, int PASSES, int EPI>
__device__ __forceinline__ void gemm_tma_body(const GemmArgs& g, const CUtensorMap* tmA, const CUtensorMap* tmB, const int bx, const int by,
                                              const int bz) {
    using SL = SmemLayoutTma<BN, PASSES>;
    constexpr int kStages = SL::kStages, kNBuf = SL::kNBuf;
    constexpr int kFullCount = PASSES == 3 ? kTmaProducerWarps : kTmaProducerWarps / 2;     // one arrival per working warp
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ uint64_t s_loaded[kStages], s_full[kStages], s_free[kStages], s_done;
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
    if (tid == 32) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&s_loaded[s], 1);
            mbar_init(&s_full[s], kFullCount);
            mbar_init(&s_free[s], 1);
        }
        mbar_init(&s_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = s_tmem;                      // columns [0, BN): accumulator
    const uint32_t tmem_a = s_tmem + BN;                 // then kNBuf A buffers of kACols columns
    constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
    pdl_launch_dependents();                             // only with this CTA's tensor memory allocated (see gemm_tc_ta_body)
    pdl_wait();
    SZB_TRACE(1);

    if (warp == kTmaProducerWarps + 1) {
        // ------------------------------------------------ TMA issuer -----------------------------------------------
        if (lane == 0) {
            for (int kb = 0; kb < n_kb; ++kb) {
                const int s = kb % kStages;
                if (kb >= kStages) mbar_wait(&s_free[s], uint32_t((kb / kStages - 1) & 1));      // MMAs of k-block kb - kStages are done
                const uint32_t st = smem_base + s * SL::kStageBytes;
#ifdef SZB_X_NOLOAD
                mbar_arrive(&s_loaded[s]);           // experiment: no copies, the tile is "there" at once
#else
                mbar_expect_tx(&s_loaded[s], uint32_t(SL::kStageBytes));
                tma_load_2d(st, tmA, kb0 + kb * BK, m0, &s_loaded[s]);
                tma_load_2d(st + SL::kATile, tmB, kb0 + kb * BK, n0, &s_loaded[s]);
#endif
            }
        }
    } else if (warp == kTmaProducerWarps) {
        // ------------------------------------------------ MMA issuer -----------------------------------------------
        if (lane == 0) {
            for (int kb = 0; kb < n_kb; ++kb) {
                const int s = kb % kStages;
                mbar_wait(&s_full[s], uint32_t((kb / kStages) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t db_hi = make_desc_k_sw128(smem_base + s * SL::kStageBytes + SL::kATile);
                const uint64_t db_lo = make_desc_k_sw128(lo_base + (kb % kNBuf) * SL::kBTile);
                const uint32_t a_hi = tmem_a + uint32_t((kb % kNBuf) * SL::kACols), a_lo = a_hi + 32;
#pragma unroll
                for (int k = 0; k < BK / UK; ++k) {
                    const uint64_t adv = uint64_t((k * UK * 4) >> 4);
                    const uint32_t acol = uint32_t(k * UK);
                    const uint32_t acc0 = (kb > 0 || k > 0) ? 1u : 0u;
#ifdef SZB_X_NOMMA
                    if (kb > 0) continue;          // experiment: one k-block of MMAs (the accumulator is defined), then bare commits
#endif
                    if (PASSES == 3) {
                        umma_tf32_ts(tmem_d, a_lo + acol, db_hi + adv, idesc, acc0);
                        umma_tf32_ts(tmem_d, a_hi + acol, db_lo + adv, idesc, 1u);
                        umma_tf32_ts(tmem_d, a_hi + acol, db_hi + adv, idesc, 1u);
                    } else {
                        umma_tf32_ts(tmem_d, a_hi + acol, db_hi + adv, idesc, acc0);
                    }
                }
                umma_commit(&s_free[s]);
            }
            // End of the tile on a barrier of its own: warps that do not walk the K loop (8-15 in single-pass mode) cannot wait
            // for "completion number n_kb / kStages" of a ring barrier by parity -- an mbarrier still in its first phase
            // answers a wait for the odd parity at once.
            if (n_kb > 0) umma_commit(&s_done);
        }
    } else if (warp < 8) {
        // ------------------------------------- A slab -> tensor memory (warps 0-7) --------------------------------
        // thread <-> row 32 (warp & 3) + lane of the tile (its TMEM lane), 16-byte chunks [4 h, 4 h + 4) of the k-block
        const int q = warp & 3, h = warp >> 2;
        const uint32_t a_lane = uint32_t(q * 32) << 16;
        for (int kb = 0; kb < n_kb; ++kb) {
            const int s = kb % kStages;
            if (kb >= kNBuf) {          // MMAs of k-block kb - kNBuf done: TMEM A buffer kb % kNBuf is free
                mbar_wait_producer(&s_free[(kb - kNBuf) % kStages], uint32_t(((kb - kNBuf) / kStages) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            mbar_wait_producer(&s_loaded[s], uint32_t((kb / kStages) & 1));
            const uint32_t st = smem_base + s * SL::kStageBytes;
#ifndef SZB_X_NOSPLIT
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
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#else
            (void)st; (void)a_lane; (void)h;
#endif
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_full[s]);
            if (kb == 0) SZB_TRACE(2);
        }
        SZB_TRACE(3);
    } else if (PASSES == 3) {
        // --------------------------------------- lo tile of B (warps 8-15) ----------------------------------------
        // thread <-> 16-byte chunk (t & 7) of rows (t >> 3) + 32 u; the raw tile itself is the hi operand
        const int t = tid - 256, lr = t >> 3, lc = t & 7;
        constexpr int kRb = BN / 32;
        for (int kb = 0; kb < n_kb; ++kb) {
            const int s = kb % kStages;
            if (kb >= kNBuf) mbar_wait_producer(&s_free[(kb - kNBuf) % kStages], uint32_t(((kb - kNBuf) / kStages) & 1));   // lo buffer kb % kNBuf is free
            mbar_wait_producer(&s_loaded[s], uint32_t((kb / kStages) & 1));
            const uint32_t src = smem_base + s * SL::kStageBytes + SL::kATile;
            const uint32_t dst = lo_base + (kb % kNBuf) * SL::kBTile;
#ifndef SZB_X_NOSPLIT
            float4 v[kRb];
#pragma unroll
            for (int u = 0; u < kRb; ++u)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                             : "r"(src + sw128_off(lr + 32 * u, lc)));
#pragma unroll
            for (int u = 0; u < kRb; ++u)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + sw128_off(lr + 32 * u, lc)), "f"(v[u].x - tf32_trunc(v[u].x)),
                             "f"(v[u].y - tf32_trunc(v[u].y)), "f"(v[u].z - tf32_trunc(v[u].z)), "f"(v[u].w - tf32_trunc(v[u].w))
                             : "memory");
#else
            (void)src; (void)dst; (void)lr; (void)lc;
#endif
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_full[s]);
        }
    }
    if (warp < kTmaProducerWarps) {
        if (n_kb > 0) mbar_wait_producer(&s_done, 0u);                                       // every MMA of the tile has completed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        SZB_TRACE(4);
        // every MMA has completed, every bulk copy was consumed by one: the ring is free to stage the output tile.
        // A warp takes its TMEM lane quarter (warp & 3) x 32 columns (warp >> 2).
        constexpr int kEpiWarps = (BN / 32) * 4;
        if (EPI == TC_SOFTMAX_CE)        // the whole row of logits sits in this tile (launcher: N <= BN = 128, one column of tiles)
            tma_epilogue_softmax(g, tmem_d, m0, n_kb > 0, reinterpret_cast<float*>(tc_smem + warp * kEpiWarpBytes),
                                 reinterpret_cast<float*>(tc_smem + kSoftmaxRedOffset));
        else if (warp < kEpiWarps)
            tc_epilogue_staged<32, EPI>(g, tmem_d + uint32_t((warp >> 2) * 32), m0, n0 + (warp >> 2) * 32, n_kb > 0,
                                        reinterpret_cast<float*>(tc_smem + warp * kEpiWarpBytes));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    SZB_TRACE(5);
    __syncthreads();
    SZB_TRACE(6);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(SL::kTmemCols) : "memory");
}

template <int BN, int PASSES, int EPI>
__global__ void __launch_bounds__(kTmaThreads) gemm_tma_kernel(const GemmArgs g, const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB) {
    gemm_tma_body<BN, PASSES, EPI>(g, &tmA, &tmB, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Optional rider of the grouped launch below.  Multi-GPU, one-shot gradient exchange (mlp.cu: p2p_exchange): the CTA that
// completes a gradient tile -- the last of its K ranges to have added its partial sums -- stores the finished tile into this
// rank's slot on every peer right away, while the other tiles are still being computed.  The update kernel then has nothing
// left to send but the [n_used, loss] tail, and its system-scope fence no longer waits for 753 KB of posted NVLink stores per
// peer.  MEASURED at N = 2 (szb_ctx::p2p_early_push): 102.5 us per step against 100.5 us without it -- the exchange is not
// bound by when the data leaves; off by default.
constexpr int kMaxPushPeers = 16;
constexpr int kMaxPushTiles = 256;
struct PeerPush {
    float* area[kMaxPushPeers];          // exchange area of every rank (this rank's own entry is not written)
    unsigned long long slot_off;         // floats: ((step & 1) * world + rank) * cap -- this rank's slot of the step
    const float* G;                      // the private gradient vector the products accumulate into
    unsigned int* counters;              // [kMaxPushTiles] tickets, zero between launches
    int world, rank;                     // world <= 1: no push
};
// Several independent products in one launch (see gemm_tc_ta_group_kernel): CTA b serves problem p with first[p] <= b < first[p + 1].
struct GroupArgsTma {
    GemmArgs g[kMaxGroup];
    CUtensorMap tmA[kMaxGroup], tmB[kMaxGroup];
    int first[kMaxGroup + 1];
    int tiles_n[kMaxGroup], tiles_m[kMaxGroup];
    int nz[kMaxGroup], tile_first[kMaxGroup];       // K ranges per tile; index of the problem's first tile ticket
    int count;
    PeerPush push;
};
template <int BN, int PASSES, int EPI>
__global__ void __launch_bounds__(kTmaThreads) gemm_tma_group_kernel(const __grid_constant__ GroupArgsTma ga) {
    const int b = blockIdx.x;
    int p = 0;
#pragma unroll
    for (int i = 1; i < kMaxGroup; ++i)
        if (i < ga.count && b >= ga.first[i]) p = i;
    const int local = b - ga.first[p];
    const int tn = ga.tiles_n[p], tm = ga.tiles_m[p];
    const int bx = local % tn, by = (local / tn) % tm;
    gemm_tma_body<BN, PASSES, EPI>(ga.g[p], &ga.tmA[p], &ga.tmB[p], bx, by, local / (tn * tm));
    if (ga.push.world > 1) {
        // every thread of the CTA has issued its reductions (the body ends with a block barrier).  Release pattern as in
        // p2p_publish: barrier -> fence -> ticket; the last K range of the tile then reads the finished sums from L2.
        __shared__ int s_last;
        const int tile = ga.tile_first[p] + by * tn + bx;
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int ticket = atomicAdd(ga.push.counters + tile, 1u);
            s_last = ticket == unsigned(ga.nz[p]) - 1u;
            if (s_last) ga.push.counters[tile] = 0u;               // the next step finds it zero
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            const GemmArgs& g = ga.g[p];
            const int m0 = by * BM, n0 = bx * BN;
            const int rows = min(BM, g.M - m0), nc4 = min(BN, g.N - n0) >> 2;      // launcher: N % 4 == 0, 16-byte aligned rows
            const size_t base = size_t(g.C - ga.push.G) + size_t(m0) * g.ldc + n0;
            for (int e = threadIdx.x; e < rows * nc4; e += blockDim.x) {
                const int r = e / nc4, c4 = e - r * nc4;
                const size_t idx = base + size_t(r) * g.ldc + 4 * c4;
                const float4 v = __ldcg(reinterpret_cast<const float4*>(ga.push.G + idx));
#pragma unroll
                for (int q = 0; q < kMaxPushPeers; ++q)
                    if (q < ga.push.world && q != ga.push.rank) *reinterpret_cast<float4*>(ga.push.area[q] + ga.push.slot_off + idx) = v;
            }
        }
    }
}

// ---- host side: tensor maps ---------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled comes from the driver (no link-time dependency on libcuda: cudaGetDriverEntryPoint).  A map
// depends only on (base, rows, cols, row stride, box rows); the buffers of a net are stable, so the few maps a step needs are
// encoded once and found again by a linear search.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

struct TensorMapKey {
    const void* base; int rows, cols, ld, box_rows;
    bool operator==(const TensorMapKey& o) const { return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows; }
};
// [rows][cols] FP32, row stride ld floats (16-byte multiple), boxes of box_rows x 32 floats, SWIZZLE_128B.  false: not encodable.
inline bool tensor_map_for(const float* base, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
    static std::mutex mu;
    static std::vector<std::pair<TensorMapKey, CUtensorMap>> cache;
    const TensorMapKey key{base, rows, cols, ld, box_rows};
    std::lock_guard<std::mutex> lock(mu);
    for (const auto& e : cache)
        if (e.first == key) { *out = e.second; return true; }
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
    const cuuint64_t strides[1] = {cuuint64_t(ld) * 4};
    const cuuint32_t box[2] = {cuuint32_t(BK), cuuint32_t(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMap tm;
    if (fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (cache.size() >= 256) cache.erase(cache.begin(), cache.begin() + 128);
    cache.emplace_back(key, tm);
    *out = tm;
    return true;
}

// *done = false (nothing launched): operands not 16-byte aligned, TMA switched off, or no driver entry point -- the caller
// falls back to launch_gemm_tc.
template <int BN, int PASSES, int EPI>
szb_status launch_gemm_tma(szb_ctx* ctx, GemmArgs g, int split_k, bool* done) {
    *done = false;
    if (g.M <= 0 || g.N <= 0) { *done = true; return SZB_OK; }
    if (!ctx->gemm_tma || g.K <= 0 || !gemm_operands_aligned(g)) return SZB_OK;
    CUtensorMap tmA, tmB;
    if (!tensor_map_for(g.A, g.M, g.K, g.lda, BM, &tmA) || !tensor_map_for(g.B, g.N, g.K, g.ldb, BN, &tmB)) return SZB_OK;
    using SL = SmemLayoutTma<BN, PASSES>;
    const int kb_total = (g.K + BK - 1) / BK;
    split_k = EPI == TC_ATOMIC ? std::max(1, std::min(split_k, kb_total)) : 1;
    g.k_chunk = ((kb_total + split_k - 1) / split_k) * BK;
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, (g.K + g.k_chunk - 1) / g.k_chunk);
    static bool attr_set[64] = {};
    if (!attr_set[ctx->device & 63]) {
        SZB_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<BN, PASSES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::kTotal));
        attr_set[ctx->device & 63] = true;
    }
    SZB_CUDA(launch_pdl(ctx, gemm_tma_kernel<BN, PASSES, EPI>, grid, dim3(kTmaThreads), size_t(SL::kTotal), g, tmA, tmB));
    SZB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    *done = true;
    return SZB_OK;
}

// Layer 3 of a training step with softmax / cross-entropy in the epilogue (tma_epilogue_softmax).  Needs the whole row of
// logits in one 128-column tile; *done = false (nothing launched) otherwise: the caller then runs the plain bias epilogue
// followed by softmax_train_kernel.
template <int PASSES>
szb_status launch_gemm_tma_softmax(szb_ctx* ctx, GemmArgs g, bool* done) {
    constexpr int BN = 128;
    *done = false;
    if (g.M <= 0 || g.N <= 0 || g.N > BN || g.K <= 0 || !ctx->gemm_tma || !gemm_operands_aligned(g)) return SZB_OK;
    CUtensorMap tmA, tmB;
    if (!tensor_map_for(g.A, g.M, g.K, g.lda, BM, &tmA) || !tensor_map_for(g.B, g.N, g.K, g.ldb, BN, &tmB)) return SZB_OK;
    using SL = SmemLayoutTma<BN, PASSES>;
    static_assert(SL::kTotal >= kSoftmaxRedOffset + 2 * 4 * 128 * 4, "room for the partial maxima and sums");
    g.k_chunk = ((g.K + BK - 1) / BK) * BK;
    static bool attr_set[64] = {};
    if (!attr_set[ctx->device & 63]) {
        SZB_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<BN, PASSES, TC_SOFTMAX_CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::kTotal));
        attr_set[ctx->device & 63] = true;
    }
    SZB_CUDA(launch_pdl(ctx, gemm_tma_kernel<BN, PASSES, TC_SOFTMAX_CE>, dim3(1, (g.M + BM - 1) / BM, 1), dim3(kTmaThreads), size_t(SL::kTotal), g, tmA, tmB));
    SZB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    *done = true;
    return SZB_OK;
}

// push (optional): early peer push of the finished tiles; *pushed tells whether the launch does it (needs 16-byte aligned rows
// of every C and N % 4 == 0, and few enough tiles for the ticket array).
template <int PASSES>
szb_status launch_gemm_tma_group(szb_ctx* ctx, const GemmArgs* gs, int count, bool* done, const PeerPush* push = nullptr, bool* pushed = nullptr) {
    if (pushed) *pushed = false;
    constexpr int BN = 128;
    *done = false;
    if (count < 1 || count > kMaxGroup || !ctx->gemm_tma) return SZB_OK;
    GroupArgsTma ga{};
    int total_tiles = 0;
    for (int p = 0; p < count; ++p) {
        if (gs[p].M <= 0 || gs[p].N <= 0 || gs[p].K <= 0 || !gemm_operands_aligned(gs[p])) return SZB_OK;
        if (!tensor_map_for(gs[p].A, gs[p].M, gs[p].K, gs[p].lda, BM, &ga.tmA[p]) || !tensor_map_for(gs[p].B, gs[p].N, gs[p].K, gs[p].ldb, BN, &ga.tmB[p]))
            return SZB_OK;
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
        ga.nz[p] = nz;
        ctas += ga.tiles_n[p] * ga.tiles_m[p] * nz;
    }
    for (int p = count; p <= kMaxGroup; ++p) ga.first[p] = ctas;
    ga.count = count;
    ga.push.world = 1;
    if (push && push->world > 1 && push->world <= kMaxPushPeers) {
        bool ok = true;
        int tiles = 0;
        for (int p = 0; p < count; ++p) {
            ga.tile_first[p] = tiles;
            tiles += ga.tiles_n[p] * ga.tiles_m[p];
            const size_t off = size_t(gs[p].C - push->G);
            ok = ok && gs[p].N % 4 == 0 && gs[p].ldc % 4 == 0 && off % 4 == 0 && (reinterpret_cast<uintptr_t>(push->G) & 15) == 0;
        }
        if (ok && tiles <= kMaxPushTiles && (push->slot_off % 4) == 0) {
            ga.push = *push;
            if (pushed) *pushed = true;
        }
    }
    using SL = SmemLayoutTma<BN, PASSES>;
    static bool attr_set[64] = {};
    if (!attr_set[ctx->device & 63]) {
        SZB_CUDA(cudaFuncSetAttribute(gemm_tma_group_kernel<BN, PASSES, TC_ATOMIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::kTotal));
        attr_set[ctx->device & 63] = true;
    }
    SZB_CUDA(launch_pdl(ctx, gemm_tma_group_kernel<BN, PASSES, TC_ATOMIC>, dim3(ctas), dim3(kTmaThreads), size_t(SL::kTotal), ga));
    SZB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    *done = true;
    return SZB_OK;
}

}  // namespace tc
}  // namespace szb
