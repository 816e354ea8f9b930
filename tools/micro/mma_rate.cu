// Microbenchmark: issue / completion rate of back-to-back tcgen05.mma kind::tf32 instructions from ONE thread, by N (64 / 128 /
// 256), operand source of A (shared memory "SS" or tensor memory "TS") and commit granularity (one commit per 12 MMAs, as the
// K loop of gemm_tc_ta_kernel issues them, or one at the very end).  Operand contents are irrelevant (zero-filled shared
// memory); the accumulator is never read.  Answers: does a 128 x 64 x 8 MMA reach its 32-clock floor, and what does the
// per-k-block commit cost?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I streamz_b200/csrc tools/micro/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gemm_tc.cuh"

namespace szb {
void set_error(const char*, ...) {}
}
using namespace szb::tc;

template <int BN, bool TS>
__global__ void __launch_bounds__(128) rate_kernel(int rounds, int per_commit, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + BN * 128) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = s_tmem, tmem_a = s_tmem + BN;
    constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
    if (tid == 0) {
        const uint32_t base = smem_u32(smem);
        const uint64_t da = make_desc_k_sw128(base), db = make_desc_k_sw128(base + 16384);
        uint32_t phase = 0;
        const long long t0 = clock64();
        int since = 0;
        for (int r = 0; r < rounds; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t adv = uint64_t((k * 8 * 4) >> 4);
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    if (TS) umma_tf32_ts(tmem_d, tmem_a + uint32_t(k * 8 + (p & 1) * 32), db + adv, idesc, 1u);
                    else umma_tf32(tmem_d, da + adv, db + adv, idesc, 1u);
                }
            }
            since += 12;
            if (per_commit > 0 && since >= per_commit) {
                umma_commit(&bar);
                mbar_wait(&bar, phase);
                phase ^= 1u;
                since = 0;
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, phase);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(512) : "memory");
}

// The K loop of gemm_tma_kernel as its issuer thread runs it: descriptors recomputed per k-block from a ring stage, alternating
// A / lo buffers, the accumulate flag computed per MMA, one commit per k-block to a ring barrier that nobody waits for.
template <int BN>
__global__ void __launch_bounds__(128) loop_kernel(int rounds, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar[4], done;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int kStage = 16384 + BN * 128, kBTile = BN * 128;
    for (int i = tid; i < (4 * kStage + 2 * kBTile) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < 4; ++s) mbar_init(&bar[s], 1);
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = s_tmem, tmem_a = s_tmem + BN;
    constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
    if (tid == 0) {
        const uint32_t smem_base = smem_u32(smem), lo_base = smem_base + 4 * kStage;
        const long long t0 = clock64();
        for (int kb = 0; kb < rounds; ++kb) {
            const int s = kb % 4;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t db_hi = make_desc_k_sw128(smem_base + s * kStage + 16384);
            const uint64_t db_lo = make_desc_k_sw128(lo_base + (kb % 2) * kBTile);
            const uint32_t a_hi = tmem_a + uint32_t((kb % 2) * 64), a_lo = a_hi + 32;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t adv = uint64_t((k * 8 * 4) >> 4);
                const uint32_t acol = uint32_t(k * 8);
                const uint32_t acc0 = (kb > 0 || k > 0) ? 1u : 0u;
                umma_tf32_ts(tmem_d, a_lo + acol, db_hi + adv, idesc, acc0);
                umma_tf32_ts(tmem_d, a_hi + acol, db_lo + adv, idesc, 1u);
                umma_tf32_ts(tmem_d, a_hi + acol, db_hi + adv, idesc, 1u);
            }
            umma_commit(&bar[s]);
        }
        umma_commit(&done);
        mbar_wait(&done, 0u);
        out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(512) : "memory");
}
template <int BN>
static void run_loop(int ctas, long long* d_out) {
    const int rounds = 200;
    const size_t smem = 4 * (16384 + BN * 128) + 2 * BN * 128 + 1024;
    cudaFuncSetAttribute(loop_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    loop_kernel<BN><<<ctas, 128, smem>>>(rounds, d_out);
    loop_kernel<BN><<<ctas, 128, smem>>>(rounds, d_out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("loop: failed: %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
    std::vector<long long> h(ctas);
    cudaMemcpy(h.data(), d_out, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (long long v : h) mean += double(v);
    mean /= ctas;
    printf("K loop as issued by gemm_tma_kernel (ring descriptors, a commit per k-block, no waits) N=%3d ctas %3d: %7.1f clk per MMA\n", BN, ctas,
           mean / (rounds * 12.0));
}

template <int BN, bool TS>
static void run(const char* name, int ctas, int per_commit, long long* d_out) {
    const int rounds = 200;
    const size_t smem = 16384 + BN * 128 + 1024;
    cudaFuncSetAttribute(rate_kernel<BN, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    rate_kernel<BN, TS><<<ctas, 128, smem>>>(rounds, per_commit, d_out);
    rate_kernel<BN, TS><<<ctas, 128, smem>>>(rounds, per_commit, d_out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); exit(1); }
    std::vector<long long> h(ctas);
    cudaMemcpy(h.data(), d_out, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (long long v : h) mean += double(v);
    mean /= ctas;
    printf("%-4s N=%3d ctas %3d commit+wait every %3d MMAs: %7.1f clk per MMA (floor %d)\n", name, BN, ctas, per_commit, mean / (rounds * 12.0), BN / 2);
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 1024 * sizeof(long long));
    for (int ctas : {1, 148}) { run_loop<64>(ctas, d_out); run_loop<128>(ctas, d_out); }
    for (int ctas : {1}) {
        for (int pc : {0, 12, 48}) {
            run<64, true>("TS", ctas, pc, d_out);
            run<128, true>("TS", ctas, pc, d_out);
            run<256, true>("TS", ctas, pc, d_out);
            run<64, false>("SS", ctas, pc, d_out);
            run<128, false>("SS", ctas, pc, d_out);
            run<256, false>("SS", ctas, pc, d_out);
        }
    }
    return 0;
}
