// Semantics probe for 2-CTA tcgen05 MMAs (cta_group::2) -- the next lever on the dense-layer GEMMs after the A operand moved
// to tensor memory: a CTA pair shares the B tile, so each CTA stages and the tensor core reads only
// half of it.
//
// STATUS: WRITTEN WITHOUT GPU ACCESS at the end of round 1 (the GPU budget was spent); it compiles for sm_100a and has NOT
// been run.  It is a stand-alone experiment, not part of the library.  Assumptions it tests (from the CUTLASS / DeepGEMM
// usage quoted in the programming guide):
//   * a cluster of two CTAs; each CTA's warp 0 executes tcgen05.alloc.cta_group::2 and gets the same TMEM base;
//   * D is 256 x N: CTA r owns rows [128 r, 128 r + 128) of A (its own [128][32] K-major SWIZZLE_128B tile) and of D (its own
//     TMEM), and stages rows [N/2 r, N/2 r + N/2) of B at the SAME shared-memory offset as its peer;
//   * only the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 with its own descriptors and idesc M = 256;
//   * tcgen05.commit.cta_group::2 ... multicast::cluster with mask 0b11 arrives on the mbarrier at the same offset in both CTAs.
// Fully serialised per k-block (stage -> cluster barrier -> MMAs -> commit -> both wait): only the semantics are probed.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/micro/umma_2cta_test.cu -o tools/micro/umma_2cta_test
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

constexpr int BM = 128, BN = 128, BK = 32, UK = 8;       // per-CTA rows, pair-wide N, K per stage, K per MMA

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return uint32_t(r) * 128u + (uint32_t(c ^ (r & 7)) << 4); }
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr >> 4) & 0x3FFF);
    d |= uint64_t(1) << 16;
    d |= uint64_t(64) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) umma_2cta_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                                  float* __restrict__ D, int K) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    unsigned char* a_tile = smem;                          // [128][32] floats
    unsigned char* b_tile = smem + BM * BK * 4;            // [BN/2][32] floats: this CTA's half of B

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync();                                        // both CTAs alive, barriers initialised
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = s_tmem;
    constexpr uint32_t idesc = make_idesc_tf32(2 * BM, BN);

    const int n_kb = K / BK;
    for (int kb = 0; kb < n_kb; ++kb) {
        const int k0 = kb * BK;
        for (int q = tid; q < BM * 8; q += 128) {          // A rows of this CTA
            const int r = q >> 3, c = q & 7;
            const float4 v = *reinterpret_cast<const float4*>(A + size_t(rank * BM + r) * K + k0 + c * 4);
            *reinterpret_cast<float4*>(a_tile + sw128_off(r, c)) = v;
        }
        for (int q = tid; q < (BN / 2) * 8; q += 128) {    // this CTA's half of the B rows
            const int r = q >> 3, c = q & 7;
            const float4 v = *reinterpret_cast<const float4*>(B + size_t(rank * (BN / 2) + r) * K + k0 + c * 4);
            *reinterpret_cast<float4*>(b_tile + sw128_off(r, c)) = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        cluster_sync();                                    // both CTAs' operands are in place
        if (rank == 0 && tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t da = make_desc_k_sw128(smem_u32(a_tile)), db = make_desc_k_sw128(smem_u32(b_tile));
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
                const uint64_t adv = uint64_t((k * UK * 4) >> 4);
                const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem_d), "l"(da + adv), "l"(db + adv), "r"(idesc), "r"(acc)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)),
                         "h"(uint16_t(3))
                         : "memory");
        }
        {   // both CTAs wait for the MMAs of this k-block on their own barrier
            const uint32_t addr = smem_u32(&s_bar), parity = uint32_t(kb & 1);
            uint32_t done = 0;
            for (uint32_t spin = 0; !done; ++spin) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(addr), "r"(parity)
                    : "memory");
                if (spin > (1u << 24)) __trap();           // never hang the GPU on a protocol mistake
            }
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        cluster_sync();                                    // nobody refills a tile the pair's MMAs might still read
    }
    // epilogue: thread t <-> row rank * 128 + t (TMEM lane t of this CTA)
    const uint32_t lane_addr = tmem_d + (uint32_t(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(lane_addr + uint32_t(c0)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[size_t(rank * BM + tid) * BN + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync();                                        // both CTAs are done with the pair's tensor memory
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(BN) : "memory");
}

int main() {
    const int M = 256, N = BN, K = 64;
    std::vector<float> hA(size_t(M) * K), hB(size_t(N) * K);
    unsigned s = 777u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return float((s >> 8) & 0xFFFF) / 32768.f - 1.f; };
    for (auto& v : hA) v = rnd();
    for (auto& v : hB) v = rnd();
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, hA.size() * 4));
    CK(cudaMalloc(&dB, hB.size() * 4));
    CK(cudaMalloc(&dD, size_t(M) * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, size_t(M) * N * 4));
    const int smem = BM * BK * 4 + (BN / 2) * BK * 4;
    CK(cudaFuncSetAttribute(umma_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_2cta_kernel<<<2, 128, smem>>>(dA, dB, dD, K);     // one cluster of two CTAs (__cluster_dims__)
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> hD(size_t(M) * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double err = 0, mag = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += double(hA[size_t(m) * K + k]) * double(hB[size_t(n) * K + k]);
            err = std::fmax(err, std::fabs(acc - double(hD[size_t(m) * N + n])));
            mag = std::fmax(mag, std::fabs(acc));
        }
    printf("2-CTA tcgen05.mma (M = 256 across the pair, N = %d, K = %d, TF32): max |err| / max |ref| = %.3e  -> %s\n", N, K, err / mag,
           err / mag < 5e-3 ? "OK" : "MISMATCH (operand split / descriptor assumptions wrong)");
    return err / mag < 5e-3 ? 0 : 1;
}
