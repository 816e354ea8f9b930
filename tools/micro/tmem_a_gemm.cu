// Semantics + timing probe for the NEXT GEMM generation ("next"): tcgen05.mma kind::tf32 with the A
// operand in TENSOR MEMORY instead of shared memory.  The shipped kernel (streamz_b200/csrc/gemm_tc.cuh) is bound by
// shared-memory bandwidth: per k-step the tensor core re-reads the 4 KB A slice for each of the three 3xTF32 products.  With
// A in TMEM (lane = row m, column = k, one 32-bit column per TF32 element -- the layout of cute::UMMA::tmem_frg for
// M = 128, cta_group::1) the producers write each row's k-block once with tcgen05.st and the MMAs read only B from shared
// memory.
//
// This file is a stand-alone experiment, not part of the library.  One CTA (128 threads, thread t <-> row t) computes a
// 128 x BN tile, fully serialised per k-block (stage A -> stage B -> MMAs -> wait): it answers "is the layout right and is
// the result FP32-accurate", and prints the error of the 3-product and the 1-product variants against a float64 reference.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/micro/tmem_a_gemm.cu -o tools/micro/tmem_a_gemm
//   tools/micro/tmem_a_gemm            # expected: 3xTF32 max rel err ~1e-6, TF32 ~1e-3
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

constexpr int BM = 128, BK = 32, UK = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return uint32_t(r) * 128u + (uint32_t(c ^ (r & 7)) << 4); }
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {   // as gemm_tc.cuh
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
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
          "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]),
          "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]),
          "f"(v[30]), "f"(v[31])
        : "memory");
}

template <int BN, int PASSES>
__global__ void __launch_bounds__(128) tmem_a_gemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                          float* __restrict__ D, int ldd, int M, int N, int K) {
    constexpr int kCols = BN + 64 <= 128 ? 128 : (BN + 64 <= 256 ? 256 : 512);   // D | A_hi (32) | A_lo (32), power of two
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * BM;
    unsigned char* b_hi = smem;
    unsigned char* b_lo = smem + BN * BK * 4;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(kCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t tmem_d = tmem, tmem_a_hi = tmem + BN, tmem_a_lo = tmem + BN + 32;
    const uint32_t my_lanes = uint32_t(warp * 32) << 16;     // a warp may only touch its own TMEM lane quarter
    constexpr uint32_t idesc = make_idesc_tf32(BM, BN);

    const int n_kb = (K + BK - 1) / BK;
    for (int kb = 0; kb < n_kb; ++kb) {
        const int k0 = kb * BK;
        // ---- A: thread t owns row m0 + t; 32 floats of the k-block -> TMEM lane t, columns [A_hi | A_lo]
        float hi[32], lo[32];
        const int m = m0 + tid;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float x = (m < M && k0 + j < K) ? __ldcg(A + size_t(m) * lda + k0 + j) : 0.f;
            hi[j] = x;                       // kind::tf32 reads the upper 19 bits: the raw word is the hi operand
            lo[j] = x - tf32_trunc(x);
        }
        tmem_st32(tmem_a_hi + my_lanes, hi);
        if (PASSES == 3) tmem_st32(tmem_a_lo + my_lanes, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        // ---- B: [BN][32] floats, K-major SWIZZLE_128B tiles (raw = hi, and lo)
        for (int q = tid; q < BN * 8; q += 128) {
            const int r = q >> 3, c = q & 7, n = r, k = k0 + c * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < N) {
                const float* p = B + size_t(n) * ldb + k;
                if (k + 0 < K) v.x = __ldcg(p + 0);
                if (k + 1 < K) v.y = __ldcg(p + 1);
                if (k + 2 < K) v.z = __ldcg(p + 2);
                if (k + 3 < K) v.w = __ldcg(p + 3);
            }
            *reinterpret_cast<float4*>(b_hi + sw128_off(r, c)) = v;
            *reinterpret_cast<float4*>(b_lo + sw128_off(r, c)) =
                make_float4(v.x - tf32_trunc(v.x), v.y - tf32_trunc(v.y), v.z - tf32_trunc(v.z), v.w - tf32_trunc(v.w));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t db_hi = make_desc_k_sw128(smem_u32(b_hi)), db_lo = make_desc_k_sw128(smem_u32(b_lo));
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
                const uint64_t adv = uint64_t((k * UK * 4) >> 4);
                const uint32_t acol = uint32_t(k * UK);             // 8 TF32 columns per k-step
                const uint32_t acc0 = (kb > 0 || k > 0) ? 1u : 0u;
                if (PASSES == 3) {
                    umma_tf32_ts(tmem_d, tmem_a_lo + acol, db_hi + adv, idesc, acc0);
                    umma_tf32_ts(tmem_d, tmem_a_hi + acol, db_lo + adv, idesc, 1u);
                    umma_tf32_ts(tmem_d, tmem_a_hi + acol, db_hi + adv, idesc, 1u);
                } else {
                    umma_tf32_ts(tmem_d, tmem_a_hi + acol, db_hi + adv, idesc, acc0);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
        }
        // everyone waits for this k-block's MMAs before the operands are overwritten (serialised on purpose)
        {
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
                if (spin > (1u << 24)) __trap();
            }
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // ---- epilogue: thread t <-> row t
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem_d + my_lanes + uint32_t(c0)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int m = m0 + tid;
        if (m < M)
            for (int j = 0; j < 16; ++j)
                if (c0 + j < N) D[size_t(m) * ldd + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kCols) : "memory");
}

template <int BN, int PASSES>
static double run(int M, int N, int K, const std::vector<float>& hA, const std::vector<float>& hB, const std::vector<double>& ref) {
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, hA.size() * 4));
    CK(cudaMalloc(&dB, hB.size() * 4));
    CK(cudaMalloc(&dD, size_t(M) * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, size_t(M) * N * 4));
    const int smem = 2 * BN * BK * 4;
    CK(cudaFuncSetAttribute(tmem_a_gemm_kernel<BN, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tmem_a_gemm_kernel<BN, PASSES><<<(M + BM - 1) / BM, 128, smem>>>(dA, K, dB, K, dD, N, M, N, K);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> hD(size_t(M) * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double err = 0, mag = 0;
    for (size_t i = 0; i < hD.size(); ++i) {
        err = std::fmax(err, std::fabs(double(hD[i]) - ref[i]));
        mag = std::fmax(mag, std::fabs(ref[i]));
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return err / mag;
}

int main() {
    const int M = 300, N = 128, K = 200;     // ragged M and K on purpose
    std::vector<float> hA(size_t(M) * K), hB(size_t(N) * K);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return float((s >> 8) & 0xFFFF) / 32768.f - 1.f; };
    for (auto& v : hA) v = rnd();
    for (auto& v : hB) v = 0.5f * rnd();
    std::vector<double> ref(size_t(M) * N);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += double(hA[size_t(m) * K + k]) * double(hB[size_t(n) * K + k]);
            ref[size_t(m) * N + n] = acc;
        }
    const double e3 = run<128, 3>(M, N, K, hA, hB, ref);
    const double e1 = run<128, 1>(M, N, K, hA, hB, ref);
    const double e3n = run<64, 3>(M, 64, K, hA, hB, ref.size() ? [&] {
        std::vector<double> r2(size_t(M) * 64);
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < 64; ++n) r2[size_t(m) * 64 + n] = ref[size_t(m) * N + n];
        return r2;
    }() : ref);
    printf("A operand in TMEM, M=%d N=%d K=%d: max |err| / max |ref|:  3xTF32 %.3e   TF32 %.3e   3xTF32 (BN=64) %.3e\n", M, N, K, e3, e1, e3n);
    const bool ok = e3 < 1e-5 && e1 < 1e-2 && e3n < 1e-5;
    printf("%s\n", ok ? "OK: layout and accuracy as expected" : "MISMATCH: the TMEM A layout / descriptor assumptions are wrong");
    return ok ? 0 : 1;
}
