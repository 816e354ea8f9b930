// Microtest: tcgen05.mma kind::i8 (s8/u8 operands, s32 accumulator in TMEM) with K-major SWIZZLE_128B operands, the
// building block of an exactly reproducible integer polyphase resampler.  One CTA, M = 128, N = 64, K = 64 (two K-steps of
// 32), all four signedness combinations accumulated the way the resampler would (hi*hi | hi*lo + lo*hi | lo*lo).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8_mma_test i8_mma_test.cu && ./i8_mma_test
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr >> 4) & 0x3FFF);
    d |= uint64_t(1) << 16;
    d |= uint64_t(64) << 32;     // SBO = 1024 B between 8-row groups
    d |= uint64_t(1) << 46;      // descriptor version (Blackwell)
    d |= uint64_t(2) << 61;      // SWIZZLE_128B
    return d;
}
// instruction descriptor for kind::i8: D = S32 (bits 4-5 = 2), a/b format 0 = u8, 1 = s8 (bits 7-9 / 10-12), K-major both
__host__ __device__ constexpr uint32_t make_idesc_i8(int m, int n, int a_signed, int b_signed) {
    return (2u << 4) | (uint32_t(a_signed) << 7) | (uint32_t(b_signed) << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return uint32_t(r) * 128u + (uint32_t(c ^ (r & 7)) << 4); }

constexpr int M = 128, N = 64, K = 64;

// A planes [M][K] bytes (hi signed, lo unsigned), B planes [N][K] (hi signed, lo unsigned); out: three accumulators [M][N]
__global__ void __launch_bounds__(128) k(const uint8_t* a_hi, const uint8_t* a_lo, const uint8_t* b_hi, const uint8_t* b_lo, int32_t* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    unsigned char* sa_hi = smem;                 // [128 rows][128 B] (K = 64 used of the 128-byte row)
    unsigned char* sa_lo = smem + 16384;
    unsigned char* sb_hi = smem + 32768;         // [64 rows][128 B]
    unsigned char* sb_lo = smem + 32768 + 8192;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * 8; i += 128) {     // 16-byte chunks: row r, chunk c (only chunks 0..3 hold data)
        const int r = i >> 3, c = i & 7;
        uint4 vh = make_uint4(0, 0, 0, 0), vl = vh;
        if (c < K / 16) { vh = *reinterpret_cast<const uint4*>(a_hi + r * K + c * 16); vl = *reinterpret_cast<const uint4*>(a_lo + r * K + c * 16); }
        *reinterpret_cast<uint4*>(sa_hi + sw128_off(r, c)) = vh;
        *reinterpret_cast<uint4*>(sa_lo + sw128_off(r, c)) = vl;
    }
    for (int i = tid; i < N * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        uint4 vh = make_uint4(0, 0, 0, 0), vl = vh;
        if (c < K / 16) { vh = *reinterpret_cast<const uint4*>(b_hi + r * K + c * 16); vl = *reinterpret_cast<const uint4*>(b_lo + r * K + c * 16); }
        *reinterpret_cast<uint4*>(sb_hi + sw128_off(r, c)) = vh;
        *reinterpret_cast<uint4*>(sb_lo + sw128_off(r, c)) = vl;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint64_t dah = make_desc_k_sw128(smem_u32(sa_hi)), dal = make_desc_k_sw128(smem_u32(sa_lo));
        const uint64_t dbh = make_desc_k_sw128(smem_u32(sb_hi)), dbl = make_desc_k_sw128(smem_u32(sb_lo));
        for (int ks = 0; ks < K / 32; ++ks) {
            const uint64_t adv = uint64_t((ks * 32) >> 4);                 // 32 bytes per K-step inside the swizzle row
            const uint32_t acc = ks > 0;
            umma_i8(tmem + 0, dah + adv, dbh + adv, make_idesc_i8(M, N, 1, 1), acc);      // hi * hi   (s8 x s8)
            umma_i8(tmem + 64, dah + adv, dbl + adv, make_idesc_i8(M, N, 1, 0), acc);     // hi * lo   (s8 x u8)
            umma_i8(tmem + 64, dal + adv, dbh + adv, make_idesc_i8(M, N, 0, 1), 1u);      // + lo * hi (u8 x s8)
            umma_i8(tmem + 128, dal + adv, dbl + adv, make_idesc_i8(M, N, 0, 0), acc);    // lo * lo   (u8 x u8)
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {   // wait for the MMAs
        uint32_t done = 0;
        for (uint32_t spin = 0; !done; ++spin) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            if (spin > (1u << 24)) __trap();
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_addr = tmem + (uint32_t(warp * 32) << 16);
    for (int accu = 0; accu < 3; ++accu) {
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(lane_addr + uint32_t(accu * 64 + c0)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; ++j) out[(size_t(accu) * M + tid) * N + c0 + j] = int32_t(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
}

int main() {
    std::vector<uint8_t> ah(M * K), al(M * K), bh(N * K), bl(N * K);
    srand(1);
    for (auto& v : ah) v = uint8_t(rand());
    for (auto& v : al) v = uint8_t(rand());
    for (auto& v : bh) v = uint8_t(rand());
    for (auto& v : bl) v = uint8_t(rand());
    uint8_t *d_ah, *d_al, *d_bh, *d_bl; int32_t* d_out;
    cudaMalloc(&d_ah, ah.size()); cudaMalloc(&d_al, al.size()); cudaMalloc(&d_bh, bh.size()); cudaMalloc(&d_bl, bl.size());
    cudaMalloc(&d_out, 3 * M * N * 4);
    cudaMemcpy(d_ah, ah.data(), ah.size(), cudaMemcpyHostToDevice); cudaMemcpy(d_al, al.data(), al.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_bh, bh.data(), bh.size(), cudaMemcpyHostToDevice); cudaMemcpy(d_bl, bl.data(), bl.size(), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<<<1, 128, 49152>>>(d_ah, d_al, d_bh, d_bl, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<int32_t> out(3 * M * N);
    cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
    long bad[3] = { 0, 0, 0 };
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            long hh = 0, mid = 0, ll = 0;
            for (int kk = 0; kk < K; ++kk) {
                const long xh = int8_t(ah[m * K + kk]), xl = al[m * K + kk], ch = int8_t(bh[n * K + kk]), cl = bl[n * K + kk];
                hh += xh * ch; mid += xh * cl + xl * ch; ll += xl * cl;
            }
            bad[0] += out[(0 * M + m) * N + n] != hh;
            bad[1] += out[(1 * M + m) * N + n] != mid;
            bad[2] += out[(2 * M + m) * N + n] != ll;
        }
    printf("mismatches: hi*hi %ld, hi*lo+lo*hi %ld, lo*lo %ld of %d each -> %s\n", bad[0], bad[1], bad[2], M * N,
           (bad[0] | bad[1] | bad[2]) ? "FAIL" : "OK (exact)");
    printf("sample: got %d %d %d\n", out[0], out[M * N], out[2 * M * N]);
    return (bad[0] | bad[1] | bad[2]) ? 1 : 0;
}
