// Register-level FFT building blocks for the 800-sample real frame of StreamZ's front end
// (reference: streamz-rs/src/lib.rs:285-301 -- complex 800-point FFT of a real frame, |X|^2 of bins 0..400).
//
// The frame is transformed as ONE 400-point complex FFT of z[n] = x[2n] + i x[2n+1] followed by the real-input
// split, instead of the reference's 800-point complex FFT with zero imaginary parts (same bins, ~2.5x fewer flops).
// 400 = 20 x 20 Cooley-Tukey; each 20-point DFT is a 4 x 5 Good-Thomas prime-factor transform (no inner twiddles),
// fully unrolled in registers.
//
// Complex numbers are handled through the tiny `cpx` abstraction below: a (re, im) pair with per-lane add / sub / mul /
// fma, multiplication by +-i as an fma with the constant (1, -1) on the swapped pair.  The same source runs on the host
// with the same rounding points (fmaf), so tests/test_fft_math.py executes bit-for-bit the arithmetic of the kernel
// (tools/fft_math_host.cpp).
//
// SZB_PACKED_F32X2=1 maps a cpx to one 64-bit register pair and every operation to ONE packed sm_100 instruction
// (add/sub/mul/fma .f32x2 -> FADD2 / FMUL2 / FFMA2).  That halves the FP32 instruction count of the butterflies
// (20-point DFT: 112 packed instead of 224 scalar instructions, swaps folded into operand selectors), but MEASURED ON
// B200 IT IS SLOWER: the whole extraction kernel takes 5.09 ms instead of 4.00 ms per 2.2 M windows, i.e. the packed
// forms issue at well under half the scalar rate.  It is therefore off by default and kept as a documented experiment.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define SZB_HD __host__ __device__ __forceinline__
#else
#define SZB_HD inline
#endif

namespace szb {

constexpr int kWindow = 800;    // lib.rs:26 WINDOW_SIZE
constexpr int kHop = 400;       // lib.rs:288
constexpr int kBins = 401;      // lib.rs:299  WINDOW_SIZE/2 + 1
constexpr int kMels = 26;       // lib.rs:27
constexpr int kMfcc = 20;       // lib.rs:28
constexpr int kFeat = 60;       // lib.rs:30-34
constexpr int kHalf = 400;      // complex FFT length
constexpr int kR = 20;          // 400 = kR * kR

// ---- packed complex (re, im) --------------------------------------------------------------------------------------
#ifndef SZB_PACKED_F32X2
#define SZB_PACKED_F32X2 0
#endif
#if defined(__CUDA_ARCH__) && SZB_PACKED_F32X2
typedef unsigned long long cpx;
__device__ __forceinline__ cpx cpack(float re, float im) { cpx r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(re), "f"(im)); return r; }
__device__ __forceinline__ float cre(cpx a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return x; }
__device__ __forceinline__ float cim(cpx a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return y; }
__device__ __forceinline__ cpx cswap(cpx a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return cpack(y, x); }
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { cpx r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { cpx r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ cpx cmul2(cpx a, cpx b) { cpx r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }   // per lane
__device__ __forceinline__ cpx cfma2(cpx a, cpx b, cpx c) { cpx r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
#else
struct alignas(8) cpx { float x, y; };
SZB_HD cpx cpack(float re, float im) { return cpx{ re, im }; }
SZB_HD float cre(cpx a) { return a.x; }
SZB_HD float cim(cpx a) { return a.y; }
SZB_HD cpx cswap(cpx a) { return cpx{ a.y, a.x }; }
SZB_HD cpx cadd(cpx a, cpx b) { return cpx{ a.x + b.x, a.y + b.y }; }
SZB_HD cpx csub(cpx a, cpx b) { return cpx{ a.x - b.x, a.y - b.y }; }
SZB_HD cpx cmul2(cpx a, cpx b) { return cpx{ a.x * b.x, a.y * b.y }; }
SZB_HD cpx cfma2(cpx a, cpx b, cpx c) { return cpx{ fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y) }; }
#endif

// b - i d = (br + di, bi - dr),  b + i d = (br - di, bi + dr),  a + conj b,  a - conj b.
// Packed: one FFMA2 with a (+-1, -+1) constant (exact).  Scalar: plain adds (a 3-register FFMA issues at half the rate
// of an FADD on this machine: expressing these as fma(x, +-1, y) cost 9 % of the whole kernel when measured).
#if defined(__CUDA_ARCH__) && SZB_PACKED_F32X2
__device__ __forceinline__ cpx csub_i(cpx b, cpx d) { return cfma2(cswap(d), cpack(1.f, -1.f), b); }
__device__ __forceinline__ cpx cadd_i(cpx b, cpx d) { return cfma2(cswap(d), cpack(-1.f, 1.f), b); }
__device__ __forceinline__ cpx cadd_conj(cpx a, cpx b) { return cfma2(b, cpack(1.f, -1.f), a); }
__device__ __forceinline__ cpx csub_conj(cpx a, cpx b) { return cfma2(b, cpack(-1.f, 1.f), a); }
#else
SZB_HD cpx csub_i(cpx b, cpx d) { return cpx{ b.x + d.y, b.y - d.x }; }
SZB_HD cpx cadd_i(cpx b, cpx d) { return cpx{ b.x - d.y, b.y + d.x }; }
SZB_HD cpx cadd_conj(cpx a, cpx b) { return cpx{ a.x + b.x, a.y - b.y }; }
SZB_HD cpx csub_conj(cpx a, cpx b) { return cpx{ a.x - b.x, a.y + b.y }; }
#endif

// x * w, w = (wr, wi)
#if defined(__CUDA_ARCH__) && SZB_PACKED_F32X2
__device__ __forceinline__ cpx cmul_tw(cpx x, float wr, float wi) { return cfma2(cswap(x), cpack(-wi, wi), cmul2(x, cpack(wr, wr))); }
#else
SZB_HD cpx cmul_tw(cpx x, float wr, float wi) { return cpx{ fmaf(x.x, wr, -(x.y * wi)), fmaf(x.x, wi, x.y * wr) }; }
#endif

// radix-4 butterfly (forward, W4 = -i), in place
SZB_HD void bfly4(cpx& x0, cpx& x1, cpx& x2, cpx& x3) {
    const cpx a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = csub(x1, x3);
    x0 = cadd(a, c);
    x2 = csub(a, c);
    x1 = csub_i(b, d);
    x3 = cadd_i(b, d);
}

// radix-5 butterfly (forward), in place
SZB_HD void bfly5(cpx& x0, cpx& x1, cpx& x2, cpx& x3, cpx& x4) {
    constexpr float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;  // cos(2pi/5), cos(4pi/5)
    constexpr float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;   // sin(2pi/5), sin(4pi/5)
    const cpx C1 = cpack(c1, c1), C2 = cpack(c2, c2), S1 = cpack(s1, s1), S2 = cpack(s2, s2), NS1 = cpack(-s1, -s1);
    const cpx t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    const cpx a1 = cfma2(C2, t2, cfma2(C1, t1, x0)), a2 = cfma2(C1, t2, cfma2(C2, t1, x0));
    const cpx b1 = cfma2(S2, t4, cmul2(S1, t3)), b2 = cfma2(NS1, t4, cmul2(S2, t3));
    x0 = cadd(cadd(x0, t1), t2);
    x1 = csub_i(a1, b1);
    x4 = cadd_i(a1, b1);
    x2 = csub_i(a2, b2);
    x3 = cadd_i(a2, b2);
}

// 20-point forward DFT, natural order in and out: X[k] = sum_n x[n] e^{-2 pi i n k / 20}.
// Good-Thomas: n = (5a + 4b) mod 20, k = (5c + 16d) mod 20, a,c in 0..3, b,d in 0..4.
SZB_HD void dft20(cpx (&z)[20]) {
    cpx u[4][5];
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        cpx r0 = z[(4 * b) % 20], r1 = z[(5 + 4 * b) % 20], r2 = z[(10 + 4 * b) % 20], r3 = z[(15 + 4 * b) % 20];
        bfly4(r0, r1, r2, r3);
        u[0][b] = r0; u[1][b] = r1; u[2][b] = r2; u[3][b] = r3;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        bfly5(u[c][0], u[c][1], u[c][2], u[c][3], u[c][4]);
#pragma unroll
        for (int d = 0; d < 5; ++d) z[(5 * c + 16 * d) % 20] = u[c][d];
    }
}

// Row of the in-place 20x20 result that holds bin k of the 400-point FFT: stage B for column k2 writes
// Z[k2 + 20 k1] to row k2*20 + k1.
SZB_HD constexpr int row_of_bin(int k) { return (k % kR) * kR + (k / kR); }

// Real-input split for the pair (k, 400-k), 1 <= k <= 200, on UNSCALED sums (E and O carry a factor 2, so the
// returned powers are 4 |X[k]|^2; the factor is folded into the mel weights).
//   za = Z[k], zb = Z[400-k], w = (wr, wi) = e^{-2 pi i k / 800}.
//   E = za + conj zb,  D = za - conj zb,  O = -i D = (D.im, -D.re),  T = w O,
//   |X[k]|^2 ~ |E + T|^2,  |X[400-k]|^2 ~ |E - T|^2.
SZB_HD void split_pair_power(cpx za, cpx zb, float wr, float wi, float& pk, float& pmk) {
    const cpx e = cadd_conj(za, zb);
    const cpx d = csub_conj(za, zb);
#if defined(__CUDA_ARCH__) && SZB_PACKED_F32X2
    const cpx t = cfma2(d, cpack(wi, wi), cmul2(cswap(d), cpack(wr, -wr)));
    const cpx x = cadd(e, t), y = csub(e, t);
    const cpx xx = cmul2(x, x), yy = cmul2(y, y);
    pk = cre(xx) + cim(xx);
    pmk = cre(yy) + cim(yy);
#else
    const float tr = fmaf(wr, d.y, wi * d.x);       // re(w O) = wr*D.im + wi*D.re
    const float ti = fmaf(wi, d.y, -(wr * d.x));    // im(w O) = wi*D.im - wr*D.re
    const float xr = e.x + tr, xi = e.y + ti, yr = e.x - tr, yi = e.y - ti;
    pk = fmaf(xr, xr, xi * xi);
    pmk = fmaf(yr, yr, yi * yi);
#endif
}

}  // namespace szb
