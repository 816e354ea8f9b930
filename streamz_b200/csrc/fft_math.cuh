// Register-level FFT building blocks for the 800-sample real frame of StreamZ's front end
// (reference: streamz-rs/src/lib.rs:285-301 -- complex 800-point FFT of a real frame, |X|^2 of bins 0..400).
//
// The frame is transformed as ONE 400-point complex FFT of z[n] = x[2n] + i x[2n+1] followed by the real-input
// split, instead of the reference's 800-point complex FFT with zero imaginary parts (same bins, ~2.5x fewer flops).
// 400 = 20 x 20 Cooley-Tukey; each 20-point DFT is a 4 x 5 Good-Thomas prime-factor transform (no inner twiddles),
// fully unrolled in registers.  Everything here is __host__ __device__ so that tests/ can run the exact same
// arithmetic on the CPU (tests/test_fft_math.py builds tools/fft_math_host.cpp).
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

// radix-4 butterfly (forward, W4 = -i), in place on 4 complex values
SZB_HD void bfly4(float& r0, float& i0, float& r1, float& i1, float& r2, float& i2, float& r3, float& i3) {
    float ar = r0 + r2, ai = i0 + i2, br = r0 - r2, bi = i0 - i2;
    float cr = r1 + r3, ci = i1 + i3, dr = r1 - r3, di = i1 - i3;
    r0 = ar + cr; i0 = ai + ci;
    r2 = ar - cr; i2 = ai - ci;
    r1 = br + di; i1 = bi - dr;   // b - i d
    r3 = br - di; i3 = bi + dr;   // b + i d
}

// radix-5 butterfly (forward), in place on 5 complex values
SZB_HD void bfly5(float& r0, float& i0, float& r1, float& i1, float& r2, float& i2, float& r3, float& i3,
                  float& r4, float& i4) {
    constexpr float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;  // cos(2pi/5), cos(4pi/5)
    constexpr float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;   // sin(2pi/5), sin(4pi/5)
    float t1r = r1 + r4, t1i = i1 + i4, t2r = r2 + r3, t2i = i2 + i3;
    float t3r = r1 - r4, t3i = i1 - i4, t4r = r2 - r3, t4i = i2 - i3;
    float a1r = fmaf(c2, t2r, fmaf(c1, t1r, r0)), a1i = fmaf(c2, t2i, fmaf(c1, t1i, i0));
    float a2r = fmaf(c1, t2r, fmaf(c2, t1r, r0)), a2i = fmaf(c1, t2i, fmaf(c2, t1i, i0));
    float b1r = fmaf(s2, t4r, s1 * t3r), b1i = fmaf(s2, t4i, s1 * t3i);
    float b2r = fmaf(-s1, t4r, s2 * t3r), b2i = fmaf(-s1, t4i, s2 * t3i);
    r0 = r0 + t1r + t2r; i0 = i0 + t1i + t2i;
    r1 = a1r + b1i; i1 = a1i - b1r;   // a1 - i b1
    r4 = a1r - b1i; i4 = a1i + b1r;   // a1 + i b1
    r2 = a2r + b2i; i2 = a2i - b2r;
    r3 = a2r - b2i; i3 = a2i + b2r;
}

// 20-point forward DFT, natural order in and out: X[k] = sum_n x[n] e^{-2 pi i n k / 20}.
// Good-Thomas: n = (5a + 4b) mod 20, k = (5c + 16d) mod 20, a,c in 0..3, b,d in 0..4.
SZB_HD void dft20(float (&re)[20], float (&im)[20]) {
    float ur[4][5], ui[4][5];
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const int n0 = (4 * b) % 20, n1 = (5 + 4 * b) % 20, n2 = (10 + 4 * b) % 20, n3 = (15 + 4 * b) % 20;
        float r0 = re[n0], i0 = im[n0], r1 = re[n1], i1 = im[n1], r2 = re[n2], i2 = im[n2], r3 = re[n3], i3 = im[n3];
        bfly4(r0, i0, r1, i1, r2, i2, r3, i3);
        ur[0][b] = r0; ui[0][b] = i0; ur[1][b] = r1; ui[1][b] = i1;
        ur[2][b] = r2; ui[2][b] = i2; ur[3][b] = r3; ui[3][b] = i3;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        bfly5(ur[c][0], ui[c][0], ur[c][1], ui[c][1], ur[c][2], ui[c][2], ur[c][3], ui[c][3], ur[c][4], ui[c][4]);
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            const int k = (5 * c + 16 * d) % 20;
            re[k] = ur[c][d];
            im[k] = ui[c][d];
        }
    }
}

// Row of the in-place 20x20 result that holds bin k of the 400-point FFT: stage B for column k2 writes
// Z[k2 + 20 k1] to row k2*20 + k1.
SZB_HD constexpr int row_of_bin(int k) { return (k % kR) * kR + (k / kR); }

// Real-input split for the pair (k, 400-k), 1 <= k <= 200, on UNSCALED sums (E and O carry a factor 2, so the
// returned powers are 4 |X[k]|^2; the factor is folded into the mel weights).
//   za = Z[k], zb = Z[400-k], w = e^{-2 pi i k / 800}.
SZB_HD void split_pair_power(float zar, float zai, float zbr, float zbi, float wr, float wi, float& pk, float& pmk) {
    float er = zar + zbr, ei = zai - zbi;       // E = Z[k] + conj Z[400-k]
    float dr = zar - zbr, di = zai + zbi;       // D = Z[k] - conj Z[400-k];  O = -i D = (di, -dr)
    float tr = fmaf(wr, di, wi * dr);           // T = w * O : re = wr*di - wi*(-dr)
    float ti = fmaf(wi, di, -(wr * dr));        //             im = wr*(-dr) + wi*di
    float xr = er + tr, xi = ei + ti, yr = er - tr, yi = ei - ti;
    pk = fmaf(xr, xr, xi * xi);
    pmk = fmaf(yr, yr, yi * yi);
}

}  // namespace szb
