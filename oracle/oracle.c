/*
 * CPU oracle / CPU baseline for the StreamZ hot path: a scalar float32 C restatement of the reference's algorithm.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may load this library.  The product (streamz_b200) never links or calls it.
 *
 * PARITY STATUS: parity unpinned at the third-party-crate boundaries (rustfft / rustdct / mel_filter / ndarray):
 * the Rust reference cannot be built here and its tests pin no numbers (lib.rs:1827-1865).  Tables (mel bank, DCT,
 * resampler taps) are passed in by the caller, who builds them with oracle/streamz_oracle.py; this file restates
 * the arithmetic that is in the reference's own source, keeping its algorithmic choices:
 *   - complex 800-point FFT of the real frame, |X|^2 on bins 0..400, DENSE 26x401 mel loop   (lib.rs:285-310)
 *   - ln(max(sum, 1e-12)), 26-point DCT-II truncated to 20                                    (lib.rs:309-315)
 *   - delta / delta-delta with edge replication, per-window z-score, std >= 1e-6              (lib.rs:212-228, 321-342)
 *   - per-window heap vectors                                                                  (lib.rs:297, 303, 325)
 *   - MLP forward per window, per-sample outer-product gradient accumulation, forward computed twice per window in
 *     the epoch loop (once for the loss, once inside train_batch)                             (lib.rs:604-620, 1013-1045)
 * It is "the reference's CPU path, restated in C" (cpu_baseline.kind = "port"), not the Rust binary.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WIN 800
#define HOP 400
#define NBINS 401
#define NMEL 26
#define NMFCC 20
#define NFEAT 60

/* ---------------------------------------------------------------------------------------------------------------- */
/* 800-point complex FFT, float32, Stockham autosort, radices 5,5,4,4,2 (stands in for rustfft plan_fft_forward(800)) */
/* ---------------------------------------------------------------------------------------------------------------- */

typedef struct { float re, im; } cpx;

static cpx g_tw[WIN];
static pthread_once_t g_tw_once = PTHREAD_ONCE_INIT;
static void init_tw(void) {
    for (int i = 0; i < WIN; ++i) {
        double a = -2.0 * M_PI * (double)i / (double)WIN;
        g_tw[i].re = (float)cos(a);
        g_tw[i].im = (float)sin(a);
    }
}

static inline cpx cmul(cpx a, cpx b) { cpx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static inline cpx cadd(cpx a, cpx b) { cpx r = { a.re + b.re, a.im + b.im }; return r; }
static inline cpx csub(cpx a, cpx b) { cpx r = { a.re - b.re, a.im - b.im }; return r; }
static inline cpx cmulnegi(cpx a) { cpx r = { a.im, -a.re }; return r; } /* a * (-i) */

/* One decimation-in-frequency Stockham pass of radix 5: n = 800, s = product of the radices already done,
 * m = n / (5 s); the twiddle of output k is W_{5m}^{p k}. */
static void pass_r5(const cpx* x, cpx* y, int n, int s) {
    int m = n / (s * 5);
    const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f; /* cos(2pi/5), cos(4pi/5) */
    const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;  /* sin(2pi/5), sin(4pi/5) */
    for (int p = 0; p < m; ++p) {
        int twstep = WIN / (m * 5);
        cpx w1 = g_tw[(p * twstep) % WIN], w2 = g_tw[(2 * p * twstep) % WIN];
        cpx w3 = g_tw[(3 * p * twstep) % WIN], w4 = g_tw[(4 * p * twstep) % WIN];
        for (int q = 0; q < s; ++q) {
            cpx x0 = x[q + s * p], x1 = x[q + s * (p + m)], x2 = x[q + s * (p + 2 * m)];
            cpx x3 = x[q + s * (p + 3 * m)], x4 = x[q + s * (p + 4 * m)];
            cpx t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
            cpx a1 = { x0.re + c1 * t1.re + c2 * t2.re, x0.im + c1 * t1.im + c2 * t2.im };
            cpx a2 = { x0.re + c2 * t1.re + c1 * t2.re, x0.im + c2 * t1.im + c1 * t2.im };
            cpx b1 = { s1 * t3.re + s2 * t4.re, s1 * t3.im + s2 * t4.im };
            cpx b2 = { s2 * t3.re - s1 * t4.re, s2 * t3.im - s1 * t4.im };
            cpx y0 = { x0.re + t1.re + t2.re, x0.im + t1.im + t2.im };
            cpx y1 = { a1.re + b1.im, a1.im - b1.re }; /* a1 - i b1 */
            cpx y4 = { a1.re - b1.im, a1.im + b1.re };
            cpx y2 = { a2.re + b2.im, a2.im - b2.re };
            cpx y3 = { a2.re - b2.im, a2.im + b2.re };
            y[q + s * (5 * p + 0)] = y0;
            y[q + s * (5 * p + 1)] = cmul(y1, w1);
            y[q + s * (5 * p + 2)] = cmul(y2, w2);
            y[q + s * (5 * p + 3)] = cmul(y3, w3);
            y[q + s * (5 * p + 4)] = cmul(y4, w4);
        }
    }
}

static void pass_r4(const cpx* x, cpx* y, int n, int s) {
    int m = n / (s * 4);
    for (int p = 0; p < m; ++p) {
        int twstep = WIN / (m * 4);
        cpx w1 = g_tw[(p * twstep) % WIN], w2 = g_tw[(2 * p * twstep) % WIN], w3 = g_tw[(3 * p * twstep) % WIN];
        for (int q = 0; q < s; ++q) {
            cpx a = x[q + s * (p + m * 0)], b = x[q + s * (p + m * 1)];
            cpx c = x[q + s * (p + m * 2)], d = x[q + s * (p + m * 3)];
            cpx apc = cadd(a, c), amc = csub(a, c), bpd = cadd(b, d), bmd = cmulnegi(csub(b, d));
            y[q + s * (4 * p + 0)] = cadd(apc, bpd);
            y[q + s * (4 * p + 1)] = cmul(cadd(amc, bmd), w1);
            y[q + s * (4 * p + 2)] = cmul(csub(apc, bpd), w2);
            y[q + s * (4 * p + 3)] = cmul(csub(amc, bmd), w3);
        }
    }
}

static void pass_r2(const cpx* x, cpx* y, int n, int s) {
    int m = n / (s * 2);
    for (int p = 0; p < m; ++p) {
        cpx w1 = g_tw[(p * (WIN / (m * 2))) % WIN];
        for (int q = 0; q < s; ++q) {
            cpx a = x[q + s * p], b = x[q + s * (p + m)];
            y[q + s * (2 * p + 0)] = cadd(a, b);
            y[q + s * (2 * p + 1)] = cmul(csub(a, b), w1);
        }
    }
}

/* in-place from the caller's point of view: result left in buf (scratch is a second 800-element buffer) */
static void fft800(cpx* buf, cpx* scratch) {
    pthread_once(&g_tw_once, init_tw);
    cpx *x = buf, *y = scratch;
    int s = 1;
    static const int radices[5] = { 5, 5, 4, 4, 2 };
    for (int i = 0; i < 5; ++i) {
        int R = radices[i];
        if (R == 4) pass_r4(x, y, WIN, s);
        else if (R == 2) pass_r2(x, y, WIN, s);
        else pass_r5(x, y, WIN, s);
        s *= R;
        cpx* t = x; x = y; y = t;
    }
    if (x != buf) memcpy(buf, x, sizeof(cpx) * WIN);
}

void so_fft800(const float* in_re, float* out_re, float* out_im) { /* exported for the oracle self-test */
    cpx a[WIN], b[WIN];
    for (int i = 0; i < WIN; ++i) { a[i].re = in_re[i]; a[i].im = 0.f; }
    fft800(a, b);
    for (int i = 0; i < WIN; ++i) { out_re[i] = a[i].re; out_im[i] = a[i].im; }
}


/* ---------------------------------------------------------------------------------------------------------------- */
/* Hardened CPU arm: the same Stockham passes with LANE = FRAME.  The scalar fft800 above runs at ~9 us per frame,     */
/* several times slower than rustfft's AVX butterflies would; here VL frames travel through every butterfly together   */
/* (structure of arrays: element e of lane l at [e * VL + l]), so the inner loops are unit-stride and gcc turns them    */
/* into AVX2 / AVX-512 arithmetic (check: make -C oracle vecreport).  Per frame the operations and their order are      */
/* exactly those of the scalar passes.  so_set_mode: 0 = scalar (as written), 1 = SIMD FFT, mel / DCT per frame in the  */
/* reference's sequential summation order (the default of the timed baseline: closest to what the Rust binary does:     */
/* rustfft is SIMD, `iter().zip().map().sum()` is not), 2 = mel and DCT batched over the lanes too (a stronger baseline */
/* than the reference itself).                                                                                          */
/* ---------------------------------------------------------------------------------------------------------------- */
#define VL 8
static int g_mode = 1;
void so_set_mode(int mode) { g_mode = mode; }
int so_get_mode(void) { return g_mode; }

typedef struct { float* re; float* im; } vbuf;   /* [WIN * VL] each */

#define VLOOP for (int l = 0; l < VL; ++l)
typedef float vf __attribute__((vector_size(VL * 4), aligned(4)));   /* GCC vector extension: one value per frame */
#define VLD(ptr, e) (*(const vf*)((ptr) + (size_t)(e) * VL))
#define VST(ptr, e) (*(vf*)((ptr) + (size_t)(e) * VL))

static void vpass_r5(vbuf x, vbuf y, int n, int s) {
    const int m = n / (s * 5);
    const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f, s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
    for (int p = 0; p < m; ++p) {
        const int twstep = WIN / (m * 5);
        const cpx w1 = g_tw[(p * twstep) % WIN], w2 = g_tw[(2 * p * twstep) % WIN], w3 = g_tw[(3 * p * twstep) % WIN],
                  w4 = g_tw[(4 * p * twstep) % WIN];
        for (int q = 0; q < s; ++q) {
            const int i0 = q + s * p, i1 = q + s * (p + m), i2 = q + s * (p + 2 * m), i3 = q + s * (p + 3 * m), i4 = q + s * (p + 4 * m);
            const vf x0r = VLD(x.re, i0), x0i = VLD(x.im, i0), x1r = VLD(x.re, i1), x1i = VLD(x.im, i1), x2r = VLD(x.re, i2),
                     x2i = VLD(x.im, i2), x3r = VLD(x.re, i3), x3i = VLD(x.im, i3), x4r = VLD(x.re, i4), x4i = VLD(x.im, i4);
            const vf t1r = x1r + x4r, t1i = x1i + x4i, t2r = x2r + x3r, t2i = x2i + x3i;
            const vf t3r = x1r - x4r, t3i = x1i - x4i, t4r = x2r - x3r, t4i = x2i - x3i;
            const vf a1r = x0r + c1 * t1r + c2 * t2r, a1i = x0i + c1 * t1i + c2 * t2i;
            const vf a2r = x0r + c2 * t1r + c1 * t2r, a2i = x0i + c2 * t1i + c1 * t2i;
            const vf b1r = s1 * t3r + s2 * t4r, b1i = s1 * t3i + s2 * t4i;
            const vf b2r = s2 * t3r - s1 * t4r, b2i = s2 * t3i - s1 * t4i;
            const vf u1r = a1r + b1i, u1i = a1i - b1r, u4r = a1r - b1i, u4i = a1i + b1r;
            const vf u2r = a2r + b2i, u2i = a2i - b2r, u3r = a2r - b2i, u3i = a2i + b2r;
            const int o = q + s * 5 * p;
            VST(y.re, o) = x0r + t1r + t2r;                      VST(y.im, o) = x0i + t1i + t2i;
            VST(y.re, o + s) = u1r * w1.re - u1i * w1.im;        VST(y.im, o + s) = u1r * w1.im + u1i * w1.re;
            VST(y.re, o + 2 * s) = u2r * w2.re - u2i * w2.im;    VST(y.im, o + 2 * s) = u2r * w2.im + u2i * w2.re;
            VST(y.re, o + 3 * s) = u3r * w3.re - u3i * w3.im;    VST(y.im, o + 3 * s) = u3r * w3.im + u3i * w3.re;
            VST(y.re, o + 4 * s) = u4r * w4.re - u4i * w4.im;    VST(y.im, o + 4 * s) = u4r * w4.im + u4i * w4.re;
        }
    }
}
static void vpass_r4(vbuf x, vbuf y, int n, int s) {
    const int m = n / (s * 4);
    for (int p = 0; p < m; ++p) {
        const int twstep = WIN / (m * 4);
        const cpx w1 = g_tw[(p * twstep) % WIN], w2 = g_tw[(2 * p * twstep) % WIN], w3 = g_tw[(3 * p * twstep) % WIN];
        for (int q = 0; q < s; ++q) {
            const int i0 = q + s * p, i1 = q + s * (p + m), i2 = q + s * (p + 2 * m), i3 = q + s * (p + 3 * m);
            const vf ar = VLD(x.re, i0), ai = VLD(x.im, i0), br = VLD(x.re, i1), bi = VLD(x.im, i1), cr = VLD(x.re, i2), ci = VLD(x.im, i2),
                     dr = VLD(x.re, i3), di = VLD(x.im, i3);
            const vf apcr = ar + cr, apci = ai + ci, amcr = ar - cr, amci = ai - ci, bpdr = br + dr, bpdi = bi + di;
            const vf bmdr = bi - di, bmdi = dr - br;                                /* (b - d) * (-i) */
            const vf u1r = amcr + bmdr, u1i = amci + bmdi, u2r = apcr - bpdr, u2i = apci - bpdi, u3r = amcr - bmdr, u3i = amci - bmdi;
            const int o = q + s * 4 * p;
            VST(y.re, o) = apcr + bpdr;                          VST(y.im, o) = apci + bpdi;
            VST(y.re, o + s) = u1r * w1.re - u1i * w1.im;        VST(y.im, o + s) = u1r * w1.im + u1i * w1.re;
            VST(y.re, o + 2 * s) = u2r * w2.re - u2i * w2.im;    VST(y.im, o + 2 * s) = u2r * w2.im + u2i * w2.re;
            VST(y.re, o + 3 * s) = u3r * w3.re - u3i * w3.im;    VST(y.im, o + 3 * s) = u3r * w3.im + u3i * w3.re;
        }
    }
}
static void vpass_r2(vbuf x, vbuf y, int n, int s) {
    const int m = n / (s * 2);
    for (int p = 0; p < m; ++p) {
        const cpx w1 = g_tw[(p * (WIN / (m * 2))) % WIN];
        for (int q = 0; q < s; ++q) {
            const int i0 = q + s * p, i1 = q + s * (p + m), o = q + s * 2 * p;
            const vf ar = VLD(x.re, i0), ai = VLD(x.im, i0), br = VLD(x.re, i1), bi = VLD(x.im, i1);
            const vf ur = ar - br, ui = ai - bi;
            VST(y.re, o) = ar + br;                              VST(y.im, o) = ai + bi;
            VST(y.re, o + s) = ur * w1.re - ui * w1.im;          VST(y.im, o + s) = ur * w1.im + ui * w1.re;
        }
    }
}
/* VL frames at once; returns the buffer that holds the result (a or b) */
static vbuf vfft800(vbuf a, vbuf b) {
    pthread_once(&g_tw_once, init_tw);
    vbuf x = a, y = b;
    int s = 1;
    static const int radices[5] = { 5, 5, 4, 4, 2 };
    for (int i = 0; i < 5; ++i) {
        const int R = radices[i];
        if (R == 4) vpass_r4(x, y, WIN, s);
        else if (R == 2) vpass_r2(x, y, WIN, s);
        else vpass_r5(x, y, WIN, s);
        s *= R;
        vbuf t = x; x = y; y = t;
    }
    return x;
}

/* MFCCs of windows [w0, w0 + VL) (lanes past n repeat the last window; their results are dropped by the caller) */
static void mfcc_group(const int16_t* pcm, size_t n, size_t w0, const float* mel, const float* dct, vbuf a, vbuf b, float* mags,
                       float* base) {
    for (int l = 0; l < VL; ++l) {
        const size_t w = w0 + (size_t)l < n ? w0 + (size_t)l : n - 1;
        const int16_t* chunk = pcm + w * HOP;
        for (int i = 0; i < WIN; ++i) { a.re[(size_t)i * VL + l] = (float)chunk[i] / 32767.0f; a.im[(size_t)i * VL + l] = 0.f; }  /* lib.rs:293-295 */
    }
    const vbuf r = vfft800(a, b);                                                                  /* lib.rs:296 */
    for (int k = 0; k < NBINS; ++k) VST(mags, k) = VLD(r.re, k) * VLD(r.re, k) + VLD(r.im, k) * VLD(r.im, k);
    float energies[NMEL][VL];
    if (g_mode >= 2) {                        /* all lanes walk the 401 bins together: same order per frame, SIMD across frames */
        for (int m = 0; m < NMEL; ++m) {
            vf sum = { 0.f };
            const float* filt = mel + (size_t)m * NBINS;
            for (int k = 0; k < NBINS; ++k) sum += filt[k] * VLD(mags, k);
            VLOOP energies[m][l] = logf(fmaxf(sum[l], 1e-12f));
        }
    } else {                                  /* one frame at a time, sequential sum: what `.zip().map().sum()` compiles to */
        float one[NBINS];                     /* this frame's mags, contiguous like the reference's per-window Vec (lib.rs:297) */
        for (int l = 0; l < VL; ++l) {
            for (int k = 0; k < NBINS; ++k) one[k] = mags[(size_t)k * VL + l];
            for (int m = 0; m < NMEL; ++m) {
                float sum = 0.f;
                const float* filt = mel + (size_t)m * NBINS;
                for (int k = 0; k < NBINS; ++k) sum += filt[k] * one[k];                          /* lib.rs:304-308 */
                energies[m][l] = logf(fmaxf(sum, 1e-12f));                                       /* lib.rs:309 */
            }
        }
    }
    for (int l = 0; l < VL && w0 + (size_t)l < n; ++l)
        for (int j = 0; j < NMFCC; ++j) {                                                         /* lib.rs:312-314 */
            float acc = 0.f;
            for (int m = 0; m < NMEL; ++m) acc += dct[j * NMEL + m] * energies[m][l];
            base[(w0 + (size_t)l) * NMFCC + j] = acc;
        }
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* Front end (lib.rs:279-345)                                                                                        */
/* ---------------------------------------------------------------------------------------------------------------- */

size_t so_n_windows(size_t n) { return n < WIN ? 0 : (n - WIN) / HOP + 1; }

/* mel: [26][401] row-major f32; dct: [20][26] row-major f32 (unscaled DCT-II rows).  out: [n][60]. Returns n. */
size_t so_extract(const int16_t* pcm, size_t len, const float* mel, const float* dct, float* out) {
    size_t n = so_n_windows(len);
    if (n == 0) return 0;
    cpx* buffer = (cpx*)malloc(sizeof(cpx) * WIN);           /* lib.rs:285 */
    cpx* scratch = (cpx*)malloc(sizeof(cpx) * WIN);
    float* base = (float*)malloc(sizeof(float) * n * NMFCC);
    if (g_mode >= 1) {
        vbuf a = { (float*)aligned_alloc(64, sizeof(float) * WIN * VL), (float*)aligned_alloc(64, sizeof(float) * WIN * VL) };
        vbuf b = { (float*)aligned_alloc(64, sizeof(float) * WIN * VL), (float*)aligned_alloc(64, sizeof(float) * WIN * VL) };
        float* vmags = (float*)aligned_alloc(64, sizeof(float) * ((NBINS * VL + 15) / 16 * 16));
        for (size_t w0 = 0; w0 < n; w0 += VL) mfcc_group(pcm, n, w0, mel, dct, a, b, vmags, base);
        free(a.re); free(a.im); free(b.re); free(b.im); free(vmags);
    } else
    for (size_t w = 0; w < n; ++w) {                          /* lib.rs:291 */
        const int16_t* chunk = pcm + w * HOP;
        for (int i = 0; i < WIN; ++i) { buffer[i].re = (float)chunk[i] / 32767.0f; buffer[i].im = 0.f; } /* lib.rs:293-295 */
        fft800(buffer, scratch);                              /* lib.rs:296 */
        float* mags = (float*)malloc(sizeof(float) * NBINS);  /* lib.rs:297: a fresh Vec per window */
        for (int k = 0; k < NBINS; ++k) mags[k] = buffer[k].re * buffer[k].re + buffer[k].im * buffer[k].im;
        float* energies = (float*)malloc(sizeof(float) * NMEL); /* lib.rs:303 */
        for (int m = 0; m < NMEL; ++m) {                      /* dense dot, lib.rs:304-308 */
            float sum = 0.f;
            const float* filt = mel + (size_t)m * NBINS;
            for (int k = 0; k < NBINS; ++k) sum += filt[k] * mags[k];
            energies[m] = logf(fmaxf(sum, 1e-12f));           /* lib.rs:309 */
        }
        for (int j = 0; j < NMFCC; ++j) {                     /* lib.rs:312-314 */
            float acc = 0.f;
            for (int m = 0; m < NMEL; ++m) acc += dct[j * NMEL + m] * energies[m];
            base[w * NMFCC + j] = acc;
        }
        free(mags);
        free(energies);
    }
    float* d1 = (float*)malloc(sizeof(float) * n * NMFCC);
    float* d2 = (float*)malloc(sizeof(float) * n * NMFCC);
    for (int pass = 0; pass < 2; ++pass) {                    /* lib.rs:212-228, 321-322 */
        const float* src = pass == 0 ? base : d1;
        float* dst = pass == 0 ? d1 : d2;
        for (size_t i = 0; i < n; ++i) {
            const float* prev = src + (i > 0 ? i - 1 : i) * NMFCC;
            const float* next = src + (i + 1 < n ? i + 1 : i) * NMFCC;
            for (int j = 0; j < NMFCC; ++j) dst[i * NMFCC + j] = (next[j] - prev[j]) / 2.0f;
        }
    }
    for (size_t i = 0; i < n; ++i) {                          /* lib.rs:324-342 */
        float* frame = out + i * NFEAT;
        memcpy(frame, base + i * NMFCC, sizeof(float) * NMFCC);
        memcpy(frame + NMFCC, d1 + i * NMFCC, sizeof(float) * NMFCC);
        memcpy(frame + 2 * NMFCC, d2 + i * NMFCC, sizeof(float) * NMFCC);
        float sum = 0.f;
        for (int j = 0; j < NFEAT; ++j) sum += frame[j];
        float mean = sum / (float)NFEAT;
        float var = 0.f;
        for (int j = 0; j < NFEAT; ++j) { float d = frame[j] - mean; var += d * d; }
        var /= (float)NFEAT;
        float sd = fmaxf(sqrtf(var), 1e-6f);
        for (int j = 0; j < NFEAT; ++j) frame[j] = (frame[j] - mean) / sd;
    }
    free(buffer); free(scratch); free(base); free(d1); free(d2);
    return n;
}

/* ---- resampler (this repo's polyphase spec; see streamz_oracle.py) ---- */
size_t so_resample(const int16_t* in, size_t n_in, uint32_t rate, const float* taps, uint32_t L, uint32_t M, uint32_t T,
                   int16_t* out) {
    size_t n_out = (size_t)(((unsigned long long)n_in * 44100ull) / rate);
    long long i0 = 0;
    uint32_t p = 0;                              /* j M = i0 L + p, advanced incrementally (M may exceed L) */
    for (size_t j = 0; j < n_out; ++j, p += M) {
        while (p >= L) { p -= L; ++i0; }
        const float* c = taps + (size_t)p * T;
        float acc = 0.f;
        for (uint32_t t = 0; t < T; ++t) {
            long long i = i0 - (long long)(T / 2 - 1) + (long long)t;
            float x = (i >= 0 && (size_t)i < n_in) ? (float)in[i] : 0.f;
            acc = fmaf(c[t], x, acc);
        }
        acc = fminf(fmaxf(acc, -32768.f), 32767.f);
        out[j] = (int16_t)acc; /* C cast truncates toward zero, like Rust `as i16` on an in-range value */
    }
    return n_out;
}

/* ---- multi-threaded batch: one clip per thread at a time, like rayon's pool over clips (main.rs:500-508) ---- */
typedef struct {
    const int16_t* pcm; const uint64_t* clip_off; const uint64_t* win_off; uint32_t n_clips;
    const float* mel; const float* dct; float* out;
    /* optional resample stage */
    uint32_t rate; const float* taps; uint32_t L, M, T;
    volatile uint32_t* next;
} batch_job;

static void* batch_worker(void* arg) {
    batch_job* jb = (batch_job*)arg;
    for (;;) {
        uint32_t c = __sync_fetch_and_add(jb->next, 1);
        if (c >= jb->n_clips) break;
        const int16_t* src = jb->pcm + jb->clip_off[c];
        size_t len = (size_t)(jb->clip_off[c + 1] - jb->clip_off[c]);
        if (jb->rate != 44100 && jb->rate != 0) {
            size_t n_out = (size_t)(((unsigned long long)len * 44100ull) / jb->rate);
            int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * (n_out ? n_out : 1));
            so_resample(src, len, jb->rate, jb->taps, jb->L, jb->M, jb->T, tmp);
            so_extract(tmp, n_out, jb->mel, jb->dct, jb->out + jb->win_off[c] * NFEAT);
            free(tmp);
        } else {
            so_extract(src, len, jb->mel, jb->dct, jb->out + jb->win_off[c] * NFEAT);
        }
    }
    return NULL;
}

/* clip_off: [n_clips+1] sample offsets into pcm; win_off: [n_clips+1] window offsets into out (caller computes).
 * rate == 44100 (or 0): no resample.  threads <= 0: 1. */
void so_extract_batch(const int16_t* pcm, const uint64_t* clip_off, const uint64_t* win_off, uint32_t n_clips,
                      const float* mel, const float* dct, uint32_t rate, const float* taps, uint32_t L, uint32_t M,
                      uint32_t T, float* out, int threads) {
    volatile uint32_t next = 0;
    batch_job jb = { pcm, clip_off, win_off, n_clips, mel, dct, out, rate, taps, L, M, T, &next };
    if (threads <= 1) { batch_worker(&jb); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, batch_worker, &jb);
    for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
    free(th);
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* SimpleNeuralNet (lib.rs:880-891, 1002-1060); sizes are generic like the reference's constructor (lib.rs:767)       */
/* ---------------------------------------------------------------------------------------------------------------- */

typedef struct {
    int n_in, h1, h2, n_out;
    float *w1, *b1, *w2, *b2, *w3, *b3; /* row-major w1[n_in][h1] w2[h1][h2] w3[h2][n_out] */
} so_net;

static void vec_mat(const float* x, const float* w, const float* b, int rows, int cols, float* y) {
    for (int j = 0; j < cols; ++j) y[j] = 0.f;
    for (int i = 0; i < rows; ++i) {
        float xi = x[i];
        const float* wr = w + (size_t)i * cols;
        for (int j = 0; j < cols; ++j) y[j] += xi * wr[j];
    }
    for (int j = 0; j < cols; ++j) y[j] += b[j];
}

static void net_forward_full(const so_net* n, const float* x, float* a1, float* h1, float* h2, float* p) {
    vec_mat(x, n->w1, n->b1, n->n_in, n->h1, a1);
    for (int j = 0; j < n->h1; ++j) h1[j] = a1[j] > 0.f ? a1[j] : 0.f;              /* lib.rs:882 */
    vec_mat(h1, n->w2, n->b2, n->h1, n->h2, h2);
    for (int j = 0; j < n->h2; ++j) h2[j] = tanhf(h2[j]);                            /* lib.rs:883 */
    vec_mat(h2, n->w3, n->b3, n->h2, n->n_out, p);
    float mx = -INFINITY;
    for (int j = 0; j < n->n_out; ++j) mx = fmaxf(mx, p[j]);                        /* lib.rs:887 */
    float sum = 0.f;
    for (int j = 0; j < n->n_out; ++j) { p[j] = expf(p[j] - mx); sum += p[j]; }
    for (int j = 0; j < n->n_out; ++j) p[j] /= sum;
}

/* probs: [B][n_out] */
void so_forward(const so_net* n, const float* x, size_t B, float* probs) {
    float* a1 = (float*)malloc(sizeof(float) * (size_t)(2 * n->h1 + n->h2));
    float *h1 = a1 + n->h1, *h2 = h1 + n->h1;
    for (size_t r = 0; r < B; ++r) net_forward_full(n, x + r * (size_t)n->n_in, a1, h1, h2, probs + r * (size_t)n->n_out);
    free(a1);
}

/* targets: [B][n_out] when per_row != 0, else one shared [n_out] vector (the reference's signature, lib.rs:1002). */
void so_train_batch(so_net* n, const float* x, size_t B, const float* targets, int per_row, float lr) {
    if (B == 0) return;                                                              /* lib.rs:1003-1005 */
    size_t s1 = (size_t)n->n_in * n->h1, s2 = (size_t)n->h1 * n->h2, s3 = (size_t)n->h2 * n->n_out;
    float* g = (float*)calloc(s1 + s2 + s3 + (size_t)(n->h1 + n->h2 + n->n_out), sizeof(float));
    float *gw1 = g, *gw2 = gw1 + s1, *gw3 = gw2 + s2, *gb1 = gw3 + s3, *gb2 = gb1 + n->h1, *gb3 = gb2 + n->h2;
    float* tmp = (float*)malloc(sizeof(float) * (size_t)(3 * n->h1 + 2 * n->h2 + 2 * n->n_out));
    float *a1 = tmp, *h1 = a1 + n->h1, *d1 = h1 + n->h1, *h2 = d1 + n->h1, *d2 = h2 + n->h2, *p = d2 + n->h2, *d3 = p + n->n_out;
    for (size_t r = 0; r < B; ++r) {
        const float* xr = x + r * (size_t)n->n_in;
        const float* t = per_row ? targets + r * (size_t)n->n_out : targets;
        net_forward_full(n, xr, a1, h1, h2, p);                                      /* lib.rs:1016-1026 */
        for (int j = 0; j < n->n_out; ++j) d3[j] = p[j] - t[j];                      /* lib.rs:1028 */
        for (int i = 0; i < n->h2; ++i) {                                            /* outer product, lib.rs:1029-1033 */
            float hi = h2[i]; float* gr = gw3 + (size_t)i * n->n_out;
            for (int j = 0; j < n->n_out; ++j) gr[j] += hi * d3[j];
        }
        for (int j = 0; j < n->n_out; ++j) gb3[j] += d3[j];
        for (int i = 0; i < n->h2; ++i) {                                            /* lib.rs:1034 */
            float acc = 0.f; const float* wr = n->w3 + (size_t)i * n->n_out;
            for (int j = 0; j < n->n_out; ++j) acc += d3[j] * wr[j];
            d2[i] = acc * (1.f - h2[i] * h2[i]);
        }
        for (int i = 0; i < n->h1; ++i) {                                            /* lib.rs:1035-1038 */
            float hi = h1[i]; float* gr = gw2 + (size_t)i * n->h2;
            for (int j = 0; j < n->h2; ++j) gr[j] += hi * d2[j];
        }
        for (int j = 0; j < n->h2; ++j) gb2[j] += d2[j];
        for (int i = 0; i < n->h1; ++i) {                                            /* lib.rs:1039-1040 */
            float acc = 0.f; const float* wr = n->w2 + (size_t)i * n->h2;
            for (int j = 0; j < n->h2; ++j) acc += d2[j] * wr[j];
            d1[i] = a1[i] > 0.f ? acc : 0.f;
        }
        for (int i = 0; i < n->n_in; ++i) {                                          /* lib.rs:1041-1044 */
            float xi = xr[i]; float* gr = gw1 + (size_t)i * n->h1;
            for (int j = 0; j < n->h1; ++j) gr[j] += xi * d1[j];
        }
        for (int j = 0; j < n->h1; ++j) gb1[j] += d1[j];
    }
    float scale = lr / (float)B;                                                     /* lib.rs:1047 */
    for (size_t i = 0; i < s1; ++i) n->w1[i] -= gw1[i] * scale;
    for (size_t i = 0; i < s2; ++i) n->w2[i] -= gw2[i] * scale;
    for (size_t i = 0; i < s3; ++i) n->w3[i] -= gw3[i] * scale;
    for (int j = 0; j < n->h1; ++j) n->b1[j] -= gb1[j] * scale;
    for (int j = 0; j < n->h2; ++j) n->b2[j] -= gb2[j] * scale;
    for (int j = 0; j < n->n_out; ++j) n->b3[j] -= gb3[j] * scale;
    free(g); free(tmp);
}

/* One epoch of lib.rs:599-622 with injected randomness.  feats [n][n_in]; labels [n]; perm [n_perm]; keep [n][n_in]
 * (u8, may be NULL).  Returns the number of surviving windows; *loss_sum gets the summed cross-entropy. */
size_t so_train_epoch(so_net* n, const float* feats, const uint32_t* labels, const uint32_t* perm, size_t n_perm,
                      size_t batch, float lr, const uint8_t* keep, double* loss_sum) {
    if (batch == 0) batch = 1;
    size_t count = 0; double loss = 0.0;
    float* xb = (float*)malloc(sizeof(float) * batch * (size_t)n->n_in);
    float* tb = (float*)malloc(sizeof(float) * batch * (size_t)n->n_out);
    float* p = (float*)malloc(sizeof(float) * (size_t)n->n_out);
    for (size_t s = 0; s < n_perm; s += batch) {
        size_t e = s + batch < n_perm ? s + batch : n_perm, nb = 0;
        for (size_t r = s; r < e; ++r) {
            size_t w = perm[r];
            float* xr = xb + nb * (size_t)n->n_in;
            int all_zero = 1;
            for (int i = 0; i < n->n_in; ++i) {                                      /* lib.rs:605-606 */
                float v = feats[w * (size_t)n->n_in + i];
                if (keep && !keep[w * (size_t)n->n_in + i]) v = 0.f;
                xr[i] = v;
                if (v != 0.f) all_zero = 0;
            }
            if (all_zero) continue;                                                  /* lib.rs:607-609 */
            float* t = tb + nb * (size_t)n->n_out;
            for (int j = 0; j < n->n_out; ++j) t[j] = 0.f;
            if (labels[w] < (uint32_t)n->n_out) t[labels[w]] = 1.f;                  /* lib.rs:592-595 */
            so_forward(n, xr, 1, p);                                                 /* lib.rs:610 */
            float l = 0.f;
            for (int j = 0; j < n->n_out; ++j) l += t[j] * logf(fmaxf(p[j], 1e-12f)); /* lib.rs:611-615 */
            loss += (double)(-l);
            ++count; ++nb;
        }
        so_train_batch(n, xb, nb, tb, 1, lr);                                        /* lib.rs:620 */
    }
    free(xb); free(tb); free(p);
    *loss_sum = loss;
    return count;
}

/* lib.rs:1389-1402: counts[argmax_last(p)] += 1 when p_max >= threshold. counts: [n_out] u64, zeroed by the caller. */
void so_identify_counts(const so_net* n, const float* feats, size_t n_win, float threshold, uint64_t* counts) {
    float* p = (float*)malloc(sizeof(float) * (size_t)n->n_out);
    for (size_t w = 0; w < n_win; ++w) {
        so_forward(n, feats + w * (size_t)n->n_in, 1, p);
        int best = 0; float bv = p[0];
        for (int j = 1; j < n->n_out; ++j) if (p[j] >= bv) { bv = p[j]; best = j; } /* last max wins */
        if (bv >= threshold) counts[best] += 1;
    }
    free(p);
}
