// Host-side construction of the constant tables used by the front-end kernels (all computed in double, stored f32).
// Reference: FeatureExtractor::new, streamz-rs/src/lib.rs:239-257 (mel bank 26x401, FFT-800 plan, DCT-II-26 plan).
#pragma once
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#include "fft_math.cuh"

namespace szb {

struct MelCsr {
    int start[kMels];            // first non-zero bin of filter m
    int len[kMels];              // number of non-zero bins
    int off[kMels + 1];          // offset of filter m's weights in w
    std::vector<float> w;        // weights as the reference holds them (f32)
};

inline double hz_to_mel_slaney(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
inline double mel_to_hz_slaney(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// Dense 26 x 401 bank with the semantics of mel::<f32>(44100, 800, Some(26), None, None, false, One) (lib.rs:240-248):
// librosa-style Slaney scale, fmin 0, fmax sr/2, rows scaled by 2 / (f[m+2] - f[m]).
inline std::vector<float> mel_filterbank_dense() {
    const double sr = 44100.0;
    std::vector<double> edges(kMels + 2);
    const double m_lo = hz_to_mel_slaney(0.0), m_hi = hz_to_mel_slaney(sr / 2.0);
    for (int i = 0; i < kMels + 2; ++i) {
        // numpy.linspace semantics: start + i * step
        double step = (m_hi - m_lo) / double(kMels + 1);
        edges[i] = mel_to_hz_slaney(i == kMels + 1 ? m_hi : m_lo + i * step);
    }
    std::vector<float> fb(size_t(kMels) * kBins, 0.f);
    for (int m = 0; m < kMels; ++m) {
        const double enorm = 2.0 / (edges[m + 2] - edges[m]);
        for (int k = 0; k < kBins; ++k) {
            const double f = k * (sr / kWindow);
            const double lower = -(edges[m] - f) / (edges[m + 1] - edges[m]);
            const double upper = (edges[m + 2] - f) / (edges[m + 2] - edges[m + 1]);
            const double v = std::fmax(0.0, std::fmin(lower, upper)) * enorm;
            fb[size_t(m) * kBins + k] = float(v);
        }
    }
    return fb;
}

inline MelCsr mel_filterbank_csr(const std::vector<float>& dense) {
    MelCsr c;
    int off = 0;
    for (int m = 0; m < kMels; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < kBins; ++k)
            if (dense[size_t(m) * kBins + k] != 0.f) { if (first < 0) first = k; last = k; }
        c.start[m] = first < 0 ? 0 : first;
        c.len[m] = first < 0 ? 0 : last - first + 1;
        c.off[m] = off;
        for (int k = 0; k < c.len[m]; ++k) c.w.push_back(dense[size_t(m) * kBins + c.start[m] + k]);
        off += c.len[m];
    }
    c.off[kMels] = off;
    return c;
}

// Unscaled DCT-II rows 0..19 over 26 inputs (rustdct process_dct2, lib.rs:313-314).
inline std::vector<float> dct2_rows() {
    std::vector<float> d(size_t(kMfcc) * kMels);
    for (int j = 0; j < kMfcc; ++j)
        for (int m = 0; m < kMels; ++m) d[size_t(j) * kMels + m] = float(std::cos(M_PI * (m + 0.5) * j / kMels));
    return d;
}

// W_400^{n1 k2}, n1 row, k2 column, interleaved (re, im).
inline std::vector<float> twiddles_400() {
    std::vector<float> t(size_t(kR) * kR * 2);
    for (int n1 = 0; n1 < kR; ++n1)
        for (int k2 = 0; k2 < kR; ++k2) {
            const double a = -2.0 * M_PI * double((n1 * k2) % kHalf) / double(kHalf);
            t[(size_t(n1) * kR + k2) * 2 + 0] = float(std::cos(a));
            t[(size_t(n1) * kR + k2) * 2 + 1] = float(std::sin(a));
        }
    return t;
}

// W_800^k for k = 0..200, interleaved (re, im).
inline std::vector<float> twiddles_800_half() {
    std::vector<float> t(size_t(201) * 2);
    for (int k = 0; k <= 200; ++k) {
        const double a = -2.0 * M_PI * double(k) / double(kWindow);
        t[size_t(k) * 2 + 0] = float(std::cos(a));
        t[size_t(k) * 2 + 1] = float(std::sin(a));
    }
    return t;
}

// ---- resampler (this repo's polyphase specification; see DESIGN.md "Resampler") ----
constexpr int kResTaps = 16;
constexpr double kResBeta = 8.0;
constexpr double kResRolloff = 0.93;

inline double bessel_i0(double x) {
    double term = 1.0, total = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 64; ++k) { term = term * q / (double(k) * double(k)); total += term; }
    return total;
}

inline void resample_ratio(uint32_t rate, uint32_t& L, uint32_t& M) {
    const uint32_t g = std::gcd(rate, 44100u);
    L = 44100u / g;
    M = rate / g;
}

// c[L][T]: tap t of phase p weighs input sample i0 - (T/2 - 1) + t; Kaiser-windowed sinc at u = p/L + T/2 - 1 - t,
// each phase normalised to unit DC gain.
inline std::vector<float> resample_taps(uint32_t rate) {
    uint32_t L, M;
    resample_ratio(rate, L, M);
    const int T = kResTaps;
    const double fc = 0.5 * kResRolloff * std::fmin(1.0, double(L) / double(M));
    const double i0b = bessel_i0(kResBeta);
    std::vector<float> c(size_t(L) * T);
    std::vector<double> row(T);
    for (uint32_t p = 0; p < L; ++p) {
        double sum = 0.0;
        for (int t = 0; t < T; ++t) {
            const double u = double(p) / double(L) + double(T / 2 - 1) - double(t);
            const double x = u / (T / 2.0);
            double win = 0.0;
            if (std::fabs(x) <= 1.0) win = bessel_i0(kResBeta * std::sqrt(std::fmax(0.0, 1.0 - x * x))) / i0b;
            const double a = 2.0 * fc * u;
            const double sinc = a == 0.0 ? 1.0 : std::sin(M_PI * a) / (M_PI * a);
            row[t] = 2.0 * fc * sinc * win;
            sum += row[t];
        }
        for (int t = 0; t < T; ++t) c[size_t(p) * T + t] = float(row[t] / sum);
    }
    return c;
}

}  // namespace szb
