// CPU harness around streamz_b200/csrc/fft_math.cuh + tables.hpp: runs the SAME register-level arithmetic and the
// same index maps the CUDA front-end kernel uses, one frame at a time, so that tests/test_fft_math.py can check them
// against the oracle without a GPU.  Not part of the product library.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../streamz_b200/csrc/tables.hpp"

using namespace szb;

extern "C" {

// frame: 800 i16 samples.  power4: 401 floats = 4 * 32767^2 * |X[k]|^2 of the reference (unscaled integer-valued input,
// unscaled split), exactly what the kernel's power stage produces.
void host_frame_power(const int16_t* frame, float* power4) {
    static const std::vector<float> tw400 = twiddles_400();
    static const std::vector<float> tw800 = twiddles_800_half();
    std::vector<cpx> S(kHalf);
    // stage A: for each n1, 20-point DFT over n2 of z[n1 + 20 n2], twiddle W400^{n1 k2}, store row k2*20 + n1
    for (int n1 = 0; n1 < kR; ++n1) {
        cpx z[20];
        for (int n2 = 0; n2 < kR; ++n2) {
            const int n = n1 + kR * n2;
            z[n2] = cpack(float(frame[2 * n]), float(frame[2 * n + 1]));
        }
        dft20(z);
        S[n1] = z[0];
        for (int k2 = 1; k2 < kR; ++k2) {
            const float* w = &tw400[(n1 * kR + k2) * 2];
            S[k2 * kR + n1] = cmul_tw(z[k2], w[0], w[1]);
        }
    }
    // stage B: for each k2, 20-point DFT over n1 in place: row k2*20 + k1 holds Z[k2 + 20 k1]
    for (int k2 = 0; k2 < kR; ++k2) {
        cpx z[20];
        for (int n1 = 0; n1 < kR; ++n1) z[n1] = S[k2 * kR + n1];
        dft20(z);
        for (int k1 = 0; k1 < kR; ++k1) S[k2 * kR + k1] = z[k1];
    }
    // real split + power
    {
        const float zr = cre(S[row_of_bin(0)]), zi = cim(S[row_of_bin(0)]);
        const float a = zr + zi, b = zr - zi;
        power4[0] = 4.f * a * a;
        power4[400] = 4.f * b * b;
    }
    for (int k = 1; k <= 200; ++k) {
        const float* w = &tw800[k * 2];
        float pk, pmk;
        split_pair_power(S[row_of_bin(k)], S[row_of_bin(kHalf - k)], w[0], w[1], pk, pmk);
        power4[k] = pk;
        power4[kHalf - k] = pmk;  // k == 200 writes the same value twice
    }
}

void host_dft20(float* re, float* im) {
    cpx z[20];
    for (int i = 0; i < 20; ++i) z[i] = cpack(re[i], im[i]);
    dft20(z);
    for (int i = 0; i < 20; ++i) { re[i] = cre(z[i]); im[i] = cim(z[i]); }
}

void host_mel_dense(float* out /*26*401*/) {
    auto d = mel_filterbank_dense();
    std::memcpy(out, d.data(), d.size() * sizeof(float));
}
void host_dct_rows(float* out /*20*26*/) {
    auto d = dct2_rows();
    std::memcpy(out, d.data(), d.size() * sizeof(float));
}
uint32_t host_resample_taps(uint32_t rate, float* out /*L*16, may be null*/) {
    uint32_t L, M;
    resample_ratio(rate, L, M);
    if (out) {
        auto c = resample_taps(rate);
        std::memcpy(out, c.data(), c.size() * sizeof(float));
    }
    return L;
}
}
