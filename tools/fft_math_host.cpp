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
    std::vector<float> sr(kHalf), si(kHalf);
    // stage A: for each n1, 20-point DFT over n2 of z[n1 + 20 n2], twiddle W400^{n1 k2}, store row k2*20 + n1
    for (int n1 = 0; n1 < kR; ++n1) {
        float re[20], im[20];
        for (int n2 = 0; n2 < kR; ++n2) {
            const int n = n1 + kR * n2;
            re[n2] = float(frame[2 * n]);
            im[n2] = float(frame[2 * n + 1]);
        }
        dft20(re, im);
        for (int k2 = 0; k2 < kR; ++k2) {
            const float wr = tw400[(n1 * kR + k2) * 2], wi = tw400[(n1 * kR + k2) * 2 + 1];
            sr[k2 * kR + n1] = re[k2] * wr - im[k2] * wi;
            si[k2 * kR + n1] = re[k2] * wi + im[k2] * wr;
        }
    }
    // stage B: for each k2, 20-point DFT over n1 in place: row k2*20 + k1 holds Z[k2 + 20 k1]
    for (int k2 = 0; k2 < kR; ++k2) {
        float re[20], im[20];
        for (int n1 = 0; n1 < kR; ++n1) { re[n1] = sr[k2 * kR + n1]; im[n1] = si[k2 * kR + n1]; }
        dft20(re, im);
        for (int k1 = 0; k1 < kR; ++k1) { sr[k2 * kR + k1] = re[k1]; si[k2 * kR + k1] = im[k1]; }
    }
    // real split + power
    {
        const float zr = sr[row_of_bin(0)], zi = si[row_of_bin(0)];
        const float a = zr + zi, b = zr - zi;
        power4[0] = 4.f * a * a;
        power4[400] = 4.f * b * b;
    }
    for (int k = 1; k <= 200; ++k) {
        const int ra = row_of_bin(k), rb = row_of_bin(kHalf - k);
        float pk, pmk;
        split_pair_power(sr[ra], si[ra], sr[rb], si[rb], tw800[2 * k], tw800[2 * k + 1], pk, pmk);
        power4[k] = pk;
        power4[kHalf - k] = pmk;  // k == 200 writes the same value twice
    }
}

void host_dft20(float* re, float* im) {
    float r[20], i[20];
    std::memcpy(r, re, sizeof r);
    std::memcpy(i, im, sizeof i);
    dft20(r, i);
    std::memcpy(re, r, sizeof r);
    std::memcpy(im, i, sizeof i);
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
