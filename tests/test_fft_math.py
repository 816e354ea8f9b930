"""The register-level arithmetic of the CUDA front end (streamz_b200/csrc/fft_math.cuh + tables.hpp), compiled for the
host by tools/fft_math_host.cpp, against numpy: same code, same index maps, no GPU needed."""
import numpy as np

from conftest import P


def test_dft20_matches_numpy(fftmath):
    r = np.random.default_rng(0)
    for _ in range(20):
        re, im = r.standard_normal(20).astype(np.float32), r.standard_normal(20).astype(np.float32)
        want = np.fft.fft(re.astype(np.float64) + 1j * im.astype(np.float64))
        fftmath.host_dft20(P(re), P(im))
        assert np.abs(re - want.real).max() < 5e-6 and np.abs(im - want.imag).max() < 5e-6


def test_frame_power_matches_rfft(fftmath, oracle):
    clip = oracle.synth_clip(2, 4, 0.2)
    square = np.where(np.arange(800) % 50 < 25, 32767, -32768).astype(np.int16)     # extreme amplitudes
    for frame in (clip[:800].copy(), clip[1000:1800].copy(), square, np.zeros(800, np.int16)):
        p4 = np.zeros(401, np.float32)
        fftmath.host_frame_power(P(frame), P(p4))
        spec = np.fft.rfft(frame.astype(np.float64))
        want = 4.0 * (spec.real ** 2 + spec.imag ** 2)
        assert np.abs(p4 - want).max() <= 2e-6 * max(want.max(), 1.0)
