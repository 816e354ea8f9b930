"""Generates the committed fixtures under tests/golden/ (run in the build container; needs torchaudio + scipy).

The reference (Rust) cannot be executed here, and its own tests hold no golden vectors for this path
(streamz-rs/src/lib.rs:1827-1865), so the fixtures are produced by an INDEPENDENT pipeline assembled from
third-party library pieces -- torchaudio's Slaney mel bank, scipy's rfft and scipy's DCT-II -- not by oracle/.
tests/test_oracle.py requires the oracle to reproduce them; the -m gpu tests require the CUDA path to.
"""
import os
import sys

import numpy as np
import scipy.fft
import torchaudio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def clip(seed: int, n: int) -> np.ndarray:
    """Small deterministic test signal that does not come from the oracle's generator: chirp + tones + noise."""
    r = np.random.default_rng(seed)
    t = np.arange(n) / 44100.0
    x = 0.3 * np.sin(2 * np.pi * (200 + 1500 * t) * t) + 0.2 * np.sin(2 * np.pi * 3300 * t + 1.0)
    x += 0.1 * np.sin(2 * np.pi * 9100 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 5 * t)) + 0.02 * r.standard_normal(n)
    return np.round(x / np.abs(x).max() * 0.6 * 32767).astype(np.int16)


def independent_features(s: np.ndarray) -> np.ndarray:
    mel = torchaudio.functional.melscale_fbanks(401, 0.0, 22050.0, 26, 44100, norm="slaney", mel_scale="slaney").T.numpy()
    n = 0 if len(s) < 800 else (len(s) - 800) // 400 + 1
    frames = np.stack([s[i * 400: i * 400 + 800] for i in range(n)]).astype(np.float64) / 32767.0
    power = np.abs(scipy.fft.rfft(frames, axis=1)) ** 2
    e = np.log(np.maximum(power @ mel.astype(np.float64).T, 1e-12))
    c = scipy.fft.dct(e, type=2, axis=1, norm=None)[:, :20] / 2.0      # scipy's DCT-II is 2x the unscaled sum
    def delta(x):
        i = np.arange(len(x))
        return (x[np.minimum(i + 1, len(x) - 1)] - x[np.maximum(i - 1, 0)]) / 2.0
    d1 = delta(c); d2 = delta(d1)
    v = np.concatenate([c, d1, d2], axis=1)
    mu = v.mean(axis=1, keepdims=True)
    sd = np.maximum(np.sqrt(((v - mu) ** 2).mean(axis=1, keepdims=True)), 1e-6)
    return (v - mu) / sd


def main():
    mel = torchaudio.functional.melscale_fbanks(401, 0.0, 22050.0, 26, 44100, norm="slaney", mel_scale="slaney").T.numpy()
    np.save(os.path.join(OUT, "mel_torchaudio_26x401.npy"), mel.astype(np.float32))
    for name, seed, n in (("a", 1, 800 + 400 * 40 + 123), ("b", 2, 800 + 400 * 3)):
        s = clip(seed, n)
        np.save(os.path.join(OUT, f"clip_{name}_i16.npy"), s)
        np.save(os.path.join(OUT, f"clip_{name}_features_f64.npy"), independent_features(s))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    sys.exit(main())
