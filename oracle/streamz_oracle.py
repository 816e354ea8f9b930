"""CPU oracle for the StreamZ hot path (numpy restatement of the reference arithmetic).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``streamz_b200``) never imports anything from ``oracle/`` and fails loudly when its CUDA library is missing.

PARITY STATUS: **parity unpinned** at the third-party-crate boundaries.  The reference is Rust and cannot be built
in this environment (no cargo/rustc; crate sources not on disk), and its own tests (lib.rs:1827-1865) pin no
numbers for this path.  The arithmetic that lives in un-vendored crates is restated from their published
definitions:
  * rustfft 6.4.0   ``plan_fft_forward(800)``      -> unnormalised forward DFT (lib.rs:249-250, 296)
  * rustdct 0.7.1   ``plan_dct2(26)``              -> unscaled DCT-II  sum x_n cos(pi (n+1/2) k / 26) (lib.rs:251-252, 313);
                                                     the scale cancels in the per-window z-score (lib.rs:328-340)
  * mel_filter 0.1.1 ``mel(44100,800,26,None,None,false,One)`` -> librosa ``filters.mel`` semantics, Slaney scale,
                                                     Slaney area normalisation (lib.rs:240-248); cross-checked here
                                                     against torchaudio ``melscale_fbanks(norm='slaney',mel_scale='slaney')``
  * rubato 0.13.0   ``FftFixedInOut``              -> NOT restated: the reference's single-call use of a pooled,
                                                     stateful resampler is ill-defined.  The resampler
                                                     here is this repo's own polyphase specification.
Everything that is in the reference's own source is followed line by line and cited per function.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np

# lib.rs:25-36
DEFAULT_SAMPLE_RATE = 44100
WINDOW_SIZE = 800
HOP_SIZE = WINDOW_SIZE // 2  # lib.rs:288
N_MELS = 26
MFCC_SIZE = 20
FEATURE_SIZE = 3 * MFCC_SIZE
DEFAULT_DROPOUT = 0.2
I16_MAX = 32767.0

# ----------------------------------------------------------------------------------------------------------------------
# Front end tables (lib.rs:239-257)
# ----------------------------------------------------------------------------------------------------------------------


def _hz_to_mel_slaney(f):
    """Slaney mel scale (librosa ``hz_to_mel(htk=False)``): linear below 1 kHz, log above."""
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    lin = f / f_sp
    with np.errstate(divide="ignore", invalid="ignore"):
        log = min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, log, lin)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr: int = DEFAULT_SAMPLE_RATE, n_fft: int = WINDOW_SIZE, n_mels: int = N_MELS,
                   dtype=np.float32) -> np.ndarray:
    """26 x 401 triangular filterbank; restates ``mel::<f32>(44100, 800, Some(26), None, None, false, One)``
    (lib.rs:240-248) with librosa semantics: fmin 0, fmax sr/2, Slaney scale, rows scaled by 2/(f[m+2]-f[m])."""
    n_bins = n_fft // 2 + 1
    fft_freqs = np.arange(n_bins, dtype=np.float64) * (sr / n_fft)
    edges = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(sr / 2.0), n_mels + 2))
    fdiff = np.diff(edges)
    ramps = edges[:, None] - fft_freqs[None, :]
    fb = np.zeros((n_mels, n_bins), dtype=np.float64)
    for m in range(n_mels):
        lower = -ramps[m] / fdiff[m]
        upper = ramps[m + 2] / fdiff[m + 1]
        fb[m] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (edges[2:n_mels + 2] - edges[:n_mels])
    fb *= enorm[:, None]
    return fb.astype(dtype)


def dct2_matrix(n_in: int = N_MELS, n_out: int = MFCC_SIZE, dtype=np.float64) -> np.ndarray:
    """Unscaled DCT-II rows 0..n_out-1: D[j, m] = cos(pi (m + 1/2) j / n_in)  (rustdct ``process_dct2``,
    lib.rs:313-314; truncation to 20 coefficients lib.rs:314)."""
    j = np.arange(n_out, dtype=np.float64)[:, None]
    m = np.arange(n_in, dtype=np.float64)[None, :]
    return np.cos(np.pi * (m + 0.5) * j / n_in).astype(dtype)


def n_windows(n_samples: int) -> int:
    """Number of 800-sample frames at hop 400 (lib.rs:289-291, 317)."""
    if n_samples < WINDOW_SIZE:
        return 0
    return (n_samples - WINDOW_SIZE) // HOP_SIZE + 1


# ----------------------------------------------------------------------------------------------------------------------
# Front end (lib.rs:167-169, 212-228, 279-345)
# ----------------------------------------------------------------------------------------------------------------------


def i16_to_f32(samples: np.ndarray, dtype=np.float32) -> np.ndarray:
    """lib.rs:167-169: ``sample as f32 / i16::MAX as f32`` (true division by 32767)."""
    return (np.asarray(samples).astype(dtype) / dtype(I16_MAX)).astype(dtype)


def downmix_to_mono(samples: np.ndarray, channels: int) -> np.ndarray:
    """lib.rs:172-183: i32 sum of each chunk of ``channels`` samples, integer division truncating toward zero by
    ``channels`` (a trailing partial chunk is divided by ``channels`` too), cast to i16."""
    s = np.asarray(samples, dtype=np.int16)
    if channels <= 1:
        return s.copy()
    n = (len(s) + channels - 1) // channels
    padded = np.zeros(n * channels, dtype=np.int32)
    padded[:len(s)] = s
    sums = padded.reshape(n, channels).sum(axis=1)
    q = np.abs(sums) // channels
    return (np.sign(sums) * q).astype(np.int16)


def mfcc_frames(samples: np.ndarray, precision: str = "f64") -> np.ndarray:
    """Frame loop of ``window_samples_with_plan`` (lib.rs:285-319): returns ``[n, 20]`` MFCCs.

    precision 'f64': mathematical definition evaluated in float64 (the parity target).
    precision 'f32': the same operations carried out in float32 in the reference's order (complex FFT of the real
                     frame, ``norm_sqr``, dense mel dot, ``ln``), used to measure the reference-precision noise floor.
    """
    s = np.asarray(samples, dtype=np.int16)
    n = n_windows(len(s))
    if n == 0:
        return np.zeros((0, MFCC_SIZE), dtype=np.float64 if precision == "f64" else np.float32)
    idx = np.arange(n)[:, None] * HOP_SIZE + np.arange(WINDOW_SIZE)[None, :]
    if precision == "f64":
        x = s.astype(np.float64)[idx] / I16_MAX
        spec = np.fft.fft(x, axis=1)[:, : WINDOW_SIZE // 2 + 1]            # lib.rs:296
        power = spec.real ** 2 + spec.imag ** 2                              # lib.rs:297-301
        mel = mel_filterbank(dtype=np.float32).astype(np.float64)           # reference holds the bank in f32
        energies = np.log(np.maximum(power @ mel.T, 1e-12))                  # lib.rs:303-310
        return energies @ dct2_matrix().T                                    # lib.rs:312-315
    import scipy.fft

    x = i16_to_f32(s)[idx].astype(np.complex64)
    spec = scipy.fft.fft(x, axis=1)[:, : WINDOW_SIZE // 2 + 1]
    assert spec.dtype == np.complex64
    power = (spec.real * spec.real + spec.imag * spec.imag).astype(np.float32)
    mel = mel_filterbank(dtype=np.float32)
    sums = np.zeros((n, N_MELS), dtype=np.float32)
    for j in range(power.shape[1]):                                          # sequential f32 accumulation, lib.rs:305-308
        sums += mel[:, j][None, :] * power[:, j][:, None]
    energies = np.log(np.maximum(sums, np.float32(1e-12))).astype(np.float32)
    return (energies.astype(np.float64) @ dct2_matrix().T).astype(np.float32)


def add_deltas(mfcc: np.ndarray) -> np.ndarray:
    """lib.rs:212-228: d[i] = (x[min(i+1, n-1)] - x[max(i-1, 0)]) / 2."""
    n = mfcc.shape[0]
    if n == 0:
        return mfcc.copy()
    nxt = mfcc[np.minimum(np.arange(n) + 1, n - 1)]
    prv = mfcc[np.maximum(np.arange(n) - 1, 0)]
    return ((nxt - prv) / mfcc.dtype.type(2.0)).astype(mfcc.dtype)


def normalise_windows(frames: np.ndarray) -> np.ndarray:
    """lib.rs:328-340: mean and population variance over the 60 values, std = max(sqrt(var), 1e-6)."""
    if frames.shape[0] == 0:
        return frames.copy()
    t = frames.dtype.type
    mean = frames.sum(axis=1, dtype=frames.dtype) / t(frames.shape[1])
    d = frames - mean[:, None]
    var = (d * d).sum(axis=1, dtype=frames.dtype) / t(frames.shape[1])
    std = np.maximum(np.sqrt(var), t(1e-6))
    return (d / std[:, None]).astype(frames.dtype)


def features_from_mfcc(base: np.ndarray) -> np.ndarray:
    """lib.rs:321-342: [mfcc | delta | delta-delta] then per-window z-score."""
    d1 = add_deltas(base)
    d2 = add_deltas(d1)
    return normalise_windows(np.concatenate([base, d1, d2], axis=1))


def extract(samples: np.ndarray, precision: str = "f64") -> np.ndarray:
    """``FeatureExtractor::extract`` (lib.rs:261-263 -> 279-345): mono i16 @ 44.1 kHz -> ``[n, 60]``."""
    return features_from_mfcc(mfcc_frames(samples, precision))


def extract_direct_dft(samples: np.ndarray) -> np.ndarray:
    """Independent check of :func:`extract`: explicit O(N^2) DFT matrix instead of an FFT (float64)."""
    s = np.asarray(samples, dtype=np.int16).astype(np.float64) / I16_MAX
    n = n_windows(len(s))
    if n == 0:
        return np.zeros((0, FEATURE_SIZE))
    t = np.arange(WINDOW_SIZE)[:, None]
    k = np.arange(WINDOW_SIZE // 2 + 1)[None, :]
    ang = -2.0 * np.pi * ((t * k) % WINDOW_SIZE) / WINDOW_SIZE
    wr, wi = np.cos(ang), np.sin(ang)
    base = np.zeros((n, MFCC_SIZE))
    mel = mel_filterbank(dtype=np.float32).astype(np.float64)
    dct = dct2_matrix()
    for w in range(n):
        fr = s[w * HOP_SIZE: w * HOP_SIZE + WINDOW_SIZE]
        re, im = fr @ wr, fr @ wi
        e = np.log(np.maximum(mel @ (re * re + im * im), 1e-12))
        base[w] = dct @ e
    return features_from_mfcc(base)


# ----------------------------------------------------------------------------------------------------------------------
# Resampler: THIS REPO'S polyphase specification (the reference's rubato call is ill-defined).
# Only the quantisation contract around it follows the reference: output length floor(n*44100/rate) (lib.rs:196),
# clamp to [-32768, 32767] and truncate toward zero (lib.rs:205-208), rate == 44100 is the identity (lib.rs:187-189).
# ----------------------------------------------------------------------------------------------------------------------

RESAMPLE_TAPS = 16          # taps per phase (input samples spanned)
RESAMPLE_KAISER_BETA = 8.0
RESAMPLE_ROLLOFF = 0.93


def resample_ratio(rate: int) -> Tuple[int, int]:
    g = math.gcd(int(rate), DEFAULT_SAMPLE_RATE)
    return DEFAULT_SAMPLE_RATE // g, int(rate) // g  # L (up), M (down)


def _bessel_i0(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    term = np.ones_like(x)
    total = np.ones_like(x)
    q = x * x / 4.0
    for k in range(1, 64):
        term = term * q / (k * k)
        total = total + term
    return total


def resample_taps(rate: int) -> np.ndarray:
    """Polyphase table ``c[L, T]`` (float32).  Output j sits at input position j*M/L = i0 + p/L; tap t weighs input
    sample i0 - (T/2 - 1) + t with a Kaiser-windowed sinc evaluated at u = p/L + T/2 - 1 - t; each phase is
    normalised to unit DC gain."""
    L, M = resample_ratio(rate)
    T = RESAMPLE_TAPS
    fc = 0.5 * RESAMPLE_ROLLOFF * min(1.0, L / M)
    p = np.arange(L, dtype=np.float64)[:, None] / L
    t = np.arange(T, dtype=np.float64)[None, :]
    u = p + (T // 2 - 1) - t
    x = u / (T / 2.0)
    win = np.where(np.abs(x) <= 1.0, _bessel_i0(RESAMPLE_KAISER_BETA * np.sqrt(np.maximum(0.0, 1.0 - x * x))), 0.0)
    win = win / _bessel_i0(np.array(RESAMPLE_KAISER_BETA))
    c = 2.0 * fc * np.sinc(2.0 * fc * u) * win
    c = c / c.sum(axis=1, keepdims=True)
    return c.astype(np.float32)


def resample_out_len(n_in: int, rate: int) -> int:
    return (int(n_in) * DEFAULT_SAMPLE_RATE) // int(rate)  # lib.rs:196


def resample_to_44100(samples: np.ndarray, rate: int, taps: Optional[np.ndarray] = None) -> np.ndarray:
    """i16 @ rate -> i16 @ 44.1 kHz.  acc = fma(c[p][t], float(x[i]), acc) for t = 0..T-1 in float32, inputs outside
    the clip are zero; out = trunc(clamp(acc, -32768, 32767))."""
    s = np.asarray(samples, dtype=np.int16)
    if rate == DEFAULT_SAMPLE_RATE:
        return s.copy()
    L, M = resample_ratio(rate)
    T = RESAMPLE_TAPS
    c = resample_taps(rate) if taps is None else np.asarray(taps, dtype=np.float32).reshape(L, T)
    n_out = resample_out_len(len(s), rate)
    if n_out == 0:
        return np.zeros(0, dtype=np.int16)
    j = np.arange(n_out, dtype=np.int64)
    pos = j * M
    i0 = pos // L
    p = pos % L
    xf = np.concatenate([np.zeros(T, np.float32), s.astype(np.float32), np.zeros(T + 1, np.float32)])
    acc = np.zeros(n_out, dtype=np.float32)
    for t in range(T):
        xi = xf[i0 - (T // 2 - 1) + t + T]
        # float32 fma emulated through float64: the product of two float32 is exact in float64
        acc = (c[p, t].astype(np.float64) * xi.astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
    y = np.clip(acc, np.float32(-32768.0), np.float32(32767.0))
    return np.trunc(y).astype(np.int16)


# ----------------------------------------------------------------------------------------------------------------------
# SimpleNeuralNet (lib.rs:745-1060)
# ----------------------------------------------------------------------------------------------------------------------


class Net:
    """Weights of ``SimpleNeuralNet`` (lib.rs:745-751): w1[I,H1] b1[H1] w2[H1,H2] b2[H2] w3[H2,C] b3[C], row-major."""

    def __init__(self, w1, b1, w2, b2, w3, b3, dtype=np.float32):
        self.dtype = dtype
        self.w1, self.b1, self.w2, self.b2, self.w3, self.b3 = (np.array(a, dtype=dtype) for a in (w1, b1, w2, b2, w3, b3))

    @classmethod
    def init(cls, n_in, h1, h2, n_out, seed=0, dtype=np.float32):
        """lib.rs:767-790: weights U(-0.5, 0.5), zero biases (the RNG stream itself is not reproducible: thread_rng)."""
        r = np.random.default_rng(seed)
        u = lambda *s: r.uniform(-0.5, 0.5, size=s).astype(np.float32)
        return cls(u(n_in, h1), np.zeros(h1), u(h1, h2), np.zeros(h2), u(h2, n_out), np.zeros(n_out), dtype=dtype)

    def copy(self, dtype=None):
        return Net(self.w1, self.b1, self.w2, self.b2, self.w3, self.b3, dtype=dtype or self.dtype)

    @property
    def n_out(self):
        return self.b3.shape[0]

    def params(self):
        return [self.w1, self.b1, self.w2, self.b2, self.w3, self.b3]


def _hidden(net: Net, x: np.ndarray):
    a1 = x @ net.w1 + net.b1
    h1 = np.where(a1 > 0, a1, 0).astype(net.dtype)       # lib.rs:882 strict > 0
    h2 = np.tanh(h1 @ net.w2 + net.b2).astype(net.dtype)  # lib.rs:883
    return a1, h1, h2


def forward(net: Net, x: np.ndarray) -> np.ndarray:
    """``SimpleNeuralNet::forward`` (lib.rs:880-891), batched over rows of ``x``: max-subtracted softmax."""
    x = np.atleast_2d(np.asarray(x, dtype=net.dtype))
    _, _, h2 = _hidden(net, x)
    z = h2 @ net.w3 + net.b3
    e = np.exp(z - z.max(axis=1, keepdims=True))
    return (e / e.sum(axis=1, keepdims=True)).astype(net.dtype)


def embed(net: Net, x: np.ndarray) -> np.ndarray:
    """``embed`` (lib.rs:895-900): second hidden layer (ReLU, tanh)."""
    return _hidden(net, np.atleast_2d(np.asarray(x, dtype=net.dtype)))[2]


def forward_embedding(net: Net, x: np.ndarray) -> np.ndarray:
    """``forward_embedding`` (lib.rs:1073-1079): ReLU on both hidden layers."""
    x = np.atleast_2d(np.asarray(x, dtype=net.dtype))
    h1 = np.maximum(x @ net.w1 + net.b1, 0)
    return np.maximum(h1 @ net.w2 + net.b2, 0).astype(net.dtype)


def gradients(net: Net, x: np.ndarray, targets: np.ndarray):
    """Summed (not averaged) gradients of lib.rs:1013-1045 for rows ``x`` and per-row target vectors ``targets``."""
    x = np.atleast_2d(np.asarray(x, dtype=net.dtype))
    a1, h1, h2 = _hidden(net, x)
    z = h2 @ net.w3 + net.b3
    e = np.exp(z - z.max(axis=1, keepdims=True))
    p = e / e.sum(axis=1, keepdims=True)
    d3 = (p - targets).astype(net.dtype)                  # lib.rs:1028
    g_w3 = h2.T @ d3                                      # lib.rs:1029-1032
    g_b3 = d3.sum(axis=0)
    d2 = (d3 @ net.w3.T) * (1 - h2 * h2)                  # lib.rs:1034
    g_w2 = h1.T @ d2
    g_b2 = d2.sum(axis=0)
    d1 = (d2 @ net.w2.T) * (a1 > 0)                       # lib.rs:1039-1040
    g_w1 = x.T @ d1
    g_b1 = d1.sum(axis=0)
    return [g.astype(net.dtype) for g in (g_w1, g_b1, g_w2, g_b2, g_w3, g_b3)], p.astype(net.dtype)


def one_hot(labels: np.ndarray, n_out: int, dtype=np.float32) -> np.ndarray:
    """Target rows as built at lib.rs:592-595: all-zero when ``label >= n_out``."""
    labels = np.asarray(labels, dtype=np.int64)
    t = np.zeros((labels.shape[0], n_out), dtype=dtype)
    ok = labels < n_out
    t[np.nonzero(ok)[0], labels[ok]] = 1
    return t


def train_batch(net: Net, x: np.ndarray, targets: np.ndarray, lr: float) -> None:
    """``train_batch`` (lib.rs:1002-1060): mean-gradient SGD step in place; ``targets`` is ``[C]`` (shared by the
    batch, as the reference) or ``[B, C]``.  Empty batch is a no-op (lib.rs:1003-1005)."""
    x = np.atleast_2d(np.asarray(x, dtype=net.dtype))
    if x.shape[0] == 0:
        return
    targets = np.asarray(targets, dtype=net.dtype)
    if targets.ndim == 1:
        targets = np.broadcast_to(targets, (x.shape[0], targets.shape[0]))
    grads, _ = gradients(net, x, targets)
    scale = net.dtype(lr) / net.dtype(x.shape[0])        # lib.rs:1047
    for w, g in zip(net.params(), grads):
        w -= (g * scale).astype(net.dtype)


def augment(samples: np.ndarray, shift: int, gain: float, noise_level: float, key: int) -> np.ndarray:
    """lib.rs:103-116 with the draws injected: out[i] = trunc(clamp(f32(s[(i+shift) % n]) * gain + noise_i * 32767)) where
    noise_i = (2 u_i - 1) * noise_level and u_i = 24 bits of splitmix64(key ^ i) (this repo's counter stream).  All
    products and the sum are rounded to float32 one at a time, like the Rust expression."""
    s = np.asarray(samples, dtype=np.int16)
    n = len(s)
    if n == 0:
        return s.copy()
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        u = (_splitmix64(np.uint64(key) ^ i) >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)
    f = np.float32
    noise = ((f(2.0) * u).astype(f) - f(1.0)).astype(f) * f(noise_level)
    val = (s[(np.arange(n) + int(shift)) % n].astype(f) * f(gain)).astype(f) + (noise.astype(f) * f(32767.0)).astype(f)
    return np.trunc(np.clip(val.astype(f), f(-32768.0), f(32767.0))).astype(np.int16)


# --- counter-based dropout stream (this repo's; the reference uses an unseeded thread_rng, lib.rs:123) ---------------

_M64 = (1 << 64) - 1


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
    return (x ^ (x >> np.uint64(31))).astype(np.uint64)


def dropout_keep_mask(seed: int, stream: int, rows: np.ndarray, n_feat: int, prob: float) -> np.ndarray:
    """keep[r, i] for window ids ``rows``: u = splitmix64(splitmix64(seed + stream * 0xD1B54A32D192ED03) ^ (row * 64 + i));
    r32 = u >> 40 (24 bits); drop when r32 * 2^-24 < prob, i.e. ``rng.gen::<f32>() < prob`` of lib.rs:125."""
    rows = np.asarray(rows, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = _splitmix64(np.array([(seed + stream * 0xD1B54A32D192ED03) & _M64], dtype=np.uint64))[0]
        ctr = rows[:, None] * np.uint64(64) + np.arange(n_feat, dtype=np.uint64)[None, :]
        u = _splitmix64(key ^ ctr)
    r = (u >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)
    if prob <= 0.0:
        return np.ones(r.shape, dtype=bool)
    return ~(r < np.float32(prob))


def train_epoch(net: Net, feats: np.ndarray, labels: np.ndarray, perm: np.ndarray, batch: int, lr: float,
                keep_mask: Optional[np.ndarray] = None) -> Tuple[float, int]:
    """One epoch of lib.rs:599-622 with the randomness injected: ``perm`` is the shuffled order (lib.rs:601),
    ``keep_mask[w]`` the dropout decisions of window ``w`` (lib.rs:606).  A window whose features are all exactly zero
    after dropout is skipped (lib.rs:607-609); loss uses the pre-update weights (lib.rs:610-617).  Returns
    (sum of losses, number of surviving windows).  ``labels`` may differ per window (superset of the reference)."""
    n_out = net.n_out
    loss_sum, count = 0.0, 0
    batch = max(1, int(batch))
    for s in range(0, len(perm), batch):
        idx = np.asarray(perm[s:s + batch], dtype=np.int64)
        x = feats[idx].astype(net.dtype).copy()
        if keep_mask is not None:
            x = np.where(keep_mask[idx], x, net.dtype(0))
        alive = ~np.all(x == 0, axis=1)
        x, lab = x[alive], np.asarray(labels)[idx][alive]
        if x.shape[0] == 0:
            continue
        t = one_hot(lab, n_out, net.dtype)
        p = forward(net, x)
        loss_sum += float(-(t * np.log(np.maximum(p, net.dtype(1e-12)))).sum())
        count += x.shape[0]
        train_batch(net, x, t, lr)
    return loss_sum, count


# ----------------------------------------------------------------------------------------------------------------------
# Aggregation (lib.rs:1285-1411)
# ----------------------------------------------------------------------------------------------------------------------


def argmax_last(p: np.ndarray) -> np.ndarray:
    """``max_by(partial_cmp)`` keeps the LAST maximal element (lib.rs:1393-1396)."""
    c = p.shape[1]
    return c - 1 - np.argmax(p[:, ::-1], axis=1)


def identify_counts(net: Net, feats: np.ndarray, threshold: float) -> np.ndarray:
    counts = np.zeros(net.n_out, dtype=np.int64)
    if feats.shape[0] == 0:
        return counts
    p = forward(net, feats)
    best = argmax_last(p)
    ok = p[np.arange(len(best)), best] >= net.dtype(threshold)   # lib.rs:1398
    np.add.at(counts, best[ok], 1)
    return counts


def speakers_from_counts(counts: Sequence[int]) -> list:
    """lib.rs:1403-1410: speakers with count > 0, stable sort by count descending (ties stay in index order)."""
    pairs = [(i, int(c)) for i, c in enumerate(counts) if c > 0]
    pairs.sort(key=lambda ic: -ic[1])
    return [i for i, _ in pairs]


def identify_speaker_list(net: Net, samples: np.ndarray, threshold: float, precision: str = "f64") -> list:
    return speakers_from_counts(identify_counts(net, extract(samples, precision).astype(net.dtype), threshold))


def identify_sums(net: Net, feats: np.ndarray) -> np.ndarray:
    """Per-class sum of window probabilities (lib.rs:1290-1297)."""
    if feats.shape[0] == 0:
        return np.zeros(net.n_out, dtype=net.dtype)
    return forward(net, feats).sum(axis=0, dtype=net.dtype)


def identify_speaker(net: Net, feats: np.ndarray) -> int:
    """lib.rs:1285-1303 on precomputed windows (0 when there are none)."""
    if feats.shape[0] == 0 or net.n_out == 0:
        return 0
    return int(argmax_last(identify_sums(net, feats)[None, :])[0])


def identify_speaker_with_threshold_feats(net: Net, feats: np.ndarray, threshold: float) -> Optional[int]:
    """lib.rs:1346-1377: None when C <= 1, no windows, or mean confidence below the threshold."""
    if net.n_out <= 1 or feats.shape[0] == 0:
        return None
    sums = identify_sums(net, feats)
    best = int(argmax_last(sums[None, :])[0])
    return best if sums[best] / net.dtype(feats.shape[0]) >= net.dtype(threshold) else None


# ----------------------------------------------------------------------------------------------------------------------
# Embeddings and cosine matching (lib.rs:131-139, 1413-1661)
# ----------------------------------------------------------------------------------------------------------------------


def normalize(v: np.ndarray) -> np.ndarray:
    """lib.rs:132-139: divide by the L2 norm when it exceeds 1e-6."""
    v = np.asarray(v).copy()
    norm = np.sqrt((v * v).sum(dtype=v.dtype))
    return v / norm if norm > 1e-6 else v


def embedding_mean(net: Net, feats: np.ndarray) -> np.ndarray:
    """extract_embedding_from_features (lib.rs:1453-1475)."""
    if feats.shape[0] == 0:
        return np.zeros(net.w2.shape[1], dtype=net.dtype)
    e = forward_embedding(net, feats)
    return normalize(e.sum(axis=0, dtype=net.dtype) / net.dtype(feats.shape[0]))


def embedding_median(net: Net, feats: np.ndarray, relu2: bool = True) -> np.ndarray:
    """median_embedding_from_features (relu2, lib.rs:1478-1500) / the reduction of extract_embedding (lib.rs:1418-1450):
    per-dimension median (mean of the two middle values for an even count), then normalize."""
    if feats.shape[0] == 0:
        return np.zeros(net.w2.shape[1], dtype=net.dtype)
    e = forward_embedding(net, feats) if relu2 else embed(net, feats)
    srt = np.sort(e, axis=0)
    n = e.shape[0]
    med = srt[n // 2] if n % 2 else (srt[n // 2 - 1] + srt[n // 2]) / net.dtype(2.0)
    return normalize(med.astype(net.dtype))


def cosine_similarity(a: np.ndarray, b: np.ndarray) -> float:
    """lib.rs:1531-1540."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    na, nb = np.sqrt((a * a).sum()), np.sqrt((b * b).sum())
    return 0.0 if na == 0 or nb == 0 else float((a * b).sum() / (na * nb))


def identify_speaker_from_embedding(emb, speaker_embeddings: dict, threshold: float):
    """lib.rs:1503-1529: best cosine match; threshold relaxed by 0.7 with fewer than 20 speakers; None for usize::MAX."""
    best_sim, best_id = -np.inf, None
    for sid, centroid in speaker_embeddings.items():
        sim = cosine_similarity(emb, centroid)
        if sim > best_sim:
            best_sim, best_id = sim, sid
    dyn = threshold * 0.7 if len(speaker_embeddings) < 20 else threshold
    return best_id if best_sim > dyn else None


def identify_speaker_cosine_emb(emb, speaker_embeds, threshold: float):
    """Decision rule shared by identify_speaker_cosine / _feats (lib.rs:1604-1661); speaker_embeds = [(mean, mean_sim, std_sim)]."""
    best_idx, best_val = None, threshold
    for i, (mean, mean_sim, std_sim) in enumerate(speaker_embeds):
        sim = cosine_similarity(emb, mean)
        if sim < mean_sim - 2.0 * std_sim:
            continue
        factor = 0.3 if len(speaker_embeds) < 200 else 1.0
        dyn = mean_sim + std_sim * factor
        if sim > 0.35 and (sim > dyn or sim > 0.5) and sim > best_val:
            best_val, best_idx = sim, i
    return best_idx


# ----------------------------------------------------------------------------------------------------------------------
# CLI orchestration (main.rs:517-519, 651-668, 750-835), restated on the oracle's primitives
# ----------------------------------------------------------------------------------------------------------------------


def burn_in_limit(dataset_size: int) -> int:
    """main.rs:517-519: ceil(n * DEFAULT_BURN_IN_FRAC) clamped to [10, 50]."""
    return int(min(50, max(10, np.ceil(np.float32(dataset_size) * np.float32(0.2)))))


def average_vectors(vectors) -> np.ndarray:
    """lib.rs:144-159."""
    acc = np.zeros_like(np.asarray(vectors[0]))
    for v in vectors:
        acc = acc + np.asarray(v)
    return normalize(acc / acc.dtype.type(len(vectors)))


def pretrain_from_features(net: Net, windows: np.ndarray, target_class: int, epochs: int, lr: float, dropout: float, batch: int,
                           seed: int) -> float:
    """lib.rs:582-628 with the randomness of the library's Python mirror: shuffles from default_rng(seed), dropout from the
    counter RNG keyed (seed, epoch).  Mean loss over the surviving windows of all epochs."""
    n = len(windows)
    if n == 0 or epochs == 0:
        return 0.0
    rng = np.random.default_rng(seed)
    labels = np.full(n, int(target_class))
    total, count = 0.0, 0
    for e in range(epochs):
        perm = rng.permutation(n)
        keep = dropout_keep_mask(seed, e, np.arange(n), windows.shape[1], dropout)
        l, c = train_epoch(net, windows, labels, perm, batch, lr, keep)
        total += l
        count += c
    return total / count if count else 0.0



# --- raw-audio training loops (lib.rs:348-397, 668-732) with this repo's seeded draws -------------------------------------


def _sm64(x: int) -> int:
    """splitmix64 on a Python int (scalar form of _splitmix64)."""
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def loop_seed(seed: int, file: int, epoch: int) -> int:
    """Seed of (file, epoch) in pretrain_network / train_from_files (include/streamz_b200.h: szb_loop_seed)."""
    return _sm64((_sm64((seed ^ 0x7EA1F11E5) & _M64) + ((file << 32) | epoch)) & _M64)


def shuffle_perm(seed: int, stream: int, n: int) -> np.ndarray:
    """The library's stand-in for `windows.shuffle(&mut thread_rng)` (lib.rs:370, 601): Fisher-Yates from the back with
    j = splitmix64(key ^ i) % (i + 1), key = splitmix64(seed + stream * 0xD1B54A32D192ED03 + 0x5F)."""
    perm = np.arange(n, dtype=np.uint32)
    key = _sm64((seed + stream * 0xD1B54A32D192ED03 + 0x5F) & _M64)
    for i in range(n - 1, 0, -1):
        j = _sm64(key ^ i) % (i + 1)
        perm[i], perm[j] = perm[j], perm[i]
    return perm


def lr_decay(lr: float, step: int) -> np.float32:
    """lr * 0.99f32.powi(step) (lib.rs:709); powi by squaring in float32, as compiler-rt's __powisf2 does."""
    a, r, b = np.float32(0.99), np.float32(1.0), int(step)
    while True:
        if b & 1:
            r = np.float32(r * a)
        b //= 2
        if b == 0:
            break
        a = np.float32(a * a)
    return np.float32(np.float32(lr) * r)


def augment_params(seed: int, n_samples: int) -> Tuple[float, float, int]:
    """Clip-level draws of augment (lib.rs:105-107) from `seed`: noise_level U(0, 0.005), gain U(0.95, 1.05),
    shift in [0, min(len, 800)) -- float32 arithmetic (szb_augment_params)."""
    k = _sm64((seed ^ 0xA06DE27) & _M64)
    f = np.float32
    u0 = f(_sm64(k ^ 1) >> 40) * f(2.0 ** -24)
    u1 = f(_sm64(k ^ 2) >> 40) * f(2.0 ** -24)
    rng_ = min(int(n_samples), WINDOW_SIZE)
    return f(f(0.005) * u0), f(f(0.95) + f(f(0.1) * u1)), (_sm64(k ^ 3) % rng_ if rng_ else 0)


def augment_seeded(samples: np.ndarray, seed: int) -> np.ndarray:
    nl, gain, shift = augment_params(seed, len(samples))
    return augment(samples, shift, gain, nl, _sm64((seed ^ 0x5EED) & _M64))


def pretrain_epoch(net: Net, samples: np.ndarray, target_class: int, lr: float, dropout: float, batch: int, epoch_seed: int,
                   precision: str = "f64") -> Tuple[float, int]:
    """One epoch of lib.rs:367-390: augment -> extract -> shuffle -> dropout / skip-all-zero / loss / train_batch chunks."""
    windows = extract(augment_seeded(samples, epoch_seed), precision).astype(np.float32)
    n = len(windows)
    if n == 0:
        return 0.0, 0
    perm = shuffle_perm(epoch_seed, 0, n)
    keep = dropout_keep_mask(epoch_seed, 0, np.arange(n), windows.shape[1], dropout)
    return train_epoch(net, windows, np.full(n, int(target_class)), perm, batch, lr, keep)


def pretrain_network(net: Net, samples: np.ndarray, target_class: int, epochs: int, lr: float, dropout: float, batch: int,
                     seed: int) -> float:
    """lib.rs:348-397; epoch e draws everything from loop_seed(seed, 0, e).  Mean loss over the windows used (0.0 if none)."""
    total, count = 0.0, 0
    for e in range(int(epochs)):
        l, c = pretrain_epoch(net, samples, target_class, lr, dropout, batch, loop_seed(seed, 0, e))
        total += l
        count += c
    return total / count if count else 0.0


def train_from_files(net: Net, clips: Sequence[np.ndarray], classes: Sequence[int], epochs: int, lr: float, dropout: float, batch: int,
                     seed: int) -> float:
    """lib.rs:668-732 in file-major order (the single-thread serialisation of the rayon loop): per (file, epoch) one
    pretrain_network epoch at lr * 0.99^step, the step counting every (file, epoch) (lib.rs:708-709)."""
    total, count, step = 0.0, 0, 0
    for f, (clip, cls) in enumerate(zip(clips, classes)):
        for e in range(int(epochs)):
            l, c = pretrain_epoch(net, clip, cls, float(lr_decay(lr, step)), dropout, batch, loop_seed(seed, f, e))
            step += 1
            total += l
            count += c
    return total / count if count else 0.0


def add_output_class(net: Net, column: np.ndarray) -> int:
    """lib.rs:797-821 with the new column given (the reference draws it from thread_rng); the new bias is 0."""
    net.w3 = np.concatenate([net.w3, np.asarray(column, net.dtype).reshape(-1, 1)], axis=1)
    net.b3 = np.concatenate([net.b3, np.zeros(1, net.dtype)])
    return net.n_out - 1


def incremental_training(net: Net, train_files: list, feature_map: dict, limit: int, conf_threshold: float, dropout: float,
                         batch: int, epochs: int, seed: int, new_columns) -> dict:
    """main.rs:750-835 in list order: embedding of the file, burn-in / labelled / matched speaker assignment with class
    growth, `epochs` epochs on the file's windows at lr 0.05 (0.01 after 1000 files), running per-speaker mean embeddings."""
    cols = iter(new_columns)
    speaker_features, embeds, log = {}, {}, []
    total_loss, count = 0.0, 0
    for entry in train_files:
        path, cls = entry[0], entry[1]
        windows = feature_map.get(path)
        if windows is None or len(windows) < 5:                  # main.rs:756-760, 829
            continue
        emb = normalize(embedding_mean(net, windows))            # main.rs:763-767
        burn = count < limit
        threshold = 0.5 if burn else conf_threshold
        if burn and cls is None:                                 # main.rs:778-785
            speaker = add_output_class(net, next(cols))
        elif cls is not None:
            speaker = int(cls)
        else:                                                    # main.rs:788-798
            matched = identify_speaker_from_embedding(emb, embeds, threshold)
            speaker = add_output_class(net, next(cols)) if matched is None or matched >= net.n_out else int(matched)
        entry[1] = speaker
        lr = 0.05 if count < 1000 else 0.01
        total_loss += pretrain_from_features(net, windows, speaker, epochs, lr, dropout, batch, seed * 7919 + count)
        speaker_features.setdefault(speaker, []).append(emb)
        embeds[speaker] = average_vectors(speaker_features[speaker])
        count += 1
        log.append((path, speaker))
    return {"total_loss": total_loss, "count": count, "speaker_embeddings": embeds, "log": log}


# ----------------------------------------------------------------------------------------------------------------------
# Synthetic audio (SURVEY.md section 8(d) generator; deterministic)
# ----------------------------------------------------------------------------------------------------------------------


def synth_clip(speaker: int, clip_id: int, seconds: float, rate: int = DEFAULT_SAMPLE_RATE) -> np.ndarray:
    """Harmonic stack with speaker-specific pitch/formants, 4-6 Hz syllable envelope, white noise at -30 dBFS,
    peak -6 dBFS, rounded to i16 mono."""
    n = int(round(seconds * rate))
    rng = np.random.default_rng(0x5A17 ^ int(clip_id))
    t = np.arange(n, dtype=np.float64) / rate
    f0 = 85.0 + (37 * speaker) % 170
    formants = [500.0 + 37.0 * ((speaker * 7) % 11), 1500.0 + 61.0 * ((speaker * 5) % 13), 2500.0 + 83.0 * ((speaker * 3) % 7)]
    x = np.zeros(n)
    for h in range(1, 21):
        f = f0 * h
        if f >= 0.45 * rate:
            break
        gain = sum(1.0 / (1.0 + ((f - fm) / (120.0 + 40.0 * i)) ** 2) for i, fm in enumerate(formants)) + 0.02
        x += gain * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi))
    env = 0.55 + 0.45 * np.sin(2 * np.pi * (4.0 + (clip_id % 5) * 0.5) * t + rng.uniform(0, 2 * np.pi))
    x = x * env
    x = x / max(np.abs(x).max(), 1e-9) * 0.5
    x = x + rng.standard_normal(n) * (10 ** (-30 / 20))
    x = x / max(np.abs(x).max(), 1e-9) * 0.5
    return np.round(x * 32767.0).astype(np.int16)
