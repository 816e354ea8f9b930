"""The oracle against known answers, the independent golden pipeline, and itself (numpy f64 vs the C float32 port).
The reference's own tests pin no numbers for this path (lib.rs:1827-1865), so the pins are: mathematical known-answer
vectors, third-party library pieces (torchaudio mel, scipy rfft/dct; tests/golden, tools/make_golden.py), and a
finite-difference check of the backward pass."""
import ctypes

import numpy as np
import pytest

from conftest import CNet, P


def test_mel_bank_matches_torchaudio_fixture(oracle, golden):
    mel = oracle.mel_filterbank()
    ref = golden["mel_torchaudio_26x401"]
    assert mel.shape == (26, 401)
    assert np.abs(mel - ref).max() < 1e-8          # torchaudio evaluates the same triangles in float32
    assert 735 <= int((mel > 0).sum()) <= 745      # ~740 non-zeros


def test_mel_bank_matches_a_second_independent_implementation(oracle):
    # transformers.audio_utils.mel_filter_bank is Hugging Face's own restatement of librosa's filter bank (Slaney scale, Slaney
    # area normalisation): a second witness, beside torchaudio's, of the semantics the oracle assumes for the mel_filter crate
    # (lib.rs:240-248).  Evaluated live (no fixture); skipped when the package cannot be imported.
    audio_utils = pytest.importorskip("transformers.audio_utils")
    hf = audio_utils.mel_filter_bank(num_frequency_bins=401, num_mel_filters=26, min_frequency=0.0, max_frequency=22050.0,
                                     sampling_rate=44100, norm="slaney", mel_scale="slaney")
    mine = oracle.mel_filterbank(dtype=np.float64)
    assert hf.shape == (401, 26) and np.abs(hf.T - mine).max() <= 1e-12


def test_oracle_matches_independent_golden(oracle, golden):
    for name in ("a", "b"):
        got = oracle.extract(golden[f"clip_{name}_i16"])
        want = golden[f"clip_{name}_features_f64"]
        assert got.shape == want.shape
        # torchaudio builds its bank in float32 (the oracle in float64, then rounds): weights agree to 6e-9 absolute,
        # which moves normalised features by ~1e-6 -- two orders below the 1e-4 parity tolerance
        assert np.abs(got - want).max() < 5e-6


def test_window_count_and_short_input(oracle):
    # lib.rs:289: fewer than 800 samples -> empty, no error
    for n, w in ((0, 0), (799, 0), (800, 1), (1199, 1), (1200, 2), (220500, 550), (441000, 1101), (2646000, 6614)):
        assert oracle.n_windows(n) == w
    assert oracle.extract(np.zeros(799, np.int16)).shape == (0, 60)


def test_silent_frame_known_answer(oracle):
    # all-zero frame: every mel energy is ln(1e-12); c0 = 26 ln(1e-12), other c_j = 0; deltas 0
    f = oracle.extract(np.zeros(800 + 400 * 3, np.int16))
    c0 = 26 * np.log(1e-12)
    v = np.zeros(60); v[0] = c0
    want = (v - v.mean()) / v.std()
    assert np.allclose(f, want[None, :], atol=1e-9)
    assert abs(f[0, 0] - (-7.681146)) < 1e-5 and abs(f[0, 1] - 0.1301889) < 1e-6


def test_single_window_has_zero_deltas(oracle):
    c = oracle.synth_clip(1, 1, 800 / 44100.0)[:800]
    base = oracle.mfcc_frames(c)
    assert base.shape == (1, 20)
    assert np.all(oracle.add_deltas(base) == 0)


def test_pure_tone_bin_power(oracle):
    # bin-centred cosine, rectangular window: |X_k|^2 = (A * 400)^2
    k, amp = 37, 0.25
    x = amp * np.cos(2 * np.pi * k * np.arange(800) / 800)
    spec = np.fft.fft(x)
    assert abs(abs(spec[k]) ** 2 - (amp * 400) ** 2) < 1e-6


def test_fft_definition_against_direct_dft(oracle):
    clip = oracle.synth_clip(2, 9, 0.2)
    assert np.abs(oracle.extract(clip) - oracle.extract_direct_dft(clip)).max() < 1e-9


def test_deltas_edge_replication(oracle):
    x = np.arange(5 * 20, dtype=np.float64).reshape(5, 20) ** 1.5
    d = oracle.add_deltas(x)
    assert np.allclose(d[0], (x[1] - x[0]) / 2) and np.allclose(d[4], (x[4] - x[3]) / 2) and np.allclose(d[2], (x[3] - x[1]) / 2)


def test_downmix_truncates_toward_zero(oracle):
    s = np.array([-3, -4, 5, 2, -32768, -32768, 7], np.int16)   # trailing partial chunk divided by channels too
    assert oracle.downmix_to_mono(s, 2).tolist() == [-3, 3, -32768, 3]
    assert oracle.downmix_to_mono(s, 1).tolist() == s.tolist()


def test_c_port_matches_numpy_oracle(oracle, oracle_c):
    clip = oracle.synth_clip(4, 21, 1.5)
    mel, dct = oracle.mel_filterbank(), oracle.dct2_matrix(dtype=np.float32)
    n = oracle.n_windows(len(clip))
    out = np.zeros((n, 60), np.float32)
    assert oracle_c.so_extract(P(clip), ctypes.c_size_t(len(clip)), P(mel), P(dct), P(out)) == n
    assert np.abs(out - oracle.extract(clip)).max() < 2e-5      # float32 noise floor of the reference's arithmetic
    assert np.abs(oracle.extract(clip, "f32") - oracle.extract(clip)).max() < 2e-5


def test_resampler_contract(oracle, oracle_c):
    # identity at 44.1 kHz (lib.rs:187-189), output length floor(n*44100/rate) (lib.rs:196), DC gain 1, C == numpy
    s = oracle.synth_clip(1, 2, 0.25, rate=16000)
    assert np.array_equal(oracle.resample_to_44100(s, 44100), s)
    for rate in (8000, 16000, 22050, 32000, 48000):
        x = oracle.synth_clip(1, 2, 0.05, rate=rate)
        y = oracle.resample_to_44100(x, rate)
        assert len(y) == len(x) * 44100 // rate
        taps = oracle.resample_taps(rate)
        L, M = oracle.resample_ratio(rate)
        yc = np.zeros(len(y), np.int16)
        oracle_c.so_resample(P(x), ctypes.c_size_t(len(x)), ctypes.c_uint32(rate), P(taps), ctypes.c_uint32(L), ctypes.c_uint32(M),
                             ctypes.c_uint32(16), P(yc))
        assert np.array_equal(y, yc)
    dc = np.full(4000, 12345, np.int16)
    y = oracle.resample_to_44100(dc, 16000)
    assert np.all(np.abs(y[100:-100].astype(int) - 12345) <= 1)
    # a 1 kHz tone stays a 1 kHz tone
    t = np.arange(16000) / 16000.0
    y = oracle.resample_to_44100(np.round(8000 * np.sin(2 * np.pi * 1000 * t)).astype(np.int16), 16000)
    spec = np.abs(np.fft.rfft(y.astype(np.float64)))
    assert abs(np.argmax(spec) * 44100.0 / len(y) - 1000.0) < 2.0


def _tiny_net(oracle, dtype=np.float64):
    # the shape of the reference's own test net (lib.rs:1834): 4 -> 3 -> 2 -> 2
    r = np.random.default_rng(7)
    return oracle.Net(r.uniform(-.5, .5, (4, 3)), r.uniform(-.1, .1, 3), r.uniform(-.5, .5, (3, 2)), r.uniform(-.1, .1, 2),
                      r.uniform(-.5, .5, (2, 2)), r.uniform(-.1, .1, 2), dtype=dtype)


def test_softmax_and_forward_known_answers(oracle):
    net = _tiny_net(oracle)
    net.w3[:] = 0; net.b3[:] = 0                       # equal logits -> uniform probabilities
    assert np.allclose(oracle.forward(net, np.array([0.1, -0.2, 0.3, 0.4])), 0.5)
    net = _tiny_net(oracle)
    x = np.array([[0.1, -0.2, 0.3, 0.4]])
    h1 = np.maximum(x @ net.w1 + net.b1, 0); h2 = np.tanh(h1 @ net.w2 + net.b2); z = h2 @ net.w3 + net.b3
    assert np.allclose(oracle.forward(net, x), np.exp(z) / np.exp(z).sum())


def test_backward_matches_finite_differences(oracle):
    net = _tiny_net(oracle)
    x = np.array([[0.1, -0.2, 0.3, 0.4], [0.7, 0.1, -0.5, 0.2]])
    t = np.array([[1.0, 0.0], [0.0, 1.0]])
    grads, _ = oracle.gradients(net, x, t)
    def loss():
        return float(-(t * np.log(oracle.forward(net, x))).sum())
    for w, g in zip(net.params(), grads):
        it = np.nditer(w, flags=["multi_index"])
        for _ in it:
            i = it.multi_index
            old = w[i]
            w[i] = old + 1e-6; lp = loss()
            w[i] = old - 1e-6; lm = loss()
            w[i] = old
            assert abs((lp - lm) / 2e-6 - g[i]) < 1e-6


def test_train_batch_one_step_weights_change(oracle):
    # the reference's `weights_change_after_training` (lib.rs:1832-1851), with the update also pinned numerically
    net = _tiny_net(oracle, np.float32)
    before = [p.copy() for p in net.params()]
    x = np.array([[0.1, -0.2, 0.3, 0.4]], np.float32)
    grads, _ = oracle.gradients(net, x, np.array([[1.0, 0.0]], np.float32))
    oracle.train_batch(net, x, np.array([1.0, 0.0], np.float32), 0.1)
    assert any(np.any(a != b) for a, b in zip(before, net.params()))
    for b, g, a in zip(before, grads, net.params()):
        assert np.allclose(a, b - 0.1 * g, atol=1e-7)
    oracle.train_batch(net, np.zeros((0, 4), np.float32), np.array([1.0, 0.0], np.float32), 0.1)   # empty batch: no-op


def test_c_port_training_matches_numpy(oracle, oracle_c):
    r = np.random.default_rng(3)
    net = oracle.Net.init(60, 32, 16, 5, seed=3)
    arrs = [np.ascontiguousarray(p.copy()) for p in net.params()]
    cnet = CNet.from_arrays(arrs)
    feats = r.standard_normal((50, 60)).astype(np.float32)
    feats[7] = 0                                         # an all-zero window is skipped (lib.rs:607-609)
    labels = r.integers(0, 6, 50).astype(np.uint32)      # label 5 >= n_out -> all-zero target (lib.rs:592-595)
    perm = r.permutation(50).astype(np.uint32)
    keep = oracle.dropout_keep_mask(11, 0, np.arange(50), 60, 0.2)
    loss_np, cnt_np = oracle.train_epoch(net, feats, labels, perm, 8, 0.05, keep)
    loss_c = ctypes.c_double()
    k8 = np.ascontiguousarray(keep.astype(np.uint8))
    cnt_c = oracle_c.so_train_epoch(ctypes.byref(cnet), P(feats), P(labels), P(perm), ctypes.c_size_t(50), ctypes.c_size_t(8),
                                    ctypes.c_float(0.05), P(k8), ctypes.byref(loss_c))
    assert cnt_c == cnt_np == 49
    assert abs(loss_c.value - loss_np) < 1e-3
    for a, b in zip(arrs, net.params()):
        assert np.abs(a - b).max() < 1e-5


def test_aggregation_rules(oracle):
    # last maximal index wins (lib.rs:1393-1396); `>=` threshold (lib.rs:1398); stable sort by count desc (lib.rs:1409)
    p = np.array([[0.4, 0.4, 0.2], [0.1, 0.45, 0.45], [0.8, 0.1, 0.1]])
    assert oracle.argmax_last(p).tolist() == [1, 2, 0]
    assert oracle.speakers_from_counts([3, 0, 5, 3, 5]) == [2, 4, 0, 3]
    net = oracle.Net.init(60, 16, 8, 4, seed=1)
    feats = np.random.default_rng(0).standard_normal((40, 60)).astype(np.float32)
    pr = oracle.forward(net, feats)
    thr = float(np.sort(pr.max(axis=1))[20])
    counts = oracle.identify_counts(net, feats, thr)
    assert counts.sum() == 20                           # exactly the windows with p_max >= thr
    assert oracle.identify_speaker_with_threshold_feats(oracle.Net.init(60, 8, 4, 1), feats, 0.0) is None   # C <= 1
    assert oracle.identify_speaker_with_threshold_feats(net, feats[:0], 0.0) is None
    assert oracle.identify_speaker(net, feats[:0]) == 0
