"""Parity of the CUDA front end against the oracle and the golden fixtures, through the C ABI (ctypes).
Tolerance: max-abs 1e-4 on normalised MFCC+delta features against the float64 oracle (BASELINE.json north_star);
the resampler + quantisation is integer-exact."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _check(sz_ex, oracle, clip):
    got = sz_ex.extract(clip)
    want = oracle.extract(clip)
    assert got.shape == want.shape and got.dtype == np.float32
    if len(want):
        assert np.abs(got - want).max() <= TOL
    return got


@pytest.fixture(scope="module")
def ex(sz, ctx):
    return sz.FeatureExtractor(ctx)


def test_golden_fixtures(ex, golden):
    for name in ("a", "b"):
        got = ex.extract(golden[f"clip_{name}_i16"])
        assert got.shape == golden[f"clip_{name}_features_f64"].shape
        assert np.abs(got - golden[f"clip_{name}_features_f64"]).max() <= TOL


@pytest.mark.parametrize("n", [0, 1, 799, 800, 801, 1199, 1200, 1201, 1999, 2000, 800 + 400 * 31, 800 + 400 * 32, 800 + 400 * 33,
                               800 + 400 * 63 + 399, 800 + 400 * 64, 44100])
def test_edge_lengths(ex, oracle, n):
    clip = oracle.synth_clip(n % 5, n, max(n, 1) / 44100.0)[:n]
    got = _check(ex, oracle, clip)
    assert got.shape[0] == oracle.n_windows(n)          # short input -> empty, no error (lib.rs:289)


def test_silence_and_extremes(ex, oracle):
    f = ex.extract(np.zeros(800 + 400 * 5, np.int16))
    assert np.allclose(f[:, 0], -7.681146, atol=1e-5) and np.allclose(f[:, 1:], 0.1301889, atol=1e-6)
    # full-scale inputs (|x| up to 32768, frame sums up to 2.6e7) with energy in every mel band.  A signal whose
    # spectrum is exactly zero in some band (e.g. a bin-periodic square wave) is NOT a parity case: ln() of pure
    # float32 rounding noise is arbitrary in the reference itself (DESIGN.md "Conditioning").
    r = np.random.default_rng(0)
    _check(ex, oracle, r.integers(-32768, 32768, 44100).astype(np.int16))
    loud = np.where(r.random(20000) < 0.5, 32767, -32768).astype(np.int16)
    _check(ex, oracle, loud)
    quiet = (oracle.synth_clip(1, 3, 0.5) // 2000).astype(np.int16)               # a few LSBs of signal
    _check(ex, oracle, quiet)


def test_five_second_and_sixty_second_clips(ex, oracle):
    c5 = oracle.synth_clip(2, 5, 5.0)
    assert _check(ex, oracle, c5).shape == (550, 60)
    c60 = np.concatenate([oracle.synth_clip(s, 60 + s, 15.0) for s in range(4)])   # mixed-speaker 60 s clip (C5)
    assert _check(ex, oracle, c60).shape == (6614, 60)


def test_ragged_batch_with_empty_and_misaligned_clips(ex, oracle):
    lens = [0, 13001, 799, 800, 44100, 2403, 1, 30000, 17777]
    clips = [oracle.synth_clip(i % 4, 100 + i, max(l, 1) / 44100.0)[:l] for i, l in enumerate(lens)]
    outs = ex.extract_batch(clips)
    for o, c in zip(outs, clips):
        w = oracle.extract(c)
        assert o.shape == w.shape
        if len(w):
            assert np.abs(o - w).max() <= TOL


def test_segmented_long_clip_is_bit_identical_to_single_clip_path(ex, oracle):
    # a clip alone in a batch is split into many window-range segments (halo recompute at the cuts); inside a
    # large batch it is one segment.  Both must produce the same bits.
    clip = oracle.synth_clip(3, 77, 8.0)
    alone = ex.extract(clip)
    filler = [oracle.synth_clip(i % 3, i, 8.0) for i in range(700)]
    in_batch = ex.extract_batch(filler[:350] + [clip] + filler[350:])[350]
    assert np.array_equal(alone, in_batch)


def test_linearity_in_gain_shifts_only_c0(ex, oracle):
    # size-independent property: doubling the signal adds ln 4 to every mel energy, i.e. 26 ln 4 to c0 only
    clip = (oracle.synth_clip(1, 9, 2.0) // 4).astype(np.int16)
    a, b = oracle.mfcc_frames(clip), oracle.mfcc_frames((clip * 2).astype(np.int16))
    assert np.allclose(b[:, 0] - a[:, 0], 26 * np.log(4.0), atol=1e-6) and np.allclose(b[:, 1:], a[:, 1:], atol=1e-6)
    fa, fb = ex.extract(clip), ex.extract((clip * 2).astype(np.int16))
    wa, wb = oracle.extract(clip), oracle.extract((clip * 2).astype(np.int16))
    assert np.abs((fb - fa) - (wb - wa)).max() <= 2 * TOL


def test_rows_are_z_scored(ex, oracle):
    f = ex.extract(oracle.synth_clip(0, 1, 10.0)).astype(np.float64)
    assert f.shape == (1101, 60)
    assert np.abs(f.mean(axis=1)).max() < 1e-5 and np.abs(f.std(axis=1) - 1).max() < 1e-4


def test_downmix(sz, ctx, oracle):
    r = np.random.default_rng(0)
    for ch, n in ((2, 20001), (3, 999), (6, 12000), (1, 100)):
        s = r.integers(-32768, 32768, n).astype(np.int16)
        assert np.array_equal(sz.downmix_to_mono(s, ch, ctx), oracle.downmix_to_mono(s, ch))


@pytest.mark.parametrize("rate", [8000, 11025, 12000, 16000, 22050, 24000, 32000, 37800, 44000, 48000])
def test_resampler_is_bit_exact(sz, ctx, oracle, native, rate):
    x = oracle.synth_clip(2, rate, 0.7, rate=rate)
    L, M = oracle.resample_ratio(rate)
    taps = np.zeros((L, 16), np.float32)
    native.check(native.lib.szb_table_resample_taps(rate, taps.ctypes.data_as(C.c_void_p), None, None))
    got = sz.resample_to_44100(x, rate, ctx)
    assert got.dtype == np.int16 and len(got) == len(x) * 44100 // rate            # lib.rs:196
    assert np.array_equal(got, oracle.resample_to_44100(x, rate, taps))            # same taps -> same bits
    indep = oracle.resample_to_44100(x, rate)                                      # oracle's own taps
    assert np.abs(got.astype(int) - indep.astype(int)).max() <= 1
    assert (got != indep).mean() < 1e-3
    loud = np.where(np.arange(3000) % 40 < 20, 32767, -32768).astype(np.int16)     # overshoot must clamp (lib.rs:207)
    assert np.array_equal(sz.resample_to_44100(loud, rate, ctx), oracle.resample_to_44100(loud, rate, taps))


def test_resample_identity_and_empty(sz, ctx):
    s = np.arange(-50, 50, dtype=np.int16)
    assert np.array_equal(sz.resample_to_44100(s, 44100, ctx), s)                  # lib.rs:187-189
    assert len(sz.resample_to_44100(np.zeros(0, np.int16), 16000, ctx)) == 0


@pytest.mark.parametrize("rate", [16000, 32000, 48000])
def test_batch_from_other_rates_equals_resample_then_extract(sz, ctx, ex, oracle, rate):
    clips = [oracle.synth_clip(i % 3, 200 + i, 0.4 + 0.37 * i, rate=rate) for i in range(5)] + [np.zeros(10, np.int16)]
    outs = ex.extract_batch(clips, rate=rate)
    for o, c in zip(outs, clips):
        r = sz.resample_to_44100(c, rate, ctx)
        assert np.array_equal(o, ex.extract(r))                                    # fused path == API composition, bit for bit
        w = oracle.extract(r)
        assert o.shape == w.shape
        if len(w):
            assert np.abs(o - w).max() <= TOL


def test_capacity_and_argument_errors(sz, ctx, native, oracle):
    clip = oracle.synth_clip(0, 0, 0.5)
    out = np.zeros((3, 60), np.float32)
    n = C.c_uint64()
    st = native.lib.szb_extract(ctx.handle, clip.ctypes.data_as(C.c_void_p), len(clip), out.ctypes.data_as(C.c_void_p), 3, C.byref(n))
    assert st == native.ERR_INVALID and b"capacity" in native.lib.szb_last_error()
    assert n.value == oracle.n_windows(len(clip))                                   # size is still reported
    assert native.lib.szb_resample_to_44100(ctx.handle, None, 0, 0, None, 0, C.byref(n)) == native.ERR_INVALID


def test_augment_is_integer_exact(sz, ctx, oracle, native):
    # lib.rs:103-116 with the random draws derived from the seed
    clip = oracle.synth_clip(1, 5, 0.6)
    for seed in (0, 1, 12345):
        nl, gain, shift = sz.augment_params(seed, len(clip))
        assert 0 <= nl < 0.005 and 0.95 <= gain < 1.05 and 0 <= shift < 800        # lib.rs:105-107
        key = int(oracle._splitmix64(np.array([seed ^ 0x5EED], dtype=np.uint64))[0])
        got = sz.augment(clip, seed, ctx)
        assert got.dtype == np.int16 and np.array_equal(got, oracle.augment(clip, shift, gain, nl, key))
    short = clip[:100]
    assert sz.augment_params(7, len(short))[2] < 100                                # shift < len for short clips
    assert len(sz.augment(np.zeros(0, np.int16), 3, ctx)) == 0
    loud = np.full(5000, 32767, np.int16)                                           # gain > 1 must clamp, not wrap
    assert sz.augment(loud, 4, ctx).max() <= 32767 and sz.augment(-loud - 1, 4, ctx).min() >= -32768


def test_pretrain_network_equals_the_oracle_loop(sz, ctx, oracle):
    # lib.rs:348-397 with every draw seeded (szb_loop_seed): augment -> extract -> shuffle -> dropout -> train_batch chunks.
    # Same augmented samples (integer-exact), same order, same dropout decisions => weights within 1e-4 of the float32 oracle.
    clip = oracle.synth_clip(1, 21, 0.35)                         # 37 windows: 5 steps of batch 8 per epoch
    onet = oracle.Net.init(60, 512, 256, 3, seed=11)
    net = sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx)
    got = sz.pretrain_network(net, clip, 2, 3, 2, 0.01, 0.2, 8, seed=17)
    want = oracle.pretrain_network(onet, clip, 2, 2, 0.01, 0.2, 8, seed=17)
    err = max(float(np.abs(a - b).max()) for a, b in zip(net.weights(), onet.params()))
    assert err <= 1e-4, err
    assert abs(got - want) <= 1e-4 * max(1.0, abs(want)), (got, want)
    assert sz.pretrain_network(net, clip[:500], 0, 3, 2, 0.01, 0.2, 8) == 0.0        # no windows -> 0.0 (lib.rs:392-396)


def test_train_from_files_equals_the_oracle_loop(sz, ctx, oracle):
    # lib.rs:668-732: per (file, epoch) one pretrain_network epoch at lr * 0.99^step; file-major order
    clips = [oracle.synth_clip(0, 31, 0.30), oracle.synth_clip(1, 32, 0.26), np.zeros(300, np.int16), oracle.synth_clip(0, 33, 0.22)]
    classes = [0, 1, 1, 0]
    onet = oracle.Net.init(60, 512, 256, 2, seed=12)
    net = sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx)
    files = [(f"spk/{i}.wav", c, k) for i, (c, k) in enumerate(zip(clips, classes))]
    got = sz.train_from_files(net, files, 2, 2, 0.05, 0.2, 8, seed=5)
    want = oracle.train_from_files(onet, clips, classes, 2, 0.05, 0.2, 8, seed=5)
    err = max(float(np.abs(a - b).max()) for a, b in zip(net.weights(), onet.params()))
    assert err <= 1e-4, err
    assert abs(got - want) <= 1e-4 * max(1.0, abs(want)), (got, want)
    assert net.file_lists() == [["spk/0.wav", "spk/3.wav"], ["spk/1.wav", "spk/2.wav"]]      # lib.rs:723
    assert float(oracle.lr_decay(0.05, 7)) < 0.05 * 0.99 ** 6                                # the step counter really decays


def test_back_to_back_dev_batches_keep_their_own_segment_tables(sz, ctx, oracle, native):
    # The *_dev entry points return without synchronising; each call's segment table / offset table must survive until its
    # own H2D copy has run (a single pinned staging buffer used to be overwritten by the next call).
    import ctypes as C
    clips_a = [oracle.synth_clip(0, 41, 0.3), oracle.synth_clip(1, 42, 0.5)]
    clips_b = [oracle.synth_clip(2, 43, 0.45), oracle.synth_clip(3, 44, 0.2), oracle.synth_clip(4, 45, 0.33)]
    ex = sz.FeatureExtractor(ctx)
    want_a, want_b = ex.extract_batch(clips_a), ex.extract_batch(clips_b)
    outs = []
    bufs = []
    for rep in range(6):                                           # A, B, A, B ... queued with no sync in between
        clips = clips_a if rep % 2 == 0 else clips_b
        pcm, off = sz.pack_clips(clips)
        d_pcm = ctx.dev_alloc(pcm.nbytes + 64)
        ctx.h2d(d_pcm, pcm)
        total = int(native.lib.szb_extract_batch_windows(native.ptr(off), len(clips), 44100))
        d_out = ctx.dev_alloc(total * 240)
        bufs.append((d_pcm, d_out, total, off))
    for d_pcm, d_out, total, off in bufs:
        woff = np.zeros(len(off), np.uint64)
        native.check(native.lib.szb_extract_batch_dev(ctx.handle, C.c_void_p(d_pcm), native.ptr(off), len(off) - 1, 44100,
                                                      C.c_void_p(d_out), total, native.ptr(woff)))
    ctx.sync()
    for rep, (d_pcm, d_out, total, off) in enumerate(bufs):
        got = np.empty((total, 60), np.float32)
        ctx.d2h(got, d_out)
        want = np.concatenate(want_a if rep % 2 == 0 else want_b)
        assert np.array_equal(got, want), rep
        ctx.dev_free(d_pcm); ctx.dev_free(d_out)


def test_fused_resampling_in_the_extraction_kernel_is_bit_identical(sz, ctx, oracle, native):
    # szb_ctx_set_fused_resample(1): the polyphase FIR runs inside the extraction kernel's staging (no 44.1 kHz intermediate in
    # HBM).  Rows must equal resample_kernel -> extract_kernel bit for bit, for ragged lengths, clips shorter than a frame after
    # resampling, several rates, and the unsupported cases must fall back silently (rates above 44.1 kHz; unaligned clip starts).
    ex = sz.FeatureExtractor(ctx)
    try:
        for rate in (16000, 8000, 22050, 32000, 11025, 48000):
            lens = [int(rate * s) & ~7 for s in (0.31, 1.0, 0.05, 2.37, 0.02)] + [8 * 123]
            clips = [oracle.synth_clip(i % 3, 70 + i, n / rate + 0.01, rate=rate)[:n] for i, n in enumerate(lens)]
            native.check(native.lib.szb_ctx_set_fused_resample(ctx.handle, 0))
            want = ex.extract_batch(clips, rate=rate)
            native.check(native.lib.szb_ctx_set_fused_resample(ctx.handle, 1))
            got = ex.extract_batch(clips, rate=rate)
            assert len(got) == len(want)
            for a, b in zip(got, want):
                assert a.shape == b.shape and np.array_equal(a, b), rate
        odd = [oracle.synth_clip(1, 90, 0.5, rate=16000)[:8003], oracle.synth_clip(2, 91, 0.5, rate=16000)[:7999]]   # 2nd clip starts unaligned
        native.check(native.lib.szb_ctx_set_fused_resample(ctx.handle, 0))
        want = ex.extract_batch(odd, rate=16000)
        native.check(native.lib.szb_ctx_set_fused_resample(ctx.handle, 1))
        got = ex.extract_batch(odd, rate=16000)
        assert all(np.array_equal(a, b) for a, b in zip(got, want))
        long_clip = [oracle.synth_clip(0, 92, 9.0, rate=16000)]                       # many tiles, several segments
        assert np.array_equal(ex.extract_batch(long_clip, rate=16000)[0], oracle_free_two_step(sz, ex, native, ctx, long_clip[0]))
    finally:
        native.check(native.lib.szb_ctx_set_fused_resample(ctx.handle, 0))


def oracle_free_two_step(sz, ex, native, ctx, clip16):
    """resample_to_44100 followed by extract through the public API (both on the GPU): the reference of the fused path."""
    native.check(native.lib.szb_ctx_set_fused_resample(ctx.handle, 0))
    out = ex.extract(sz.resample_to_44100(clip16, 16000, ctx))
    native.check(native.lib.szb_ctx_set_fused_resample(ctx.handle, 1))
    return out
