"""BASELINE.json's full-size configurations, checked through size-independent properties (the oracle needs seconds per
clip, so at these sizes it only sees a sample):

  configs[1]  10 000 x 10-s clips at 16 kHz -> 11 010 000 windows (the bench workload), device resident:
              identical clips give bit-identical rows wherever they sit in the batch (scheduling / segmentation
              independence), every row is z-scored (lib.rs:328-340), a sample of clips matches the float64 oracle to 1e-4;
  configs[2]  1 000 000 cached windows, 100 speakers, batch 4096: every window survives the input dropout (lib.rs:605-609
              skips only all-zero rows), the epoch loss falls epoch over epoch on separable data, the trained weights stay
              finite and the GPU forward pass with them matches the oracle's (same argmax, probabilities 5e-5)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_configs1_full_size_extraction_properties(sz, ctx, oracle, native):
    torch = pytest.importorskip("torch")
    sys.path.insert(0, ROOT)
    import bench                                               # only its synthetic-clip generator is used
    dev = torch.device("cuda", 0)
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs ~10 GB of device memory")
    n_clips, n_in, rate, distinct = 10000, 160000, 16000, 16
    base = bench.synth_clips_device(torch, dev, distinct, n_in, rate, seed=3)
    pcm = base.repeat(n_clips // distinct, 1).contiguous()     # clip c == base[c % 16]
    off = np.arange(n_clips + 1, dtype=np.uint64) * n_in
    total = int(native.lib.szb_extract_batch_windows(native.ptr(off), n_clips, rate))
    per_clip = oracle.n_windows(oracle.resample_out_len(n_in, rate))
    assert per_clip == 1101 and total == n_clips * per_clip
    feats = torch.empty((total, 60), dtype=torch.float32, device=dev)
    woff = np.zeros(n_clips + 1, np.uint64)
    torch.cuda.synchronize()
    native.check(native.lib.szb_extract_batch_dev(ctx.handle, C.c_void_p(pcm.data_ptr()), native.ptr(off), n_clips, rate,
                                                  C.c_void_p(feats.data_ptr()), total, native.ptr(woff)))
    ctx.sync()
    assert np.array_equal(woff, np.arange(n_clips + 1, dtype=np.uint64) * per_clip)
    f = feats.view(n_clips // distinct, distinct, per_clip, 60)
    assert bool(torch.isfinite(feats).all())
    for g in range(0, n_clips // distinct, 25):                # chunks keep the boolean temporaries small
        assert torch.equal(f[g:g + 25], f[0:1].expand(min(25, n_clips // distinct - g), -1, -1, -1)), g
    mean = feats.mean(dim=1)
    var = feats.var(dim=1, unbiased=False)
    assert float(mean.abs().max()) < 2e-5 and float((var - 1).abs().max()) < 2e-4
    for i in (0, 7, 15):                                       # a sample against the float64 oracle
        want = oracle.extract(oracle.resample_to_44100(base[i].cpu().numpy(), rate))
        got = f[0, i].cpu().numpy()
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-4


def test_configs4_identification_sweep_shards_by_window_range(sz, ctx, oracle):
    """BASELINE configs[4]: a 60-s mixed clip (6 614 windows) cut into window ranges for 1 / 2 / 4 / 8 ranks.  Every range's
    rows are bit-identical to the rows of the whole-clip extraction, so the per-rank histograms add up to the single-GPU one
    and the speaker list is the same (the ranks run one after the other here; each only sees its own range + halo)."""
    from streamz_b200.sharding import shard_windows
    clip = np.concatenate([oracle.synth_clip(s, 60 + s, 15.0) for s in range(4)])          # four speakers x 15 s
    ex = sz.FeatureExtractor(ctx)
    full = ex.extract(clip)
    assert full.shape == (6614, 60)
    onet = oracle.Net.init(60, 512, 256, 6, seed=31)
    net = sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx)
    # a few epochs so that the histogram is not degenerate
    labels = (np.arange(len(full)) * 4 // len(full)).astype(np.uint32)
    data = sz.DeviceFeatures(ctx, full, labels)
    for e in range(2):
        sz.train_epoch(net, data, np.random.default_rng(e).permutation(len(full)).astype(np.uint32), 64, 0.02, seed=1, stream=e)
    data.close()
    thr = 0.5
    want = sz.identify_counts(net, full, thr)
    assert want.sum() > 0
    for world in (1, 2, 4, 8):
        total = np.zeros_like(want)
        for rank, (w0, w1) in enumerate(shard_windows(len(full), world)):
            assert np.array_equal(ex.extract_range(clip, w0, w1), full[w0:w1])
            total += sz.identify_counts_sharded(net, clip, thr, rank, world, ex)
        assert np.array_equal(total, want)
        assert sz.speakers_from_counts(total) == sz.identify_speaker_list(net, clip, thr, ex)
    # edge ranges: empty, single window, the clip's first and last windows
    assert ex.extract_range(clip, 10, 10).shape == (0, 60)
    for w0, w1 in ((0, 1), (0, 3), (6613, 6614), (6610, 6614), (1, 2), (2, 5)):
        assert np.array_equal(ex.extract_range(clip, w0, w1), full[w0:w1])
    short = clip[:800 + 400 * 2]                                                            # 3 windows: every range clamps at both edges
    fs = ex.extract(short)
    for w0 in range(3):
        for w1 in range(w0 + 1, 4):
            assert np.array_equal(ex.extract_range(short, w0, w1), fs[w0:w1])


def test_configs2_full_size_training_properties(sz, ctx, oracle):
    n, n_spk, batch = 1_000_000, 100, 4096
    r = np.random.default_rng(11)
    labels = r.integers(0, n_spk, n).astype(np.uint32)
    centres = r.standard_normal((n_spk, 60)).astype(np.float32)
    x = centres[labels] + 0.7 * r.standard_normal((n, 60)).astype(np.float32)
    x = ((x - x.mean(axis=1, keepdims=True)) / x.std(axis=1, keepdims=True)).astype(np.float32)   # rows as the front end emits them
    onet = oracle.Net.init(60, 512, 256, n_spk, seed=5)
    net = sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx)
    data = sz.DeviceFeatures(ctx, x, labels)
    losses = []
    for epoch in range(3):
        perm = np.random.default_rng(100 + epoch).permutation(n).astype(np.uint32)
        loss, used = sz.train_epoch(net, data, perm, batch, 0.01, dropout=0.2, seed=9, stream=epoch)
        assert used == n                                       # P(all 60 inputs dropped) = 0.2^60
        losses.append(loss / used)
    data.close()
    assert all(np.isfinite(losses)) and losses[0] > losses[1] > losses[2], losses
    w = net.weights()
    assert all(np.isfinite(a).all() for a in w)
    trained = oracle.Net(*w, dtype=np.float64)
    sample = x[:: n // 2000]
    got, want = net.forward(sample), oracle.forward(trained, sample)
    assert np.abs(got - want).max() <= 5e-5
    srt = np.sort(want, axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > 1e-4                   # identical labels outside the near-tie margin
    assert np.array_equal(got.argmax(axis=1)[clear], want.argmax(axis=1)[clear])
    acc = float((got.argmax(axis=1) == labels[:: n // 2000]).mean())
    assert acc > 0.5, acc                                      # the separable classes are being learnt
