"""Parity of the SimpleNeuralNet kernels and the aggregation against the oracle, through the C ABI.

Tolerances, per arithmetic mode of the dense layers (szb_net_set_precision):
  '3xtf32' (default; tcgen05 tensor cores, split TF32): probabilities max-abs 5e-5, weights after one step 1e-5,
           after an epoch 1e-4, identical argmax labels (outside a 1e-4 near-tie margin) and per-class counts;
  'fp32'   (CUDA cores): probabilities 1e-5, same weight tolerances;
  'tf32'   (tensor cores, single pass): probabilities 2e-2, weights after one step 1e-3, identical argmax outside a
           2e-2 margin."""
PROB_TOL = {"fp32": 1e-5, "3xtf32": 5e-5, "tf32": 2e-2}
STEP_TOL = {"fp32": 1e-5, "3xtf32": 1e-5, "tf32": 1e-3}
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair(sz, ctx, oracle, dims, seed=0):
    onet = oracle.Net.init(*dims, seed=seed)
    onet.b1[:] = np.random.default_rng(seed).uniform(-.1, .1, dims[1])
    onet.b2[:] = np.random.default_rng(seed + 1).uniform(-.1, .1, dims[2])
    onet.b3[:] = np.random.default_rng(seed + 2).uniform(-.1, .1, dims[3])
    return onet, sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx)


def _werr(net, onet):
    return max(float(np.abs(a - b).max()) for a, b in zip(net.weights(), onet.params()))


@pytest.mark.parametrize("mode", ["3xtf32", "fp32", "tf32"])
@pytest.mark.parametrize("dims,B", [((60, 512, 256, 100), 300), ((60, 512, 256, 2), 64), ((60, 512, 256, 1000), 33),
                                    ((4, 3, 2, 2), 5), ((60, 512, 256, 1), 7), ((17, 33, 65, 9), 130), ((60, 512, 256, 100), 5000),
                                    ((60, 200, 72, 10), 300)])     # 16-byte-aligned rows, ragged 32-column epilogue chunks
def test_forward(sz, ctx, oracle, dims, B, mode):
    onet, net = _pair(sz, ctx, oracle, dims, seed=B)
    net.set_precision(mode)
    x = np.random.default_rng(B).standard_normal((B, dims[0])).astype(np.float32)
    q64 = oracle.forward(onet.copy(np.float64), x)
    p = net.forward(x)
    assert p.shape == q64.shape and np.abs(p - q64).max() <= PROB_TOL[mode]
    assert np.abs(p.sum(axis=1) - 1).max() < 1e-5
    margin = np.sort(q64, axis=1)[:, -1] - (np.sort(q64, axis=1)[:, -2] if dims[3] > 1 else 0)
    clear = margin > 2 * PROB_TOL[mode]
    assert np.array_equal(p.argmax(axis=1)[clear], q64.argmax(axis=1)[clear])       # identical labels off near-ties
    assert np.abs(net.forward(x[0]) - p[0]).max() <= (1e-6 if mode != "tf32" else 1e-2)   # single-window call


def test_reference_unit_test_weights_change(sz, ctx):
    # lib.rs:1832-1851
    net = sz.SimpleNeuralNet(4, 3, 2, 2, seed=3, ctx=ctx)
    before = net.weights()
    assert all(np.all(b == 0) for b in (before[1], before[3], before[5]))           # zero biases (lib.rs:772-777)
    assert all(np.abs(w).max() <= 0.5 and np.abs(w).max() > 0 for w in (before[0], before[2], before[4]))
    net.train_batch([[0.1, -0.2, 0.3, 0.4]], [1.0, 0.0], 0.1)
    after = net.weights()
    assert any(np.any(a != b) for a, b in zip(before[:4], after[:4])), "weights did not change after training step"


@pytest.mark.parametrize("mode", ["3xtf32", "fp32", "tf32"])
@pytest.mark.parametrize("dims,B", [((60, 512, 256, 100), 8), ((60, 512, 256, 100), 4096), ((4, 3, 2, 2), 1), ((60, 512, 256, 3), 577),
                                    ((60, 200, 72, 10), 300), ((60, 512, 256, 101), 130)])
def test_train_batch_shared_target(sz, ctx, oracle, dims, B, mode):
    onet, net = _pair(sz, ctx, oracle, dims, seed=B)
    net.set_precision(mode)
    x = np.random.default_rng(B).standard_normal((B, dims[0])).astype(np.float32)
    t = np.zeros(dims[3], np.float32); t[dims[3] // 2] = 1
    net.train_batch(x, t, 0.01); oracle.train_batch(onet, x, t, 0.01)
    assert _werr(net, onet) <= STEP_TOL[mode]
    if mode == "tf32":
        return
    soft = np.random.default_rng(1).dirichlet(np.ones(dims[3])).astype(np.float32)   # any target vector, not only one-hot
    net.train_batch(x, soft, 0.05); oracle.train_batch(onet, x, soft, 0.05)
    assert _werr(net, onet) <= 2e-5
    net.train_batch(np.zeros((0, dims[0]), np.float32), t, 0.01)                    # empty batch: no-op (lib.rs:1003)
    net.train(x[0], t, 0.01); oracle.train_batch(onet, x[:1], t, 0.01)              # train == batch of one (lib.rs:954)
    assert _werr(net, onet) <= 3e-5


def test_train_batch_labels_dropout_and_skip(sz, ctx, oracle):
    onet, net = _pair(sz, ctx, oracle, (60, 512, 256, 10), seed=4)
    r = np.random.default_rng(4)
    x = r.standard_normal((64, 60)).astype(np.float32)
    labels = r.integers(0, 12, 64).astype(np.uint32)              # 10, 11 >= n_out -> all-zero target (lib.rs:592-595)
    keep = r.random((64, 60)) >= 0.2
    keep[5] = False                                               # window left all-zero -> skipped (lib.rs:607-609)
    x[9] = 0
    loss, used = net.train_batch_labels(x, labels, 0.01, keep)
    oloss, oused = oracle.train_epoch(onet, x, labels, np.arange(64), 64, 0.01, keep)
    assert used == oused == 62
    assert abs(loss - oloss) <= 1e-3 * max(1.0, abs(oloss))
    assert _werr(net, onet) <= 1e-5


def test_epoch_with_library_dropout_stream(sz, ctx, oracle):
    onet, net = _pair(sz, ctx, oracle, (60, 512, 256, 7), seed=9)
    r = np.random.default_rng(9)
    n = 1000
    feats = r.standard_normal((n, 60)).astype(np.float32)
    labels = r.integers(0, 7, n).astype(np.uint32)
    data = sz.DeviceFeatures(ctx, feats, labels)
    tot, cnt = 0.0, 0
    for epoch in range(2):
        perm = r.permutation(n).astype(np.uint32)
        loss, used = sz.train_epoch(net, data, perm, 96, 0.02, dropout=0.2, seed=77, stream=epoch)   # 96 does not divide 1000
        keep = oracle.dropout_keep_mask(77, epoch, np.arange(n), 60, 0.2)
        oloss, oused = oracle.train_epoch(onet, feats, labels, perm, 96, 0.02, keep)
        assert used == oused == n
        assert abs(loss - oloss) <= 1e-3 * abs(oloss)
    data.close()
    assert _werr(net, onet) <= 1e-4


def test_pretrain_from_features_reduces_loss(sz, ctx, oracle):
    # C1-shaped: windows of two synthetic speakers, file-sequential training as train_from_feature_map (lib.rs:632-665)
    ex = sz.FeatureExtractor(ctx)
    fmap = {f"spk{s}_{i}.wav": ex.extract(oracle.synth_clip(s, 10 * s + i, 1.0)) for s in range(2) for i in range(2)}
    files = [(p, int(p[3])) for p in sorted(fmap)]
    net = sz.SimpleNeuralNet(60, 512, 256, 2, seed=1, ctx=ctx)
    first = sz.train_from_feature_map(net, fmap, files, 1, 0.01, 0.2, 8, seed=1)
    later = sz.train_from_feature_map(net, fmap, files, 3, 0.01, 0.2, 8, seed=2)
    assert np.isfinite(first) and later < first
    assert net.file_lists() == [["spk0_0.wav", "spk0_1.wav"], ["spk1_0.wav", "spk1_1.wav"]]
    assert sz.pretrain_from_features(net, np.zeros((0, 60), np.float32), 0, 2, 3, 0.01, 0.2, 8) == 0.0   # lib.rs:623-627


def test_identify_counts_sums_and_list(sz, ctx, oracle):
    onet, net = _pair(sz, ctx, oracle, (60, 512, 256, 6), seed=12)
    clip = np.concatenate([oracle.synth_clip(s, 40 + s, 2.0) for s in range(3)])
    feats = oracle.extract(clip).astype(np.float32)
    p64 = oracle.forward(onet.copy(np.float64), feats)
    thr = 0.5
    near = np.abs(p64.max(axis=1) - thr) < 1e-4
    srt = np.sort(p64, axis=1)
    near |= (srt[:, -1] - srt[:, -2]) < 1e-4
    assert near.sum() == 0, "test data has a near-tie; change the seed"
    counts = sz.identify_counts(net, feats, thr)
    assert np.array_equal(counts, oracle.identify_counts(onet, feats, thr))
    sums = sz.identify_sums(net, feats)
    assert np.abs(sums - oracle.identify_sums(onet, feats)).max() <= 1e-3
    ex = sz.FeatureExtractor(ctx)
    assert sz.identify_speaker_list(net, clip, thr) == oracle.identify_speaker_list(onet, clip, thr)
    assert sz.identify_speaker_list(net, clip[:500], thr) == []                      # no windows -> empty list
    assert sz.identify_speaker(net, clip, ex) == oracle.identify_speaker(onet, feats)
    assert sz.identify_speaker_with_threshold(net, clip, 0.0, ex) == oracle.identify_speaker_with_threshold_feats(onet, feats, 0.0)
    assert sz.identify_speaker_with_threshold(net, clip, 1.1, ex) is None
    one = sz.SimpleNeuralNet(60, 512, 256, 1, ctx=ctx)
    assert sz.identify_speaker_with_threshold(one, clip, 0.0, ex) is None            # lib.rs:1316-1318
    assert sz.identify_speaker_with_threshold_feats(net, feats[:0], 0.0) is None     # lib.rs:1363-1365


def test_tie_break_takes_last_index(sz, ctx):
    # identical output columns -> equal probabilities -> the LAST index is counted (lib.rs:1393-1396)
    r = np.random.default_rng(0)
    w3 = np.repeat(r.uniform(-.5, .5, (256, 1)).astype(np.float32), 4, axis=1)
    net = sz.SimpleNeuralNet.from_weights(r.uniform(-.5, .5, (60, 512)), np.zeros(512), r.uniform(-.5, .5, (512, 256)), np.zeros(256),
                                          w3, np.zeros(4), ctx=ctx)
    feats = r.standard_normal((50, 60)).astype(np.float32)
    assert sz.identify_counts(net, feats, 0.25).tolist() == [0, 0, 0, 50]
    assert sz.identify_counts(net, feats, 0.26).tolist() == [0, 0, 0, 0]             # `>=` threshold (lib.rs:1398)


def test_add_output_class(sz, ctx, oracle):
    onet, net = _pair(sz, ctx, oracle, (60, 512, 256, 3), seed=2)
    col = np.random.default_rng(5).uniform(-.5, .5, 256).astype(np.float32)
    net.add_output_class(col)
    w = net.weights()
    assert net.output_size() == 4 and w[4].shape == (256, 4) and w[5].shape == (4,)
    assert np.array_equal(w[4][:, :3], onet.w3) and np.array_equal(w[4][:, 3], col)  # lib.rs:803-809
    assert np.array_equal(w[5][:3], onet.b3) and w[5][3] == 0                         # lib.rs:812-815
    net.add_output_class()                                                            # random column in [-0.5, 0.5)
    w = net.weights()
    assert net.output_size() == 5 and np.abs(w[4][:, 4]).max() <= 0.5 and np.abs(w[4][:, 4]).max() > 0
    x = np.random.default_rng(1).standard_normal((9, 60)).astype(np.float32)
    assert net.forward(x).shape == (9, 5)


def test_model_npz_round_trip_and_numpy_compat(sz, ctx, oracle, tmp_path):
    onet, net = _pair(sz, ctx, oracle, (60, 512, 256, 5), seed=8)
    net.record_training_file(0, "a/x.wav"); net.record_training_file(0, "a/y.wav"); net.record_training_file(0, "a/x.wav")
    net.record_training_file(3, "b/z.mp3")
    path = str(tmp_path / "model.npz")
    net.save(path)
    z = np.load(path)                                                                # numpy reads the reference's layout
    names = set(z.files)
    assert {"w1", "b1", "w2", "b2", "sample_rate", "bits", "num_speakers", "w3_1", "b3_5", "speaker_0_files"} <= names
    assert z["w1"].shape == (60, 512) and z["w1"].dtype == np.float32 and z["sample_rate"].dtype == np.int64
    assert int(z["num_speakers"][0]) == 5 and int(z["sample_rate"][0]) == 44100 and int(z["bits"][0]) == 16
    for k in range(5):
        assert np.array_equal(z[f"w3_{k + 1}"], onet.w3[:, k]) and z[f"b3_{k + 1}"][0] == onet.b3[k]   # lib.rs:1091-1098
    assert bytes(z["speaker_0_files"]).decode() == "a/x.wav\na/y.wav" and len(z["speaker_1_files"]) == 0
    back = sz.SimpleNeuralNet.load(path, ctx=ctx)
    assert all(np.array_equal(a, b) for a, b in zip(back.weights(), net.weights()))
    assert back.file_lists()[0] == ["a/x.wav", "a/y.wav"] and back.file_lists()[3] == ["b/z.mp3"]
    # numpy-written files: members get a ".npy" suffix; legacy dense w3/b3 (lib.rs:1199-1207)
    p2 = str(tmp_path / "legacy.npz")
    np.savez(p2, w1=onet.w1, b1=onet.b1, w2=onet.w2, b2=onet.b2, w3=onet.w3, b3=onet.b3, sample_rate=np.array([16000], np.int64),
             bits=np.array([16], np.int64))
    leg = sz.SimpleNeuralNet.load(p2, ctx=ctx)
    assert leg.sample_rate == 16000 and all(np.array_equal(a, b) for a, b in zip(leg.weights(), onet.params()))
    with pytest.raises(Exception):
        sz.SimpleNeuralNet.load(str(tmp_path / "missing.npz"), ctx=ctx)


def test_cached_features(sz, ctx, oracle, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    ex = sz.FeatureExtractor(ctx)
    clip = oracle.synth_clip(1, 1, 0.5)
    calls = []
    def loader(p):
        calls.append(p)
        return clip
    a = sz.load_cached_features("d/s.wav", loader, ex)
    b = sz.load_cached_features("d/s.wav", loader, ex)
    assert len(calls) == 1 and np.array_equal(a, b)
    assert np.array_equal(np.load(tmp_path / "feature_cache" / "d_s.wav.npy"), a)     # lib.rs:550-579
    assert len(sz.load_cached_features("short.wav", lambda p: clip[:100], ex)) == 0
    assert not (tmp_path / "feature_cache" / "short.wav.npy").exists()                # empty sets are not written (lib.rs:573)


def test_model_npz_keeps_embeddings_and_encoding_layer(sz, ctx, oracle, tmp_path):
    # lib.rs:1099-1127 (save) / 1168-1264 (load): speaker_embeddings, speaker_mean_sims, speaker_std_sims, w4_k / b4_k.
    # The CLI sets the embeddings right before saving (main.rs:845-856); load -> save must lose nothing.
    onet, net = _pair(sz, ctx, oracle, (60, 512, 256, 3), seed=4)
    r = np.random.default_rng(3)
    embeds = [(r.standard_normal(256).astype(np.float32), float(r.random()), float(r.random())) for _ in range(3)]
    net.set_embeddings(embeds)
    w4, b4 = r.standard_normal((40, 6)).astype(np.float32), r.standard_normal(6).astype(np.float32)   # its own row count (lib.rs:1211-1214)
    net.set_encoding_layer(w4, b4)
    path = str(tmp_path / "model.npz")
    net.save(path)
    z = np.load(path)
    assert z["speaker_embeddings"].shape == (3, 256) and z["speaker_embeddings"].dtype == np.float32
    assert np.array_equal(z["speaker_embeddings"], np.stack([e[0] for e in embeds]))
    assert np.array_equal(z["speaker_mean_sims"], np.array([e[1] for e in embeds], np.float32))
    assert np.array_equal(z["speaker_std_sims"], np.array([e[2] for e in embeds], np.float32))
    for k in range(6):
        assert np.array_equal(z[f"w4_{k + 1}"], w4[:, k]) and z[f"b4_{k + 1}"][0] == b4[k]
    # member order of the reference writer (lib.rs:1084-1127)
    order = list(z.files)
    assert order.index("b3_3") < order.index("w4_1") < order.index("speaker_0_files") < order.index("speaker_embeddings")
    back = sz.SimpleNeuralNet.load(path, ctx=ctx)
    got = back.embeddings()
    assert len(got) == 3 and all(np.array_equal(a[0], b[0]) and a[1] == np.float32(b[1]) and a[2] == np.float32(b[2])
                                 for a, b in zip(got, embeds))
    gw4, gb4 = back.encoding_layer()
    assert np.array_equal(gw4, w4) and np.array_equal(gb4, b4)
    p2 = str(tmp_path / "again.npz")
    back.save(p2)
    assert open(path, "rb").read() == open(p2, "rb").read()          # load -> save is the identity on the file
    fresh = sz.SimpleNeuralNet(60, 8, 4, 2, ctx=ctx)
    assert fresh.embeddings() == [] and fresh.encoding_layer() is None
    p3 = str(tmp_path / "plain.npz")
    fresh.save(p3)
    assert "speaker_embeddings" not in np.load(p3).files and "w4_1" not in np.load(p3).files   # only when non-empty (lib.rs:1114)


def test_npz_and_npy_readers_reject_malformed_files(sz, ctx, native, tmp_path):
    # every one of these used to over-read, wrap an index or throw through the C ABI
    import ctypes as C
    import struct
    good = str(tmp_path / "m.npz")
    sz.SimpleNeuralNet(60, 8, 4, 2, ctx=ctx).save(good)
    raw = bytearray(open(good, "rb").read())

    def load_status(data):
        p = str(tmp_path / "bad.npz")
        open(p, "wb").write(bytes(data))
        h, sr, bits = C.c_void_p(), C.c_uint32(), C.c_uint32()
        st = native.lib.szb_net_load(ctx.handle, p.encode(), C.byref(h), C.byref(sr), C.byref(bits))
        if st == 0:
            native.lib.szb_net_destroy(h)
        return st

    assert load_status(raw) == 0
    eocd = raw.rfind(b"PK\x05\x06")
    cd = struct.unpack_from("<I", raw, eocd + 16)[0]
    bad = bytearray(raw); struct.pack_into("<H", bad, cd + 28, 0xFFFF)          # name length past the end of the file
    assert load_status(bad) == native.ERR_IO
    bad = bytearray(raw); struct.pack_into("<I", bad, cd + 42, 0xFFFFFFF0)      # local header offset that wraps in 32 bits
    assert load_status(bad) == native.ERR_IO
    bad = bytearray(raw); struct.pack_into("<I", bad, cd + 20, 0xFFFFFF00); struct.pack_into("<I", bad, cd + 24, 0xFFFFFF00)
    assert load_status(bad) == native.ERR_IO                                      # member size past the end
    assert load_status(raw[:40]) == native.ERR_IO

    def npy_status(header):
        p = str(tmp_path / "bad.npy")
        h = header.encode()
        open(p, "wb").write(b"\x93NUMPY\x01\x00" + struct.pack("<H", len(h)) + h + b"\x00" * 64)
        rows, cols = C.c_uint64(), C.c_uint64()
        buf = np.zeros(16, np.float32)
        return native.lib.szb_npy_read_f32(p.encode(), native.ptr(buf), 16, C.byref(rows), C.byref(cols))

    assert npy_status("{'descr': '<f4', 'fortran_order': False, 'shape': (2, 8), }\n") == 0
    assert npy_status("{'descr': '<f4', 'fortran_order': False, 'shape': (4611686018427387904, 4), }\n") == native.ERR_IO   # count * 4 wraps
    assert npy_status("{'descr' '<f4', 'fortran_order' False, 'shape' (2, 8), }\n") == native.ERR_IO                      # no ':' at all
    assert npy_status("{'descr': '<f4', 'fortran_order':") == native.ERR_IO
    assert npy_status("{'descr': '<f4', 'fortran_order': False, 'shape': (2, 8") == native.ERR_IO


def test_batched_identification_equals_per_clip_histograms(sz, ctx, oracle):
    # szb_identify_counts_batch_dev: one pass over the concatenated windows of many clips, per-window clip lookup;
    # counts[c] must equal identify_counts of clip c (lib.rs:1389-1402), empty clips included
    r = np.random.default_rng(21)
    onet, net = _pair(sz, ctx, oracle, (60, 512, 256, 7), seed=13)
    sizes = [0, 1, 37, 0, 1101, 5, 64, 0, 300, 0]
    clips = [r.standard_normal((n, 60)).astype(np.float32) for n in sizes]
    for thr in (0.0, 0.3, 0.9):
        got = sz.identify_counts_batch(net, clips, thr)
        assert got.shape == (len(sizes), 7) and got.dtype == np.uint32
        for c, w in enumerate(clips):
            want = sz.identify_counts(net, w, thr) if len(w) else np.zeros(7, np.int64)
            assert np.array_equal(got[c].astype(np.int64), want), (thr, c)
        assert np.array_equal(got.astype(np.int64).sum(axis=0), oracle.identify_counts(onet, np.concatenate(clips), thr)) or thr > 0.0
    assert sz.identify_counts_batch(net, [], 0.5).shape == (0, 7)


def test_small_batch_epochs_replay_a_captured_graph_with_identical_results(sz, ctx, oracle, native, monkeypatch):
    # Batches of 33..256 rows are launch-latency-bound: the epoch replays a captured two-step CUDA graph.
    # Same kernels, same order: the weights must equal those of plain launches (SZB_NO_GRAPHS=1 context) to within the order of
    # the split-K reductions, and the oracle's to 1e-4, over several epochs with odd step counts and a ragged last batch.
    r = np.random.default_rng(3)
    BATCH = 48
    n = BATCH * 37 + 5
    feats = r.standard_normal((n, 60)).astype(np.float32)
    labels = r.integers(0, 3, n).astype(np.uint32)
    onet = oracle.Net.init(60, 512, 256, 3, seed=2)
    perms = [r.permutation(n).astype(np.uint32) for _ in range(3)]
    monkeypatch.setenv("SZB_NO_GRAPHS", "1")
    plain_ctx = sz.Context(0)
    monkeypatch.delenv("SZB_NO_GRAPHS")
    results = []
    for c in (ctx, plain_ctx):
        net = sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=c)
        data = sz.DeviceFeatures(c, feats, labels)
        tot, cnt = 0.0, 0
        launches0 = c.launch_count
        for e, perm in enumerate(perms):
            loss, used = sz.train_epoch(net, data, perm, BATCH, 0.01 * 0.99 ** e, dropout=0.2, seed=5, stream=e)   # lr and key change per epoch
            tot += loss; cnt += used
        results.append((net.weights(), tot, cnt, c.launch_count - launches0, int(native.lib.szb_ctx_graph_launch_count(c.handle))))
        data.close()
        net.close()                                            # a net must not outlive its context
    (wg, lg, cg, ng, gg), (wp, lp, cp, npl, gp) = results
    assert gg > 0 and gp == 0                                  # the default context really replayed graphs, the other one did not
    assert cg == cp and abs(lg - lp) <= 1e-4 * abs(lp) and ng == npl                      # same work, same launch count claimed
    assert max(float(np.abs(a - b).max()) for a, b in zip(wg, wp)) <= 1e-6
    for e, perm in enumerate(perms):
        keep = oracle.dropout_keep_mask(5, e, np.arange(n), 60, 0.2)
        oracle.train_epoch(onet, feats, labels, perm, BATCH, 0.01 * 0.99 ** e, keep)
    assert max(float(np.abs(a - b).max()) for a, b in zip(wg, onet.params())) <= 1e-4
    plain_ctx.close()


def test_persistent_small_batch_kernel_equals_the_step_kernels_and_the_oracle(sz, ctx, oracle, native, monkeypatch):
    # Batches of <= 32 rows (the reference's default is 8, main.rs:36): the epoch runs as ONE persistent cooperative kernel
    # (train_small.cu, FP32 CUDA cores, grid barriers between the layers).  Against the oracle: <= 1e-5 after one step, <= 1e-4
    # after epochs; against the eleven-kernel path (SZB_NO_SMALL_KERNEL=1) within FP32 reassociation; ragged last batches, batch
    # sizes 1 / 8 / 32, windows dropped to all-zero, labels out of range, the reference's 4-3-2-2 net, 7 and 200 classes.
    monkeypatch.setenv("SZB_NO_SMALL_KERNEL", "1")
    plain_ctx = sz.Context(0)
    monkeypatch.delenv("SZB_NO_SMALL_KERNEL")
    r = np.random.default_rng(8)
    for dims, batch, n, prob in (((60, 512, 256, 2), 8, 8 * 9 + 3, 0.2), ((60, 512, 256, 7), 32, 32 * 3 + 1, 0.2), ((60, 512, 256, 200), 1, 5, 0.0),
                                 ((4, 3, 2, 2), 8, 19, 0.5), ((60, 512, 256, 3), 8, 8, 0.97)):
        feats = r.standard_normal((n, dims[0])).astype(np.float32)
        labels = r.integers(0, dims[3] + 1, n).astype(np.uint32)          # includes label == C: all-zero target (lib.rs:592-595)
        base = oracle.Net.init(*dims, seed=4)
        outs = []
        for c in (ctx, plain_ctx):
            net = sz.SimpleNeuralNet.from_weights(*base.params(), ctx=c)
            data = sz.DeviceFeatures(c, feats, labels)
            l0 = c.launch_count
            tot, cnt = 0.0, 0
            for e in range(2):
                perm = np.random.default_rng(e).permutation(n).astype(np.uint32)
                loss, used = sz.train_epoch(net, data, perm, batch, 0.05, dropout=prob, seed=3, stream=e)
                tot += loss; cnt += used
            outs.append((net.weights(), tot, cnt, c.launch_count - l0, net.forward(feats[:4])))
            data.close(); net.close()
        onet = base.copy()
        otot, ocnt = 0.0, 0
        for e in range(2):
            perm = np.random.default_rng(e).permutation(n).astype(np.uint32)
            keep = oracle.dropout_keep_mask(3, e, np.arange(n), dims[0], prob)
            l, k = oracle.train_epoch(onet, feats, labels, perm, batch, 0.05, keep)
            otot += l; ocnt += k
        (ws, ls, cs, launches_small, ps), (wp, lp, cp, launches_plain, pp) = outs
        assert launches_small == 2 and launches_plain > 2, (launches_small, launches_plain)      # one launch per epoch
        assert cs == cp == ocnt, (dims, cs, cp, ocnt)
        assert abs(ls - otot) <= 1e-4 * max(1.0, abs(otot)) and abs(lp - otot) <= 1e-4 * max(1.0, abs(otot))
        assert max(float(np.abs(a - b).max()) for a, b in zip(ws, onet.params())) <= 1e-4, dims
        assert max(float(np.abs(a - b).max()) for a, b in zip(ws, wp)) <= 1e-4, dims
        assert np.abs(ps - oracle.forward(onet, feats[:4])).max() <= 1e-4        # the tensor-core forward sees the updated weights
    plain_ctx.close()


def test_fused_step_launch_structure_equals_the_eleven_launch_step_and_the_oracle(sz, ctx, oracle, monkeypatch):
    # Round 2's step was eleven launches (batch kernel, 3 forward GEMMs, softmax, dW3, dX2, dW2, dX1, dW1, update).  SZB_STEP_FUSE
    # bits fold three of them away: 1 = softmax / cross-entropy in the layer-3 epilogue (n_out <= 128), 2 = the next batch
    # prepared by extra CTAs of the update kernel (batches > 256 rows), 4 = input gradients first, then the three weight
    # gradients as one grouped launch.  Every combination must give the eleven-launch result (to FP32 reassociation: the
    # split-K ranges and the order of the softmax sum differ) and the oracle's, with ragged last batches (77 rows: rows of the
    # transposed operands are no longer 16-byte aligned, the grouped launch falls back), windows dropped to all-zero, labels
    # out of range, and more classes than one tile holds (200: softmax stays a kernel of its own).
    r = np.random.default_rng(12)
    for dims, batch, n in (((60, 512, 256, 100), 300, 300 * 3 + 77), ((60, 512, 256, 7), 260, 260 * 2 + 128), ((60, 512, 256, 200), 512, 1024 + 260)):
        feats = r.standard_normal((n, dims[0])).astype(np.float32)
        feats[5] = 0                                                            # a window that is all-zero before dropout
        labels = r.integers(0, dims[3] + 1, n).astype(np.uint32)                # includes label == C: all-zero target (lib.rs:592-595)
        base = oracle.Net.init(*dims, seed=6)
        perms = [r.permutation(n).astype(np.uint32) for _ in range(2)]
        outs = {}
        for bits in (0, 1, 2, 4, 7):
            monkeypatch.setenv("SZB_STEP_FUSE", str(bits))
            c = sz.Context(0)
            net = sz.SimpleNeuralNet.from_weights(*base.params(), ctx=c)
            data = sz.DeviceFeatures(c, feats, labels)
            l0 = c.launch_count
            tot, cnt = 0.0, 0
            for e, perm in enumerate(perms):
                loss, used = sz.train_epoch(net, data, perm, batch, 0.02, dropout=0.2, seed=11, stream=e)
                tot += loss; cnt += used
            outs[bits] = (net.weights(), tot, cnt, c.launch_count - l0)
            data.close(); net.close(); c.close()
        monkeypatch.delenv("SZB_STEP_FUSE")
        onet = base.copy()
        otot, ocnt = 0.0, 0
        for e, perm in enumerate(perms):
            keep = oracle.dropout_keep_mask(11, e, np.arange(n), dims[0], 0.2)
            l, k = oracle.train_epoch(onet, feats, labels, perm, batch, 0.02, keep)
            otot += l; ocnt += k
        w0, l0_, c0, n0 = outs[0]
        for bits in (1, 2, 4, 7):
            w, l, c, nl = outs[bits]
            assert c == c0 == ocnt, (dims, bits, c, c0, ocnt)
            assert abs(l - l0_) <= 1e-5 * abs(l0_) and abs(l - otot) <= 1e-3 * abs(otot), (dims, bits)
            assert max(float(np.abs(a - b).max()) for a, b in zip(w, w0)) <= 2e-6, (dims, bits)
            assert max(float(np.abs(a - b).max()) for a, b in zip(w, onet.params())) <= 1e-4, (dims, bits)
            assert nl < n0 or (bits == 1 and dims[3] > 128), (dims, bits, nl, n0)      # launches really went away
        assert outs[7][3] <= min(outs[b][3] for b in (1, 2, 4))
