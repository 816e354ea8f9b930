"""Parity of the embedding path (embed / forward_embedding, mean and median embeddings, cosine matching)
against the oracle.  Tolerances: embeddings max-abs 5e-5 (3xTF32 default) on unit-norm vectors; medians are exact order
statistics of the per-window values, so they inherit the per-value tolerance; decisions identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair(sz, ctx, oracle, C=5, seed=3):
    onet = oracle.Net.init(60, 512, 256, C, seed=seed)
    r = np.random.default_rng(seed)
    onet.b1[:] = r.uniform(-.1, .1, 512); onet.b2[:] = r.uniform(-.1, .1, 256)
    return onet, sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx)


@pytest.mark.parametrize("mode", ["3xtf32", "fp32"])
def test_embed_and_forward_embedding(sz, ctx, oracle, mode):
    onet, net = _pair(sz, ctx, oracle)
    net.set_precision(mode)
    x = np.random.default_rng(0).standard_normal((333, 60)).astype(np.float32)
    o64 = onet.copy(np.float64)
    # pre-activations of the second layer reach |z| ~ 20 for N(0,1) inputs and U(-0.5,0.5) weights: the tolerance is
    # relative to that scale (3xTF32 carries ~1e-5 relative error over a 512-term dot product, FP32 ~1e-6)
    rel = {"3xtf32": 2e-5, "fp32": 2e-6}[mode]
    want_relu = oracle.forward_embedding(o64, x)
    scale = float(np.abs(want_relu).max())
    assert np.abs(sz.embed(net, x) - oracle.embed(o64, x)).max() <= rel * scale                # ReLU, tanh  (lib.rs:895-900)
    assert np.abs(sz.forward_embedding(net, x) - want_relu).max() <= rel * scale               # ReLU, ReLU  (lib.rs:1073-1079)
    assert sz.embed(net, x[0]).shape == (256,) and sz.embedding_size(net) == 256


@pytest.mark.parametrize("n", [0, 1, 2, 7, 550, 1101])
def test_mean_and_median_embeddings(sz, ctx, oracle, n):
    onet, net = _pair(sz, ctx, oracle)
    feats = np.random.default_rng(n).standard_normal((n, 60)).astype(np.float32)
    o64 = onet.copy(np.float64)
    got_mean, want_mean = sz.extract_embedding_from_features(net, feats), oracle.embedding_mean(o64, feats)
    assert got_mean.shape == (256,) and np.abs(got_mean - want_mean).max() <= 5e-5
    for relu2 in (True, False):
        got, want = sz.median_embedding_from_features(net, feats, relu2), oracle.embedding_median(o64, feats, relu2)
        assert np.abs(got - want).max() <= 5e-5
    if n:
        assert abs(np.linalg.norm(got_mean) - 1) < 1e-5        # normalised (lib.rs:1473)
    else:
        assert not got_mean.any()                              # no windows: zero vector (lib.rs:1429-1431)


def test_median_is_an_exact_order_statistic(sz, ctx):
    # identity-like net so the embedding equals the input: w1 = I (60 -> first 60 of 512), w2 picks them back
    w1 = np.zeros((60, 512), np.float32); w1[np.arange(60), np.arange(60)] = 1
    w2 = np.zeros((512, 256), np.float32); w2[np.arange(60), np.arange(60)] = 1
    net = sz.SimpleNeuralNet.from_weights(w1, np.zeros(512), w2, np.zeros(256), np.zeros((256, 2)), np.zeros(2), ctx=ctx).set_precision("fp32")
    r = np.random.default_rng(5)
    for n in (4, 5, 1000):
        x = np.abs(r.standard_normal((n, 60))).astype(np.float32)   # positive: ReLU is the identity
        x[:, 3] = 0.25                                              # a column of ties
        got = sz.median_embedding_from_features(net, x, relu2=True)
        med = np.median(x.astype(np.float32), axis=0).astype(np.float32)
        want = np.zeros(256, np.float32); want[:60] = med
        want /= np.linalg.norm(want)
        assert np.abs(got - want).max() <= 1e-6


def test_cosine_matching_rules(sz, ctx, oracle):
    onet, net = _pair(sz, ctx, oracle)
    ex = sz.FeatureExtractor(ctx)
    clips = {s: oracle.synth_clip(s, 70 + s, 1.5) for s in range(3)}
    feats = {f"spk{s}.wav": ex.extract(c) for s, c in clips.items()}
    for s in range(3):
        net.record_training_file(s, f"spk{s}.wav")
    embeds = sz.compute_speaker_embeddings(net, feats)                         # lib.rs:1555-1599
    assert len(embeds) == 5 and embeds[0][0].shape == (256,) and embeds[4][1] == 0.0
    assert all(abs(m - 1) < 1e-5 and s < 1e-5 for _, m, s in embeds[:3])       # one file per speaker: sim 1, std 0
    o64 = onet.copy(np.float64)
    for s in range(3):
        w = feats[f"spk{s}.wav"]
        emb = sz.extract_embedding_from_features(net, w)
        assert abs(sz.cosine_similarity(emb, embeds[s][0]) - oracle.cosine_similarity(emb, embeds[s][0])) < 1e-6
        want = oracle.identify_speaker_cosine_emb(oracle.embedding_mean(o64, w), [(e.astype(np.float64), m, sd) for e, m, sd in embeds], 0.3)
        assert sz.identify_speaker_cosine_feats(net, embeds, w, 0.3) == want
        got = sz.identify_speaker_cosine(net, embeds, clips[s], 0.3, ex)
        want2 = oracle.identify_speaker_cosine_emb(oracle.embedding_median(o64, w.astype(np.float64), relu2=False),
                                                   [(e.astype(np.float64), m, sd) for e, m, sd in embeds], 0.3)
        assert got == want2
    cents = {s: embeds[s][0] for s in range(3)}
    e0 = sz.extract_embedding_from_features(net, feats["spk0.wav"])
    assert sz.identify_speaker_from_embedding(e0, cents, 0.99) == oracle.identify_speaker_from_embedding(e0, cents, 0.99)
    assert sz.identify_speaker_from_embedding(e0, cents, 2.0) is None           # nothing passes -> usize::MAX
    assert sz.identify_speaker_cosine_feats(net, [], feats["spk0.wav"], 0.3) is None
    assert sz.cosine_similarity(np.zeros(4), np.ones(4)) == 0.0                 # lib.rs:1535-1537
