"""the CLI's training run (streamz-rs/src/main.rs:490-520, 651-668, 750-835) on top of the C ABI, against
the oracle's restatement of the same flow with the same shuffles, dropout masks and new output columns.

Decisions (which class every file ends up in, when classes are added) must be identical; losses and weights agree to the
tolerances of the step kernels accumulated over the ~700 optimiser steps of the run."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scenario(oracle, rate=44100):
    # (speaker, label in train_files.txt): two labelled files, then unlabelled files of known and unknown speakers
    plan = [(0, 0), (1, 1), (0, None), (2, None), (1, None), (0, None), (2, None), (1, None)]
    clips, files = {}, []
    for i, (spk, label) in enumerate(plan):
        path = f"clip{i}_spk{spk}.wav"
        clips[path] = (oracle.synth_clip(spk, 500 + i, 1.25, rate=rate), rate)
        files.append([path, label])
    clips["short.wav"] = (oracle.synth_clip(1, 900, 0.05, rate=rate), rate)      # 3 windows: skipped (main.rs:756)
    files.insert(4, ["short.wav", None])
    files.append(["missing.wav", None])                                          # no audio: skipped (main.rs:829)
    return clips, files


def test_burn_in_limit_and_helpers(sz, oracle):
    from streamz_b200 import orchestrate as orc
    for n in (1, 10, 49, 51, 100, 249, 251, 5000):
        assert orc.burn_in_limit(n) == oracle.burn_in_limit(n) == min(50, max(10, int(np.ceil(np.float32(n) * np.float32(0.2)))))
    assert orc.burn_in_limit(1000, override=7) == 7
    assert orc.count_speakers([("a", 3), ("b", None), ("c", 3), ("d", 0)]) == 2
    v = [np.array([3.0, 4.0], np.float32), np.array([0.0, 2.0], np.float32)]
    assert np.allclose(orc.average_vectors(v), oracle.average_vectors(v), atol=1e-7)


def test_training_run_matches_the_oracle_flow(sz, ctx, oracle):
    from streamz_b200 import orchestrate as orc
    clips, files = _scenario(oracle)
    ex = sz.FeatureExtractor(ctx)
    fmap = orc.extract_feature_map(clips, ex)
    assert set(fmap) == set(clips) and fmap["short.wav"].shape == (oracle.n_windows(len(clips["short.wav"][0])), 60)
    for p, (pcm, _) in list(clips.items())[:3]:                                   # one batched call == per-clip extraction
        assert np.array_equal(fmap[p], ex.extract(pcm))

    onet = oracle.Net.init(60, 512, 256, 2, seed=21)
    net = sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx)
    onet = onet.copy(dtype=np.float64)
    ofmap = {p: f.astype(np.float64) for p, f in fmap.items()}
    new_cols = [np.random.default_rng(70 + i).uniform(-0.5, 0.5, 256).astype(np.float32) for i in range(8)]
    seed, init_epochs, limit = 5, 6, 3

    # initial training over the labelled files (main.rs:651-668)
    loss0 = orc.initial_training(net, fmap, files, epochs=init_epochs, seed=seed)
    oloss, k = 0.0, 0
    for p, c in files:
        if c is not None:
            oloss += oracle.pretrain_from_features(onet, ofmap[p], c, init_epochs, 0.01, 0.2, 8, seed + k)
            k += 1
    assert k == 2 and abs(loss0 - oloss / k) <= 2e-3 * max(1.0, abs(oloss / k))

    # incremental pass with burn-in (main.rs:750-835)
    gfiles, ofiles = [list(f) for f in files], [list(f) for f in files]
    got = orc.incremental_training(net, gfiles, fmap, limit, seed=seed, new_columns=new_cols)
    want = oracle.incremental_training(onet, ofiles, ofmap, limit, 0.8, 0.2, 8, 5, seed, new_cols)
    assert [(p, s) for p, _, s in got["log"] if s is not None] == want["log"]     # same class for every file, same growth
    assert [f[1] for f in gfiles] == [f[1] for f in ofiles]                       # train_files.txt would be rewritten alike
    hows = {p: h for p, h, _ in got["log"]}
    assert hows["short.wav"] == "too short" and hows["missing.wav"] == "missing"
    assert hows[files[2][0]] == "new (burn-in)"                                   # unlabelled file inside the burn-in phase
    assert got["count"] == want["count"] == 8 and net.output_size() == onet.n_out
    assert abs(got["total_loss"] - want["total_loss"]) <= 5e-3 * max(1.0, abs(want["total_loss"]))
    for a, b in zip(net.weights(), onet.params()):
        assert np.abs(a - b).max() <= 2e-3
    for sid, e in want["speaker_embeddings"].items():
        assert abs(float(np.dot(got["speaker_embeddings"][sid], e)) - 1.0) <= 1e-3
    # the files of one synthetic speaker that were matched (not burn-in) share a class
    by_spk = {}
    for (p, how, s) in got["log"]:
        if how == "matched":
            by_spk.setdefault(p.split("_spk")[1], set()).add(s)
    assert all(len(v) == 1 for v in by_spk.values())

    # evaluation by cosine similarity to the stored embeddings (main.rs:583-640) runs on the same pieces
    targets = [(p, s) for p, _, s in got["log"] if s is not None]
    rep = orc.evaluate(net, got["speaker_embeddings"], targets, fmap, conf_threshold=0.5)
    assert all(0.0 <= rep[k] <= 1.0 for k in ("accuracy", "precision", "recall", "f1"))
