"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/streamz_b200.h declares,
its GPU-free entry points (size queries, tables, dropout stream, npy IO) agree with the oracle, and it fails loudly
without a GPU instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, P


def _declared():
    text = open(os.path.join(ROOT, "include", "streamz_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(szb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native):
    names = _declared()
    assert len(names) >= 50
    for n in names:
        assert hasattr(native.lib, n), f"{n} is declared in include/streamz_b200.h but not exported"
    assert set(native.SIGNATURES) == set(names), "ctypes signature table and header disagree"
    assert b"sm_100a" in native.lib.szb_version()


def test_size_queries(native, oracle):
    for n in (0, 799, 800, 1199, 1200, 441000, 2646000):
        assert native.lib.szb_num_windows(n) == oracle.n_windows(n)
    for n, rate in ((160000, 16000), (12345, 32000), (999, 48000), (7, 8000), (0, 16000)):
        assert native.lib.szb_resample_out_len(n, rate) == oracle.resample_out_len(n, rate)
    off = np.array([0, 160000, 160000, 160100, 480100], dtype=np.uint64)   # ragged, one empty and one too-short clip
    want = sum(oracle.n_windows(oracle.resample_out_len(int(b - a), 16000)) for a, b in zip(off[:-1], off[1:]))
    assert native.lib.szb_extract_batch_windows(P(off), 4, 16000) == want
    assert native.lib.szb_extract_batch_windows(P(off), 4, 44100) == sum(oracle.n_windows(int(b - a)) for a, b in zip(off[:-1], off[1:]))


def test_tables_match_oracle(native, oracle):
    mel = np.zeros((26, 401), np.float32); native.check(native.lib.szb_table_mel(P(mel)))
    dct = np.zeros((20, 26), np.float32); native.check(native.lib.szb_table_dct(P(dct)))
    assert np.array_equal(mel, oracle.mel_filterbank())
    assert np.array_equal(dct, oracle.dct2_matrix(dtype=np.float32))
    for rate in (8000, 11025, 16000, 22050, 32000, 48000, 96000):
        L, M = C.c_uint32(), C.c_uint32()
        native.check(native.lib.szb_table_resample_taps(rate, None, C.byref(L), C.byref(M)))
        assert (L.value, M.value) == oracle.resample_ratio(rate)
        taps = np.zeros((L.value, 16), np.float32)
        native.check(native.lib.szb_table_resample_taps(rate, P(taps), None, None))
        ref = oracle.resample_taps(rate)
        # two independent double-precision evaluations: equal up to the last float32 bit
        assert np.abs(taps - ref).max() <= 1.2e-7 * np.abs(ref).max()


def test_dropout_stream_matches_oracle(native, oracle, sz):
    rows = np.array([0, 1, 5, 123456, 2 ** 31 + 3], dtype=np.uint64)
    for seed, stream, p in ((0, 0, 0.2), (42, 7, 0.5), (2 ** 63 + 11, 3, 0.01)):
        got = sz.dropout_keep_mask(seed, stream, rows, 60, p)
        want = oracle.dropout_keep_mask(seed, stream, rows, 60, p)
        assert np.array_equal(got, want)
    big = sz.dropout_keep_mask(1, 0, np.arange(4000), 60, 0.2)
    assert abs((~big).mean() - 0.2) < 0.01            # drop rate = p, no rescale (lib.rs:119-129)
    assert sz.dropout_keep_mask(1, 0, np.arange(10), 60, 0.0).all()


def test_npy_round_trip_against_numpy(sz, tmp_path):
    a = np.random.default_rng(0).standard_normal((37, 60)).astype(np.float32)
    path = str(tmp_path / "f.npy")
    sz.write_npy(path, a)
    assert np.array_equal(np.load(path), a)            # numpy reads what we write (feature_cache layout, lib.rs:574-576)
    np.save(path, a[:5])
    assert np.array_equal(sz.read_npy(path), a[:5])    # we read what numpy writes
    np.save(path, a.astype(np.float64))
    with pytest.raises(Exception):
        sz.read_npy(path)                              # wrong dtype is an error, like read_npy::<Array2<f32>> (lib.rs:564)
    assert sz.feature_cache_path("data/spk1/a b.wav") == "feature_cache/data_spk1_a b.wav.npy"
    assert sz.feature_cache_path("c:\\x\\y.mp3") == "feature_cache/c:_x_y.mp3.npy"


def test_no_gpu_means_loud_failure_not_fallback(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    st = native.lib.szb_ctx_create(0, None, C.byref(h))
    assert st == native.ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in native.lib.szb_last_error()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "streamz_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "streamz_oracle" not in text.replace("oracle/streamz_oracle.py", "") and "liboracle" not in text, f


def test_seeded_loop_draws_match_the_oracle(native, oracle):
    """Host-side draws of the raw-audio training loops (szb_loop_seed, szb_shuffle_perm, szb_lr_decay, szb_augment_params):
    the library and the oracle must agree bit for bit, or the GPU parity tests of pretrain_network / train_from_files
    (lib.rs:348-397, 668-732) would compare different runs.  No GPU needed: these are plain host functions."""
    import ctypes as C
    for n in (0, 1, 2, 17, 550):
        p = np.zeros(max(1, n), np.uint32)
        native.check(native.lib.szb_shuffle_perm(9, 3, n, native.ptr(p)))
        assert np.array_equal(p[:n], oracle.shuffle_perm(9, 3, n))
        assert sorted(p[:n].tolist()) == list(range(n))
    for step in range(0, 80, 7):
        assert np.float32(native.lib.szb_lr_decay(0.01, step)) == oracle.lr_decay(0.01, step)
    assert abs(float(oracle.lr_decay(0.01, 50)) - 0.01 * 0.99 ** 50) < 1e-8
    for args in [(0, 0, 0), (1, 2, 3), (2 ** 63 + 5, 7, 99)]:
        assert native.lib.szb_loop_seed(*args) == oracle.loop_seed(*args)
    nl, g, sh = C.c_float(), C.c_float(), C.c_uint64()
    for seed in (0, 5, oracle.loop_seed(3, 1, 2)):
        for n in (22050, 300, 0):
            native.check(native.lib.szb_augment_params(seed, n, C.byref(nl), C.byref(g), C.byref(sh)))
            o = oracle.augment_params(seed, n)
            assert (np.float32(nl.value), np.float32(g.value), sh.value) == (o[0], o[1], o[2])
