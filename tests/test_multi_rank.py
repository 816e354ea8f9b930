"""World-size-2 checks of the multi-GPU decomposition on the CPU (gloo): clip sharding, and the identity the NCCL path
relies on -- all-reducing [gradient sum | surviving-window count] over batch slices reproduces the single-process step."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_clips_covers_everything_once():
    sys.path.insert(0, ROOT)
    from streamz_b200.sharding import shard_clips
    r = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        lengths = r.integers(1000, 500000, 1000)
        parts = shard_clips(lengths, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(lengths)
        assert all(a[1] == b[0] for a, b in zip(parts[:-1], parts[1:]))
        loads = [int(lengths[a:b].sum()) for a, b in parts]
        assert max(loads) - min(loads) <= 2 * lengths.max()
    assert shard_clips([5, 5], 4)[-1][1] == 2            # more ranks than clips: some ranks get nothing


def test_shard_windows_partitions_a_clip():
    from streamz_b200.sharding import shard_windows
    for n in (0, 1, 5, 550, 6614):
        for world in (1, 2, 3, 4, 8):
            parts = shard_windows(n, world)
            assert parts[0][0] == 0 and parts[-1][1] == n and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def test_shard_batches_partitions_each_global_batch():
    from streamz_b200.sharding import shard_batches
    perm = np.random.default_rng(1).permutation(1000)
    for world in (2, 3, 8):
        locs = [shard_batches(perm, 96, r, world) for r in range(world)]
        assert len({len(s) for _, s in locs}) == 1       # same number of steps on every rank
        for step, s0 in enumerate(range(0, 1000, 96)):
            got = np.concatenate([loc[sum(sz[:step]):sum(sz[:step + 1])] for loc, sz in locs])
            assert np.array_equal(np.sort(got), np.sort(perm[s0:s0 + 96]))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import streamz_oracle as orc
    from streamz_b200.sharding import shard_batches
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = np.random.default_rng(7)
    n = 300
    feats = r.standard_normal((n, 60)).astype(np.float32); feats[11] = 0
    labels = r.integers(0, 5, n).astype(np.uint32)
    perm = r.permutation(n).astype(np.uint32)
    keep = orc.dropout_keep_mask(3, 0, np.arange(n), 60, 0.2)
    net = orc.Net.init(60, 32, 16, 5, seed=2)
    local, sizes = shard_batches(perm, 64, rank, world)
    pos = 0
    for b in sizes:
        idx = local[pos:pos + b].astype(np.int64); pos += b
        x = np.where(keep[idx], feats[idx], np.float32(0))
        alive = ~np.all(x == 0, axis=1)
        x, lab = x[alive], labels[idx][alive]
        if len(x):
            grads, _ = orc.gradients(net, x, orc.one_hot(lab, 5))
        else:
            grads = [np.zeros_like(p) for p in net.params()]
        flat = torch.from_numpy(np.concatenate([g.ravel() for g in grads] + [np.array([len(x)], np.float32)]))
        dist.all_reduce(flat)                                       # the one collective of the path
        flat = flat.numpy()
        used = flat[-1]
        if used > 0:
            o = 0
            for p in net.params():
                p -= (flat[o:o + p.size].reshape(p.shape) * np.float32(0.05 / used)).astype(np.float32)
                o += p.size
    if rank == 0:
        np.savez(out_path, *net.params())
    dist.destroy_process_group()


def test_allreduced_slices_equal_single_process_step(tmp_path, oracle):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "dp.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = np.random.default_rng(7)
    n = 300
    feats = r.standard_normal((n, 60)).astype(np.float32); feats[11] = 0
    labels = r.integers(0, 5, n).astype(np.uint32)
    perm = r.permutation(n).astype(np.uint32)
    keep = oracle.dropout_keep_mask(3, 0, np.arange(n), 60, 0.2)
    net = oracle.Net.init(60, 32, 16, 5, seed=2)
    oracle.train_epoch(net, feats, labels, perm, 64, 0.05, keep)
    dp = np.load(out)
    for i, p in enumerate(net.params()):
        assert np.abs(dp[f"arr_{i}"] - p).max() < 2e-6
