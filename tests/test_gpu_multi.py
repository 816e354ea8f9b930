"""2 / 4 / 8 ranks on as many GPUs (one process per GPU): batch-parallel training must reproduce the single-GPU epoch with
any gradient exchange (NCCL all-reduces, or the one-shot / two-shot / packet peer-memory exchange fused into the update kernel) and leave
bit-identical replicas; clip-sharded extraction must reproduce the single-GPU features.  A world size is skipped when
fewer GPUs are visible (the driver's 1-GPU box skips all three; `gpurun --gpus 8` runs all three, log under profiles/)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpu_count():
    import torch
    return torch.cuda.device_count()


def _worker(rank, world, port, tmp):
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    import streamz_b200 as sz
    from streamz_b200.sharding import shard_batches, shard_clips
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)        # plumbing only: carries the NCCL unique id
    ctx = sz.Context(rank)
    uid = [sz.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)
    d = np.load(os.path.join(tmp, "data.npz"))
    net = sz.SimpleNeuralNet.from_weights(*[d[f"p{i}"] for i in range(6)], ctx=ctx)
    data = sz.DeviceFeatures(ctx, d["feats"], d["labels"])
    tot, cnt = 0.0, 0
    for epoch in range(2):
        local, sizes = shard_batches(d[f"perm{epoch}"], 96, rank, world)
        loss, used = sz.train_epoch_steps(net, data, local, sizes, 0.02, dropout=0.2, seed=77, stream=epoch)
        tot += loss; cnt += used
    # the same two epochs again with the two-shot peer-memory exchange fused into the update kernel instead of NCCL,
    # on the tensor-core path (sgd_fused_kernel) and on the FP32 CUDA-core path (sgd_p2p_kernel)
    w2, w3, w4, w6 = None, None, None, None
    peer = True
    for mode, proto in (("3xtf32", "one-shot"), ("3xtf32", "two-shot"), ("3xtf32", "ll"), ("fp32", "auto")):
        peer = ctx.comm_peer_exchange(True, proto) and peer
        net2 = sz.SimpleNeuralNet.from_weights(*[d[f"p{i}"] for i in range(6)], ctx=ctx).set_precision(mode)
        for epoch in range(2):
            local, sizes = shard_batches(d[f"perm{epoch}"], 96, rank, world)
            sz.train_epoch_steps(net2, data, local, sizes, 0.02, dropout=0.2, seed=77, stream=epoch)
        if proto == "one-shot":
            w2 = net2.weights()
        elif proto == "two-shot":
            w4 = net2.weights()
        elif proto == "ll":
            w6 = net2.weights()
        else:
            w3 = net2.weights()
    # large batches (264 rows per rank): the update kernel of a step also prepares the next step's batch with extra CTAs that
    # take no part in the exchange (SZB_STEP_FUSE bit 2); at world 4 / 8 the 1000 windows make a single, smaller step
    ctx.comm_peer_exchange(True, "auto")
    net5 = sz.SimpleNeuralNet.from_weights(*[d[f"p{i}"] for i in range(6)], ctx=ctx)
    for epoch in range(2):
        local, sizes = shard_batches(d[f"perm{epoch}"], 264 * world, rank, world)
        sz.train_epoch_steps(net5, data, local, sizes, 0.02, dropout=0.2, seed=77, stream=epoch)
    w5 = net5.weights()
    ctx.comm_peer_exchange(False)
    # extraction: each rank takes its clip range, no collective
    clips = [d[f"clip{i}"] for i in range(6)]
    lo, hi = shard_clips([len(c) for c in clips], world)[rank]
    feats = sz.FeatureExtractor(ctx).extract_batch(clips[lo:hi]) if hi > lo else []
    np.savez(os.path.join(tmp, f"out{rank}.npz"), *net.weights(), peer=peer, **{f"q{i}": w for i, w in enumerate(w2)},
             **{f"r{i}": w for i, w in enumerate(w3)}, **{f"s{i}": w for i, w in enumerate(w4)},
             **{f"t{i}": w for i, w in enumerate(w5)}, **{f"u{i}": w for i, w in enumerate(w6)}, loss=tot, used=cnt, lo=lo, hi=hi, **{f"f{lo + i}": f for i, f in enumerate(feats)})
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_training_and_extraction_match_single_gpu(sz, ctx, oracle, tmp_path, world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    r = np.random.default_rng(9)
    n = 1000
    feats = r.standard_normal((n, 60)).astype(np.float32)
    labels = r.integers(0, 7, n).astype(np.uint32)
    onet = oracle.Net.init(60, 512, 256, 7, seed=9)
    perms = [r.permutation(n).astype(np.uint32) for _ in range(2)]
    clips = [oracle.synth_clip(i % 3, 300 + i, 0.5 + 0.3 * i) for i in range(6)]
    np.savez(str(tmp_path / "data.npz"), feats=feats, labels=labels, perm0=perms[0], perm1=perms[1],
             **{f"p{i}": p for i, p in enumerate(onet.params())}, **{f"clip{i}": c for i, c in enumerate(clips)})
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    # single-GPU run of the same epochs
    net = sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx)
    data = sz.DeviceFeatures(ctx, feats, labels)
    tot, cnt = 0.0, 0
    for epoch in range(2):
        loss, used = sz.train_epoch(net, data, perms[epoch], 96, 0.02, dropout=0.2, seed=77, stream=epoch)
        tot += loss; cnt += used
    net32 = sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx).set_precision("fp32")
    for epoch in range(2):
        sz.train_epoch(net32, data, perms[epoch], 96, 0.02, dropout=0.2, seed=77, stream=epoch)
    netb = sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx)
    for epoch in range(2):
        sz.train_epoch(netb, data, perms[epoch], 264 * world, 0.02, dropout=0.2, seed=77, stream=epoch)
    outs = [np.load(str(tmp_path / f"out{k}.npz")) for k in range(world)]
    for k in range(world):
        assert int(outs[k]["used"]) == cnt and abs(float(outs[k]["loss"]) - tot) <= 1e-3 * abs(tot)
        for i, w in enumerate(net.weights()):
            assert np.abs(outs[k][f"arr_{i}"] - w).max() <= 1e-5          # both replicas == the single-GPU result
            assert bool(outs[k]["peer"])                                  # B200 boxes have NVLink peer access
            assert np.abs(outs[k][f"q{i}"] - w).max() <= 1e-5             # ... with either gradient exchange
        for i, w in enumerate(netb.weights()):
            assert np.abs(outs[k][f"t{i}"] - w).max() <= 1e-5             # large batches: next batch prepared inside the update kernel
        for i, w in enumerate(net32.weights()):
            assert np.abs(outs[k][f"r{i}"] - w).max() <= 1e-5             # FP32 path: sgd_p2p_kernel
    for k in range(1, world):
        for i in range(6):                                                # every slice is summed once, in rank order, and
            assert np.array_equal(outs[0][f"q{i}"], outs[k][f"q{i}"])     # broadcast (two-shot) or summed by every rank in the same
            assert np.array_equal(outs[0][f"r{i}"], outs[k][f"r{i}"])     # order (one-shot): bit-identical replicas
            assert np.array_equal(outs[0][f"s{i}"], outs[k][f"s{i}"])
            assert np.array_equal(outs[0][f"t{i}"], outs[k][f"t{i}"])
            assert np.array_equal(outs[0][f"u{i}"], outs[k][f"u{i}"])
    for i in range(6):
        assert np.array_equal(outs[0][f"q{i}"], outs[0][f"s{i}"])         # the protocols add in the same order: same bits
        assert np.array_equal(outs[0][f"q{i}"], outs[0][f"u{i}"])         # ... also with the flag inside the packets (no flag round)
    single = sz.FeatureExtractor(ctx).extract_batch(clips)
    seen = 0
    for k in range(world):
        for i in range(int(outs[k]["lo"]), int(outs[k]["hi"])):
            assert np.array_equal(outs[k][f"f{i}"], single[i]); seen += 1
    assert seen == 6
