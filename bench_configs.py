"""The BASELINE.json configurations that are not the driver's bench line (bench.py --config c1 | c4 | c5), one JSON line each.

c1  configs[0], the reference's own CPU-runnable case: 4 synthetic 5-s clips @ 44.1 kHz, 2 speakers, extraction + initial
    training (train_from_feature_map, main.rs:658-666: every file for 60 epochs, batch 8, lr 0.01, dropout 0.2).  Wall time of
    the whole flow through the host API on the GPU beside the C port on one host thread (the reference trains under a
    write lock).  This is the latency-bound regime the reference actually runs: 16 560 steps of 8 windows.
c4  configs[3], extract -> train -> identify on synthetic 10-s clips @ 16 kHz sharded over the ranks (default 12 500 clips per
    rank = 100 000 at 8 ranks; 1 000 speakers): features stay on the producing GPU, one epoch of batch-parallel training with
    the fused gradient exchange, then identify_speaker_list's histogram for every clip of the shard.
c5  configs[4], identification sweep: one mixed clip of {7.3 s, 60 s, 8 min, 64 min} = {0.8 k, 6.6 k, 53 k, 424 k} windows cut
    into window ranges over the ranks (szb_extract_range_dev + szb_identify_counts_dev), per-class counts added on the host.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _dist():
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    return torch, dist, rank, local, world, dev


def _max_over_ranks(torch, dist, dev, world, v):
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import streamz_oracle as orc   # synthetic clips + the CPU baseline leg only
    return orc


def _oracle_lib():
    import subprocess
    so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(so)
    lib.so_train_epoch.restype = C.c_size_t
    lib.so_extract.restype = C.c_size_t
    return lib


class _SoNet(C.Structure):   # oracle.c so_net
    _fields_ = [("n_in", C.c_int), ("h1", C.c_int), ("h2", C.c_int), ("n_out", C.c_int)] + [(k, C.c_void_p) for k in
                                                                                           ("w1", "b1", "w2", "b2", "w3", "b3")]


def run_c1(args, print_line):
    torch, dist, rank, local, world, dev = _dist()
    import streamz_b200 as sz
    orc = _oracle()
    epochs, batch, lr, dropout = 60, 8, 0.01, 0.2
    clips = [orc.synth_clip(s % 2, 500 + s, 5.0) for s in range(4)]
    files = [(f"clip{i}.wav", i % 2) for i in range(4)]
    ctx = sz.Context(local)
    ex = sz.FeatureExtractor(ctx)
    def flow(seed):
        t0 = time.perf_counter()
        feats = ex.extract_batch(clips)                                                   # main.rs:500-508
        t1 = time.perf_counter()
        net = sz.SimpleNeuralNet(60, 512, 256, 2, seed=seed, ctx=ctx)                      # main.rs:640-649
        fmap = {p: f for (p, _), f in zip(files, feats)}
        loss = sz.train_from_feature_map(net, fmap, files, epochs, lr, dropout, batch, seed=seed)   # main.rs:658-666
        ctx.sync()
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, loss, sum(len(f) for f in feats), net
    flow(0)[4].close()                                                                    # warm-up (allocations, first launches)
    l0 = ctx.launch_count
    t_ext, t_train, loss, n_win, net = flow(1)
    launches = ctx.launch_count - l0
    net.close()
    # CPU port, one thread: extraction + the same schedule through so_train_epoch
    lib = _oracle_lib()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    mel, dct = orc.mel_filterbank(), orc.dct2_matrix(dtype=np.float32)
    c0 = time.perf_counter()
    cfeats = []
    for c in clips:
        out = np.empty((orc.n_windows(len(c)), 60), np.float32)
        lib.so_extract(P(c), C.c_size_t(len(c)), P(mel), P(dct), P(out))
        cfeats.append(out)
    c1 = time.perf_counter()
    r = np.random.default_rng(1)
    arrs = [np.ascontiguousarray(a, np.float32) for a in (r.uniform(-.5, .5, (60, 512)), np.zeros(512), r.uniform(-.5, .5, (512, 256)),
                                                           np.zeros(256), r.uniform(-.5, .5, (256, 2)), np.zeros(2))]
    cnet = _SoNet(60, 512, 256, 2, *[a.ctypes.data for a in arrs])
    closs = C.c_double()
    cpu_epochs = 4                                   # bounded sample: 4 of the 60 epochs of every file, scaled (the loop is uniform)
    for (pth, cls), f in zip(files, cfeats):
        lab = np.full(len(f), cls, np.uint32)
        for e in range(cpu_epochs):
            perm = r.permutation(len(f)).astype(np.uint32)
            keep = (r.random((len(f), 60)) >= dropout).astype(np.uint8)
            lib.so_train_epoch(C.byref(cnet), P(f), P(lab), P(perm), C.c_size_t(len(f)), C.c_size_t(batch), C.c_float(lr), P(keep), C.byref(closs))
    c2 = time.perf_counter()
    cpu_total = (c1 - c0) + (c2 - c1) * epochs / cpu_epochs
    steps = sum((len(f) + batch - 1) // batch for f in cfeats) * epochs
    line = {"metric": "wall seconds, extraction + initial training (BASELINE configs[0])", "value": t_ext + t_train, "unit": "s", "n_gpus": 1,
            "steps": 1, "warmup": 1, "ms_per_step": (t_ext + t_train) * 1e3, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c1: 4 synthetic 5-s clips @ 44.1 kHz, 2 speakers, MFCC+delta windows + SimpleNeuralNet 60 epochs per file, "
                                   "batch 8, lr 0.01, dropout 0.2 (main.rs:658-666 with BASELINE's epoch count)", "windows": n_win,
                       "optimizer_steps": steps},
            "e2e": {"value": t_ext + t_train, "unit": "s", "h2d_bytes_per_step": int(sum(len(c) for c in clips) * 2 + n_win * 240 * 1),
                    "d2h_bytes_per_step": int(n_win * 240), "note": "host API end to end: clips in host memory, features returned to the host, "
                    "then uploaded once per file for training"},
            "gpu_launches": int(launches),
            "breakdown": {"extract_s": t_ext, "train_s": t_train, "us_per_optimizer_step": t_train / steps * 1e6, "mean_loss": loss},
            "cpu_baseline": {"value": cpu_total, "unit": "s", "cores": 1, "kind": "port",
                             "sample": f"extraction of the 4 clips ({c1 - c0:.2f} s) + {cpu_epochs} of the {epochs} epochs per file "
                                       f"({c2 - c1:.2f} s, scaled x{epochs // cpu_epochs}); one thread: the reference trains under a write lock "
                                       "(main.rs:803) and extraction of 4 clips is a 4-task rayon loop"}}
    print_line(line)
    return 0


def _synth_device(torch, dev, n_clips, seed, speakers):
    import bench
    old = bench.N_SPEAKERS
    bench.N_SPEAKERS = speakers
    try:
        return bench.synth_clips_device(torch, dev, n_clips, 160000, 16000, seed=seed)
    finally:
        bench.N_SPEAKERS = old


def run_c4(args, print_line):
    torch, dist, rank, local, world, dev = _dist()
    import streamz_b200 as sz
    from streamz_b200 import _native as N
    n_clips = args.clips if args.clips != 10000 else 12500
    speakers, batch = 1000, 4096
    stream = torch.cuda.Stream(device=dev)
    ctx = sz.Context(local, stream=stream.cuda_stream)
    pcm = _synth_device(torch, dev, n_clips, rank, speakers)
    off = np.arange(n_clips + 1, dtype=np.uint64) * 160000
    total = int(N.lib.szb_extract_batch_windows(N.ptr(off), n_clips, 16000))
    feats = torch.empty((total, 60), dtype=torch.float32, device=dev)
    woff = np.zeros(n_clips + 1, np.uint64)
    per_clip = total // n_clips
    labels = ((torch.arange(n_clips, device=dev) + rank * n_clips) % speakers).to(torch.int32).repeat_interleave(per_clip).contiguous()
    net = sz.SimpleNeuralNet(60, 512, 256, speakers, seed=7, ctx=ctx)
    peer = False
    if world > 1:
        uid = [sz.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        peer = ctx.comm_peer_exchange(True)
    perm = np.random.default_rng(5 + rank).permutation(total).astype(np.uint32)
    n_steps = (total + batch - 1) // batch
    sizes = np.full(n_steps, batch, np.uint32); sizes[-1] = total - batch * (n_steps - 1)
    loss, used = C.c_double(), C.c_uint64()
    counts = np.zeros(speakers, np.uint64)
    batch_counts = np.zeros((n_clips, speakers), np.uint32)
    ident_clips = min(n_clips, 2000)
    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    def run(timed):
        t = {}
        sync(); t0 = time.perf_counter()
        N.check(N.lib.szb_extract_batch_dev(ctx.handle, C.c_void_p(pcm.data_ptr()), N.ptr(off), n_clips, 16000, C.c_void_p(feats.data_ptr()), total, N.ptr(woff)))
        ctx.sync(); t["extract_s"] = time.perf_counter() - t0
        sync(); t0 = time.perf_counter()
        ns = n_steps if timed else 8                                   # warm-up pass: eight steps
        rows = int(sizes[:ns].sum())
        N.check(N.lib.szb_net_train_epoch_steps_dev(net._h, C.c_void_p(feats.data_ptr()), C.c_void_p(labels.data_ptr()), total, N.ptr(perm), rows,
                                                    N.ptr(sizes), ns, 0.01, 0.2, 99, 0, None, C.byref(loss), C.byref(used)))
        ctx.sync(); t["train_s"] = time.perf_counter() - t0
        sync(); t0 = time.perf_counter()
        hits = 0
        for c in range(ident_clips if timed else 16):                  # identify_speaker_list's histogram, one call per clip (lib.rs:1389-1402)
            N.check(N.lib.szb_identify_counts_dev(net._h, C.c_void_p(feats.data_ptr() + int(woff[c]) * 240), per_clip, 0.0, N.ptr(counts)))
            hits += int(np.argmax(counts) == (c + rank * n_clips) % speakers)
        t["identify_s"] = time.perf_counter() - t0
        t["top1_hits"] = hits
        sync(); t0 = time.perf_counter()                               # the same histograms for the WHOLE shard in one batched call
        nb = n_clips if timed else 64
        N.check(N.lib.szb_identify_counts_batch_dev(net._h, C.c_void_p(feats.data_ptr()), N.ptr(woff), nb, 0.0, N.ptr(batch_counts)))
        t["identify_batched_s"] = time.perf_counter() - t0
        t["batched_equals_per_clip"] = bool(np.array_equal(batch_counts[min(nb, ident_clips if timed else 16) - 1].astype(np.uint64), counts))
        return t
    run(False)
    l0 = ctx.launch_count
    t = run(True)
    launches = ctx.launch_count - l0
    ext = _max_over_ranks(torch, dist, dev, world, t["extract_s"]); trn = _max_over_ranks(torch, dist, dev, world, t["train_s"])
    idn = _max_over_ranks(torch, dist, dev, world, t["identify_s"])
    idb = _max_over_ranks(torch, dist, dev, world, t["identify_batched_s"])
    if rank == 0:
        audio_s = world * n_clips * 10
        ident_per_clip_full = idn * n_clips / ident_clips
        ident_full = idb
        line = {"metric": "audio-seconds/sec, extract + 1 training epoch + identification (BASELINE configs[3])",
                "value": audio_s / (ext + trn + ident_full), "unit": "audio-s/s", "n_gpus": world, "steps": 1, "warmup": 1,
                "ms_per_step": (ext + trn + ident_full) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"c4: {n_clips} synthetic 10-s clips @ 16 kHz per GPU ({world * n_clips} in all), {speakers} speakers; "
                                       "extraction device-resident, 1 epoch batch 4096 per GPU, identification histogram per clip",
                           "windows_per_gpu": total, "train_steps": int(n_steps)},
                "gpu_launches": int(launches),
                "breakdown": {"extract_s": ext, "train_epoch_s": trn, "train_windows_per_s": world * total / trn, "us_per_train_step": trn / n_steps * 1e6,
                              "identify_batched_s": idb, "identify_clips_per_s": world * n_clips / idb,
                              "identify_per_clip_api_s_measured": idn, "identify_per_clip_api_clips_measured": ident_clips,
                              "identify_per_clip_api_s_all_clips": ident_per_clip_full, "batched_equals_per_clip": t["batched_equals_per_clip"], "grad_exchange": "two-shot peer-memory" if peer else ("NCCL" if world > 1 else "none"),
                              "mean_loss": loss.value / max(1, used.value), "top1_hits_rank0": t["top1_hits"]},
                "e2e": None, "cpu_baseline": None,
                "note": "clips and features stay on the GPUs (no host copy in this flow); identification = one batched call over the whole "
                        "shard (szb_identify_counts_batch_dev, per-clip histograms read back to the host); the per-clip API is timed beside it "
                        f"on the first {ident_clips} clips of every shard and scaled; the CPU arm of this flow is the sum of the c2 line's "
                        "extraction and MLP baselines"}
        print_line(line)
    if world > 1:
        ctx.comm_peer_exchange(False)
        dist.destroy_process_group()
    return 0


def run_c5(args, print_line):
    torch, dist, rank, local, world, dev = _dist()
    import streamz_b200 as sz
    from streamz_b200 import _native as N
    from streamz_b200.sharding import shard_windows
    orc = _oracle()
    ctx = sz.Context(local)
    speakers = 5
    onet = orc.Net.init(60, 512, 256, speakers, seed=31)
    net = sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx)
    unit = np.concatenate([orc.synth_clip(s, 60 + s, 12.0) for s in range(5)])           # 60 s, five speakers
    sweep = []
    for target in (800, 6614, 53000, 424000):
        n_samples = 800 + 400 * (target - 1)
        clip = np.tile(unit, n_samples // len(unit) + 1)[:n_samples]
        n_win = int(N.lib.szb_num_windows(n_samples))
        w0, w1 = shard_windows(n_win, world)[rank]
        d_feats = ctx.dev_alloc(max(1, (w1 - w0) * 240))
        counts = np.zeros(speakers, np.uint64)
        def once():
            N.check(N.lib.szb_extract_range_dev(ctx.handle, N.ptr(clip), n_samples, w0, w1, C.c_void_p(d_feats), w1 - w0))
            N.check(N.lib.szb_identify_counts_dev(net._h, C.c_void_p(d_feats), w1 - w0, 0.5, N.ptr(counts)))
            t = torch.from_numpy(counts.astype(np.int64)).to(dev)
            if world > 1:
                dist.all_reduce(t)                          # plumbing for the host-side sum of <= C integers
            return t.cpu().numpy()
        once()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            total_counts = once()
        dt = _max_over_ranks(torch, dist, dev, world, (time.perf_counter() - t0) / reps)
        ctx.dev_free(d_feats)
        entry = {"windows": n_win, "audio_s": n_samples / 44100.0, "ms": dt * 1e3, "windows_per_s": n_win / dt, "counts": total_counts.tolist()}
        if rank == 0 and n_win <= 7000:                      # the CPU port on the same clip, one thread (lib.rs:1383-1411 is serial per clip)
            lib = _oracle_lib()
            P = lambda a: a.ctypes.data_as(C.c_void_p)
            mel, dct = orc.mel_filterbank(), orc.dct2_matrix(dtype=np.float32)
            out = np.empty((n_win, 60), np.float32)
            cn = _SoNet(60, 512, 256, speakers, *[a.ctypes.data for a in onet.params()])
            cc = np.zeros(speakers, np.uint64)
            c0 = time.perf_counter()
            lib.so_extract(P(clip), C.c_size_t(n_samples), P(mel), P(dct), P(out))
            lib.so_identify_counts(C.byref(cn), P(out), C.c_size_t(n_win), C.c_float(0.5), P(cc))
            entry["cpu_ms_one_thread"] = (time.perf_counter() - c0) * 1e3
            entry["cpu_counts_equal"] = bool(np.array_equal(cc.astype(np.int64), total_counts))
        sweep.append(entry)
    if rank == 0:
        big = sweep[-1]
        line = {"metric": "identification windows/sec on one long clip sharded by window range (BASELINE configs[4])", "value": big["windows_per_s"],
                "unit": "windows/s", "n_gpus": world, "steps": 5, "warmup": 1, "ms_per_step": big["ms"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "c5: mixed-speaker clips of 0.8 k / 6.6 k / 53 k / 424 k windows; every rank extracts and classifies its "
                                       "window range (+2-frame halo), per-class counts summed across ranks; value = the 424 k-window point"},
                "sweep": sweep,
                "e2e": {"value": big["windows_per_s"], "unit": "windows/s", "h2d_bytes_per_step": int(big["windows"] * 800 / world), "d2h_bytes_per_step": speakers * 8,
                        "note": "already end to end: the clip is in host memory, each rank uploads only its range, the counts come back to the host"},
                "cpu_baseline": {"value": (sweep[1]["windows"] / (sweep[1].get("cpu_ms_one_thread", float("nan")) * 1e-3)), "unit": "windows/s", "cores": 1,
                                 "kind": "port", "sample": "the 60-s clip (6 614 windows): so_extract + so_identify_counts on one thread (identify_speaker_list is "
                                                           "serial per clip in the reference)"}}
        print_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run(args, print_line):
    return {"c1": run_c1, "c4": run_c4, "c5": run_c5}[args.config](args, print_line)
