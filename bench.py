#!/usr/bin/env python
"""Benchmark of the StreamZ hot path on B200 (contract: see the task statement; one JSON line on stdout from rank 0).

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): feature extraction of 10 000 synthetic
10-s clips at 16 kHz, resampled to 44.1 kHz and turned into normalised MFCC+delta windows ([n, 60] f32).  One "step" is
one pass over the whole batch.  With N GPUs every rank processes its own 10 000 clips (weak scaling; clips are
independent, no collective on the data path).  `value` times the device-resident pass (inputs already in HBM) with CUDA
events on the library's stream; `e2e` times the host-buffer C-ABI call szb_extract_batch with pinned host memory, the
H2D copy of the PCM and the D2H copy of the features inside the timed region.  The second half of the metric (MLP train
windows/s, configs[2]: 1 M cached windows, 100 speakers, batch 4096) is reported under "mlp".

  --impl reference   times the reference's CPU path (the C restatement in oracle/, all host threads) on a bounded sample
                     of the same workload; the Rust reference itself cannot be built in this image.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATE = 16000
CLIP_SECONDS = 10
N_CLIPS = 10000
N_SPEAKERS = 100
BYTES_PER_AUDIO_S_FUSED = 2 * RATE + 26460          # 2 r + 110.25 * 240
BYTES_PER_WINDOW_44K = 1040                         # 800 B read + 240 B written
MLP_WINDOWS, MLP_SPEAKERS, MLP_BATCH = 1_000_000, 100, 4096


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel, windows_per_launch):
    """DRAM bytes (read + write) per launch of `kernel` from the committed `ncu --set full` capture, or (None, why)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))[kernel]
        per_win = (d["dram_bytes_read"] + d["dram_bytes_write"]) / d["windows_per_launch"]
        how = f"ncu dram__bytes_read.sum + dram__bytes_write.sum, {d['source']}"
        if d["windows_per_launch"] != windows_per_launch:
            how += f"; scaled from a {d['windows_per_launch']}-window launch"
        return int(per_win * windows_per_launch), how
    except Exception as e:
        return None, f"no capture ({e!r})"


# ---- synthetic multi-speaker clips, generated on the device (plumbing, outside every timed region) --------------------
def synth_clips_device(torch, device, n_clips, n_samples, rate, seed):
    """[n_clips, n_samples] int16: harmonic stack with speaker-specific pitch and formants, syllable envelope, -30 dBFS
    noise, peak -6 dBFS (same recipe as oracle.synth_clip, vectorised)."""
    g = torch.Generator(device=device)
    g.manual_seed(0x5A17 ^ seed)
    out = torch.empty((n_clips, n_samples), dtype=torch.int16, device=device)
    t = torch.arange(n_samples, device=device, dtype=torch.float32) / rate
    chunk = 250
    for c0 in range(0, n_clips, chunk):
        nc = min(chunk, n_clips - c0)
        ids = torch.arange(c0, c0 + nc, device=device)
        spk = (ids % N_SPEAKERS).to(torch.float32)
        f0 = 85.0 + torch.remainder(37.0 * spk, 170.0)
        fm = torch.stack([500.0 + 37.0 * torch.remainder(spk * 7, 11), 1500.0 + 61.0 * torch.remainder(spk * 5, 13),
                          2500.0 + 83.0 * torch.remainder(spk * 3, 7)], dim=1)
        bw = torch.tensor([120.0, 160.0, 200.0], device=device)
        x = torch.zeros((nc, n_samples), device=device)
        for h in range(1, 21):
            f = f0 * h
            gain = (1.0 / (1.0 + ((f[:, None] - fm) / bw) ** 2)).sum(dim=1) + 0.02
            gain = torch.where(f < 0.45 * rate, gain, torch.zeros_like(gain))
            ph = torch.rand((nc, 1), generator=g, device=device) * 6.2831853
            x += gain[:, None] * torch.sin(6.2831853 * f[:, None] * t[None, :] + ph)
        env_f = 4.0 + torch.remainder(ids, 5).to(torch.float32) * 0.5
        x *= 0.55 + 0.45 * torch.sin(6.2831853 * env_f[:, None] * t[None, :] + torch.rand((nc, 1), generator=g, device=device) * 6.28)
        x = x / x.abs().amax(dim=1, keepdim=True).clamp_min(1e-9) * 0.5
        x += torch.randn((nc, n_samples), generator=g, device=device) * (10 ** (-30 / 20))
        x = x / x.abs().amax(dim=1, keepdim=True).clamp_min(1e-9) * 0.5
        out[c0:c0 + nc] = torch.round(x * 32767.0).to(torch.int16)
    return out


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- reference arm / cpu baseline: the oracle's C restatement on the host cores ------------------------------------------
CPU_MODES = {0: "scalar FFT-800 (as oracle.c was timed in round 1)",
             1: "FFT-800 vectorised 8 frames wide (AVX2, rustfft-class), mel / DCT per frame in the reference's sequential sum order",
             2: "FFT, power and mel batched 8 frames wide -- stronger than the reference's own loops"}


def cpu_reference_rate(n_clips_sample, threads, steps=1, warmup=0, seed=1, mode=1):
    """audio-seconds/s of the restated reference CPU path (resample 16k->44.1k + extract) on `n_clips_sample` clips.
    mode: see CPU_MODES (oracle.c so_set_mode); 1 is the arm the speed-up is quoted against."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import streamz_oracle as orc  # bench's cpu_baseline / reference leg is allowed to run the oracle
    so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(so)
    lib.so_set_mode(C.c_int(mode))
    n = RATE * CLIP_SECONDS
    base = [orc.synth_clip(s, 1000 + s, CLIP_SECONDS, rate=RATE) for s in range(min(8, n_clips_sample))]
    pcm = np.concatenate([base[i % len(base)] for i in range(n_clips_sample)])
    off = (np.arange(n_clips_sample + 1, dtype=np.uint64) * n)
    wins = orc.n_windows(orc.resample_out_len(n, RATE))
    woff = (np.arange(n_clips_sample + 1, dtype=np.uint64) * wins)
    out = np.empty((n_clips_sample * wins, 60), np.float32)
    mel, dct, taps = orc.mel_filterbank(), orc.dct2_matrix(dtype=np.float32), orc.resample_taps(RATE)
    L, M = orc.resample_ratio(RATE)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    def run():
        lib.so_extract_batch(P(pcm), P(off), P(woff), C.c_uint32(n_clips_sample), P(mel), P(dct), C.c_uint32(RATE), P(taps),
                             C.c_uint32(L), C.c_uint32(M), C.c_uint32(16), P(out), C.c_int(threads))
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    return n_clips_sample * CLIP_SECONDS / dt, dt


def cpu_mlp_rate(n_windows, speakers, batch, seed=3):
    """train windows/s of the restated reference training loop (oracle.c so_train_epoch: per-window forward for the loss, then
    train_batch with per-sample outer-product gradients, lib.rs:599-622, 1002-1060) on ONE host thread -- the reference trains
    under a write lock (main.rs:803, lib.rs:710), so one thread is what it uses."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(so)
    lib.so_train_epoch.restype = C.c_size_t

    class SoNet(C.Structure):   # oracle.c so_net
        _fields_ = [("n_in", C.c_int), ("h1", C.c_int), ("h2", C.c_int), ("n_out", C.c_int)] + [(k, C.c_void_p) for k in
                                                                                               ("w1", "b1", "w2", "b2", "w3", "b3")]
    r = np.random.default_rng(seed)
    dims = (60, 512, 256, speakers)
    arrs = [r.uniform(-.5, .5, (dims[0], dims[1])), np.zeros(dims[1]), r.uniform(-.5, .5, (dims[1], dims[2])), np.zeros(dims[2]),
            r.uniform(-.5, .5, (dims[2], dims[3])), np.zeros(dims[3])]
    arrs = [np.ascontiguousarray(a, np.float32) for a in arrs]
    net = SoNet(*dims, *[a.ctypes.data for a in arrs])
    feats = r.standard_normal((n_windows, 60)).astype(np.float32)
    labels = r.integers(0, speakers, n_windows).astype(np.uint32)
    perm = r.permutation(n_windows).astype(np.uint32)
    keep = (r.random((n_windows, 60)) >= 0.2).astype(np.uint8)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    loss = C.c_double()
    t0 = time.perf_counter()
    used = lib.so_train_epoch(C.byref(net), P(feats), P(labels), P(perm), C.c_size_t(n_windows), C.c_size_t(batch), C.c_float(0.01),
                              P(keep), C.byref(loss))
    dt = time.perf_counter() - t0
    return n_windows / dt, dt, int(used)


def cpu_sample_for(seconds, threads, cap, mode=1):
    """Number of clips whose CPU extraction takes about `seconds` on this host (probe with a few clips per thread first)."""
    probe = int(min(cap, max(16, threads * 4)))
    rate, _ = cpu_reference_rate(probe, threads, mode=mode)
    return int(min(cap, max(probe, rate * seconds / CLIP_SECONDS)))


def cpu_baseline_block(cores, seconds, cap):
    """cpu_baseline of the extraction metric: the arm the ratio is quoted against (mode 1) plus the scalar and the fully
    batched variants on short samples, each with its per-core cost, so the "x" beside it can be read for what it is."""
    wins_per_clip = 1101
    sample = cpu_sample_for(seconds, cores, cap, mode=1)
    rate, dt = cpu_reference_rate(sample, cores, steps=1, warmup=0, mode=1)
    out = {"value": rate, "unit": "audio-s/s", "cores": cores, "kind": "port",
           "us_per_window_per_core": dt * cores / (sample * wins_per_clip) * 1e6,
           "sample": f"{sample} of {cap} clips, C restatement of lib.rs:186-345 (oracle/oracle.c), one clip per thread, {dt:.2f} s; "
                     + CPU_MODES[1], "variants": {}}
    for mode in (0, 2):
        try:
            smp = cpu_sample_for(min(3.0, seconds), cores, cap, mode=mode)
            r, d = cpu_reference_rate(smp, cores, steps=1, warmup=0, mode=mode)
            out["variants"]["scalar" if mode == 0 else "batched"] = {
                "value": r, "us_per_window_per_core": d * cores / (smp * wins_per_clip) * 1e6, "sample": f"{smp} clips, {d:.2f} s; " + CPU_MODES[mode]}
        except Exception as e:
            out["variants"][str(mode)] = {"error": repr(e)}
    return out


def pin_to_gpu_numa_node(torch, local):
    """Multi-rank runs share one host: bind this rank's threads to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI
    function) BEFORE the pinned staging buffers are allocated, so they are first-touched on the GPU's own NUMA node and the
    host-buffer e2e path of every rank copies through its local memory controller and PCIe root.  Returns a short description
    for the JSON line, or None when the topology cannot be read."""
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        dev_id = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0"
        cpus = open(os.path.join(path, "local_cpulist")).read().strip()
        node = open(os.path.join(path, "numa_node")).read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return {"numa_node": int(node), "cpus": cpus, "bound": bool(ids)}
    except Exception as e:
        return {"error": repr(e)}


def run_mlp(torch, dist, sz, N, ctx, dev, rank, world, feats, total):
    """configs[2]: one epoch over 1 M cached windows per GPU, 100 speakers, batch 4096 per GPU, lr 0.01, dropout 0.2.
    At N > 1 the epoch is timed with BOTH gradient exchanges (overlapped NCCL all-reduces; the two-shot peer-memory exchange
    fused into the update kernel), each on a fresh net from the same seed; the faster one is the reported number and both
    timings stay in the line.  Device time (CUDA events on the library's stream), max over ranks."""
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    nwin = MLP_WINDOWS
    src = feats[:nwin] if total >= nwin else torch.randn((nwin, 60), generator=g, device=dev)
    src = src.contiguous()
    labels = torch.randint(0, MLP_SPEAKERS, (nwin,), generator=g, device=dev, dtype=torch.int32)
    if world > 1:
        uid = [sz.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
    perm = np.random.default_rng(5).permutation(nwin).astype(np.uint32)
    loss, used = C.c_double(), C.c_uint64()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_epoch(exchange):
        if world > 1:
            active = ctx.comm_peer_exchange(exchange != "nccl", exchange if exchange != "nccl" else "auto")
            if exchange != "nccl" and not active:
                return None
        net = sz.SimpleNeuralNet(60, 512, 256, MLP_SPEAKERS, seed=7, ctx=ctx)   # default arithmetic: 3xTF32 on tcgen05
        def epoch(n_rows):
            N.check(N.lib.szb_net_train_epoch_dev(net._h, C.c_void_p(src.data_ptr()), C.c_void_p(labels.data_ptr()), nwin, N.ptr(perm),
                                                  n_rows, MLP_BATCH, 0.01, 0.2, 99, 0, None, C.byref(loss), C.byref(used)))
        epoch(nwin)                 # warm-up: one epoch of the timed size (the library captures its two-step graph and sizes its
                                    # permutation buffer on the first epoch of a size; a training run has 60-100 epochs, main.rs:36)
        barrier()
        launches0 = ctx.launch_count
        ctx.timer_start()
        epoch(nwin)
        ms = ctx.timer_stop()
        launches = ctx.launch_count - launches0
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        # replicas must hold identical bits after the epoch: compare a checksum of the raw weight bytes across ranks
        w = np.concatenate([a.ravel() for a in net.weights()])
        digest = [int(np.bitwise_xor.reduce(w.view(np.uint32).astype(np.uint64) * (np.arange(w.size, dtype=np.uint64) | np.uint64(1))))]
        same = True
        if world > 1:
            digests = [None] * world
            dist.all_gather_object(digests, digest[0])
            same = all(d == digests[0] for d in digests)
        out = {"ms_per_epoch": float(tt.item()), "us_per_step": float(tt.item()) * 1e3 / ((nwin + MLP_BATCH - 1) // MLP_BATCH),
               "mean_loss": loss.value / max(1, used.value), "replicas_identical": bool(same), "launches": int(launches)}
        net.close()
        return out

    runs = {}
    if world == 1:
        runs["none (1 GPU)"] = timed_epoch("none")
    else:
        runs["NCCL all-reduce per layer, overlapped"] = timed_epoch("nccl")
        for proto in ("two-shot", "one-shot", "ll"):
            r = timed_epoch(proto)
            if r is not None:
                name = "packet (flag inside every 8-byte store, no flag round)" if proto == "ll" else proto
                runs[f"{name} peer-memory exchange fused into the update kernel"] = r
        ctx.comm_peer_exchange(False)
    best = min(runs, key=lambda k: runs[k]["ms_per_epoch"])
    ms_epoch = runs[best]["ms_per_epoch"]
    steps = (nwin + MLP_BATCH - 1) // MLP_BATCH
    flop_per_win = 909312 + 1536 * MLP_SPEAKERS
    tflops = nwin * flop_per_win / (ms_epoch * 1e-3) / 1e12            # per GPU, algorithmic (each product counted once)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf32_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
    mlp = {"train_windows_per_s": world * nwin / (ms_epoch * 1e-3), "ms_per_epoch": ms_epoch, "us_per_step": ms_epoch * 1e3 / steps,
           "windows": nwin, "speakers": MLP_SPEAKERS, "batch_per_gpu": MLP_BATCH, "mean_loss": runs[best]["mean_loss"],
           "tflops": world * tflops, "gpu_launches": runs[best]["launches"],
           "precision": "3xTF32 (tcgen05 kind::tf32, split hi/lo, FP32-equivalent); tflops counts algorithmic FLOPs once",
           "kernels": "gemm_tma_kernel: TMA operand fetch, A operand and accumulator in tensor memory; weight gradients as one grouped launch",
           "grad_exchange": best, "exchange_timings": runs, "replicas_identical": all(r["replicas_identical"] for r in runs.values()),
           "roofline": {"bound": "tensor", "achieved": tflops, "peak": tf32_peak, "unit": "TFLOP/s", "frac": tflops / tf32_peak,
                        "executed": 3 * tflops, "traffic": None,
                        "peak_source": ("measured" if peaks else "fallback") + " dense bf16 burst / 2 (kind::tf32 runs at half the bf16 rate)",
                        "note": "per GPU; achieved = algorithmic FLOPs of a training step (1 062 912 per window: forward, dX and dW of "
                                "every layer, each product once) / step time; the 3xTF32 split executes 3x that on the tensor pipe "
                                "(`executed`); the step is latency-bound: %d dependent launches of 4096-row GEMMs per step (DESIGN.md 6)"
                                % round(runs[best]["launches"] / max(1, steps))},
           "workload": "configs[2]: 1M cached windows, 100 speakers, batch 4096 per GPU, 1 epoch, lr 0.01, dropout 0.2"}
    # end to end through the host API: features and labels start in pinned host memory, are uploaded inside the timed region
    # (szb_memcpy_h2d), the epoch runs, the loss comes back (szb_net_train_epoch_dev returns it on the host)
    try:
        h_src = torch.empty((nwin, 60), dtype=torch.float32, pin_memory=True); h_src.copy_(src)
        h_lab = torch.empty((nwin,), dtype=torch.int32, pin_memory=True); h_lab.copy_(labels)
        d_src = torch.empty_like(src); d_lab = torch.empty_like(labels)
        net = sz.SimpleNeuralNet(60, 512, 256, MLP_SPEAKERS, seed=7, ctx=ctx)
        if world > 1:
            if best.startswith("NCCL"):
                ctx.comm_peer_exchange(False)
            else:
                ctx.comm_peer_exchange(True, "ll" if best.startswith("packet") else best.split(" ")[0])
        def e2e_epoch(n_rows):
            N.check(N.lib.szb_memcpy_h2d(ctx.handle, C.c_void_p(d_src.data_ptr()), C.c_void_p(h_src.data_ptr()), nwin * 240))
            N.check(N.lib.szb_memcpy_h2d(ctx.handle, C.c_void_p(d_lab.data_ptr()), C.c_void_p(h_lab.data_ptr()), nwin * 4))
            N.check(N.lib.szb_net_train_epoch_dev(net._h, C.c_void_p(d_src.data_ptr()), C.c_void_p(d_lab.data_ptr()), nwin, N.ptr(perm),
                                                  n_rows, MLP_BATCH, 0.01, 0.2, 99, 0, None, C.byref(loss), C.byref(used)))
        e2e_epoch(nwin)
        barrier()
        t0 = time.perf_counter()
        e2e_epoch(nwin)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ctx.comm_peer_exchange(False)
        mlp["e2e"] = {"value": world * nwin / float(tt.item()), "unit": "train windows/s", "h2d_bytes_per_step": int(nwin * 244),
                      "d2h_bytes_per_step": 16, "note": "step = one epoch: upload of the 1 M windows + labels from pinned host memory, "
                      "245 training steps, loss and count read back; wall clock, max over ranks"}
        net.close()
    except Exception as e:
        mlp["e2e"] = {"error": repr(e)}
    return mlp


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # bounded sample: about 3 s of host work per step keeps K + W steps within a couple of minutes on any host
    sample = cpu_sample_for(3.0, cores, N_CLIPS)
    rate, dt = cpu_reference_rate(sample, cores, steps=args.steps, warmup=args.warmup, mode=1)
    line = {"impl": "reference", "metric": "audio-seconds/sec MFCC+delta extraction", "value": rate, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"extract {N_CLIPS} x {CLIP_SECONDS}-s clips 16k->44.1k (BASELINE configs[1]); each step = "
                                   f"{sample}-clip sample of it on the host CPU", "clips_per_step": sample},
            "cpu_baseline": {"value": rate, "unit": "audio-s/s", "cores": cores, "kind": "port",
                             "us_per_window_per_core": dt * cores / (sample * 1101) * 1e6,
                             "sample": f"{sample} of {N_CLIPS} clips per step, C restatement of lib.rs:186-345, one clip per thread; " + CPU_MODES[1]},
            "e2e": {"value": rate, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print_line(line)
    return 0


def main():
    # Libraries print to stdout (NCCL's version banner does, on every rank): the contract is ONE JSON line there, so
    # everything written to fd 1 during the run goes to stderr and the JSON line is written to the saved descriptor.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    global print_line
    def print_line(obj):
        os.write(saved, (json.dumps(obj) + "\n").encode())
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=N_CLIPS, help="clips per GPU (default: the full configs[1] batch)")
    ap.add_argument("--no-mlp", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    ap.add_argument("--config", default="c2", choices=["c2", "c1", "c4", "c5"],
                    help="c2 (default): BASELINE configs[1] + configs[2], the driver's line; c1 / c4 / c5: the other BASELINE configs "
                         "(configs[0], [3], [4]), each printing its own JSON line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config != "c2":
        import bench_configs
        return bench_configs.run(args, print_line)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: streamz_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(torch, local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import streamz_b200 as sz
    from streamz_b200 import _native as N

    stream = torch.cuda.Stream(device=dev)
    ctx = sz.Context(local, stream=stream.cuda_stream)
    n_clips, n_in = args.clips, RATE * CLIP_SECONDS
    audio_s = n_clips * CLIP_SECONDS

    pcm = synth_clips_device(torch, dev, n_clips, n_in, RATE, seed=rank)
    off = (np.arange(n_clips + 1, dtype=np.uint64) * n_in)
    total = int(N.lib.szb_extract_batch_windows(N.ptr(off), n_clips, RATE))
    feats = torch.empty((total, 60), dtype=torch.float32, device=dev)
    woff = np.zeros(n_clips + 1, np.uint64)
    torch.cuda.synchronize()

    def step_dev():
        N.check(N.lib.szb_extract_batch_dev(ctx.handle, C.c_void_p(pcm.data_ptr()), N.ptr(off), n_clips, RATE,
                                            C.c_void_p(feats.data_ptr()), total, N.ptr(woff)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        barrier()
        launches0 = ctx.launch_count
        ctx.timer_start()
        for _ in range(steps):
            fn()
        ms = ctx.timer_stop()
        barrier()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ctx.launch_count - launches0

    # ---- device-resident value + per-launch kernel timing for the roofline ------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.kernel_timing(True)
    for _ in range(args.warmup):
        step_dev()
    ctx.sync()
    ctx.kernel_timing_read(reset=True)
    ms_total, launches = timed(step_dev, args.steps, 0)
    k_ms, k_n = ctx.kernel_timing_read(reset=True)
    ctx.kernel_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * audio_s / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer C ABI ---------------------------------------------------------------------------
    h_pcm = torch.empty((n_clips * n_in,), dtype=torch.int16, pin_memory=True)
    h_pcm.copy_(pcm.reshape(-1))
    h_feats = torch.empty((total, 60), dtype=torch.float32, pin_memory=True)
    torch.cuda.synchronize()

    def step_e2e():
        N.check(N.lib.szb_extract_batch(ctx.handle, C.c_void_p(h_pcm.data_ptr()), N.ptr(off), n_clips, RATE,
                                        C.c_void_p(h_feats.data_ptr()), total, N.ptr(woff)))

    e2e_steps = args.e2e_steps or args.steps
    for _ in range(max(1, args.warmup - 1)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()                      # returns after the D2H copy of the features has completed
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * audio_s / float(t.item())
    checksum = float(h_feats[:: max(1, total // 997)].double().abs().sum().item())

    # ---- MLP training throughput (configs[2]) ----------------------------------------------------------------------------------
    mlp = None
    if not args.no_mlp:
        try:
            mlp = run_mlp(torch, dist, sz, N, ctx, dev, rank, world, feats, total)
        except Exception as e:  # the headline metric must still be reported
            mlp = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peaks()
    launches_per_step = max(1, int(round(k_n / max(1, args.steps))))
    windows_per_launch = total // launches_per_step
    algo_bytes = windows_per_launch * BYTES_PER_WINDOW_44K
    k_avg_ms = k_ms / max(1, k_n)
    achieved = algo_bytes / (k_avg_ms * 1e-3) / 1e9 if k_n else None
    traffic, traffic_src = measured_traffic("extract_kernel", windows_per_launch)
    # the whole step (resample + extract) against SURVEY.md 8(d)'s fused figure: 2 r + 26 460 bytes per audio-second
    step_algo = BYTES_PER_AUDIO_S_FUSED * audio_s
    step_achieved = step_algo / (ms_per_step * 1e-3) / 1e9
    rs_traffic, _ = measured_traffic("resample_kernel", total)
    step_traffic = (traffic * launches_per_step + rs_traffic) if (traffic and rs_traffic) else None
    line = {
        "metric": "audio-seconds/sec MFCC+delta extraction", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"extract {n_clips} x {CLIP_SECONDS}-s clips 16k->44.1k per GPU (BASELINE configs[1])",
                   "clips_per_gpu": n_clips, "windows_per_gpu": total, "audio_seconds_per_step": world * audio_s,
                   "l2": "inputs (3.2 GB) and outputs (2.6 GB) per step exceed the 126 MB L2; no flush needed"},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(n_clips * n_in * 2), "d2h_bytes_per_step": int(total * 240),
                "steps": e2e_steps, "checksum": checksum, "per_gpu": e2e_value / world, "host_binding_rank0": numa},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_src, "kernel": "extract_kernel", "launch_ms": k_avg_ms, "launches_timed": int(k_n),
                     "algorithmic_bytes_per_launch": int(algo_bytes), "peak_source": peak_src,
                     "note": "1040 B/window (800 read + 240 written) x windows per launch / CUDA-event launch time",
                     "step": {"achieved": step_achieved, "frac": step_achieved / peak, "algorithmic_bytes_per_step": int(step_algo),
                              "traffic": step_traffic, "ms": ms_per_step,
                              "note": "whole step (resample_kernel + extract_kernel) against the fused figure of SURVEY.md 8(d), "
                                      "58 460 B per audio-second at 16 kHz; traffic = measured DRAM bytes of both kernels: the 44.1 kHz "
                                      "i16 intermediate makes one round trip through HBM (DESIGN.md 5: fusing it away was measured "
                                      "slower, neither kernel is HBM-bound)"}},
        "clocks": clocks, "mlp": mlp,
    }
    if not args.no_cpu and world >= 1:
        cores = os.cpu_count() or 1
        try:
            line["cpu_baseline"] = cpu_baseline_block(cores, 10.0, n_clips)   # about 10 s + 2 x 3 s of CPU work
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)}
        if isinstance(mlp, dict) and "error" not in mlp and rank == 0:
            try:                                               # bounded: two batches of the configs[2] shape, one thread
                wps, dt, used = cpu_mlp_rate(2 * MLP_BATCH, MLP_SPEAKERS, MLP_BATCH)
                mlp["cpu_baseline"] = {"value": wps, "unit": "train windows/s", "cores": 1, "kind": "port",
                                       "sample": f"{2 * MLP_BATCH} windows (2 steps of batch {MLP_BATCH}), C restatement of lib.rs:599-622 + "
                                                 f"1002-1060 (oracle/oracle.c so_train_epoch), one thread as the reference trains under a "
                                                 f"write lock, {dt:.2f} s"}
            except Exception as e:
                mlp["cpu_baseline"] = {"error": repr(e)}
    print_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
