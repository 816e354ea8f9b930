"""Device-resident configs[1] pass (10 000 x 10-s clips, 16 kHz -> 44.1 kHz -> features) under different L2-ring settings:
   python tools/gpu_frontend_sweep.py [chunk_mb:streams ...]   e.g. 0:1 32:1 32:2"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import streamz_b200 as sz
from streamz_b200 import _native as N
import bench
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = sz.Context(0, stream=stream.cuda_stream)
n_clips = int(os.environ.get("CLIPS", "10000"))
n_in = 160000
pcm = bench.synth_clips_device(torch, dev, n_clips, n_in, 16000, seed=0)
off = (np.arange(n_clips + 1, dtype=np.uint64) * n_in)
total = int(N.lib.szb_extract_batch_windows(N.ptr(off), n_clips, 16000))
feats = torch.empty((total, 60), dtype=torch.float32, device=dev)
woff = np.zeros(n_clips + 1, np.uint64)
torch.cuda.synchronize()
def step():
    N.check(N.lib.szb_extract_batch_dev(ctx.handle, C.c_void_p(pcm.data_ptr()), N.ptr(off), n_clips, 16000, C.c_void_p(feats.data_ptr()), total, N.ptr(woff)))
ref = None
# "fused" = FIR inside the extraction kernel (no intermediate); "mb:streams" = two-kernel path, L2 ring of mb MB (0 = one chunk)
cfgs = [("fused",) if a == "fused" else tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(0, 1), ("fused",)]
steps = int(os.environ.get("STEPS", "5"))
for cfg in cfgs:
    fused = cfg[0] == "fused"
    mb, st = (0, 1) if fused else cfg
    N.check(N.lib.szb_ctx_set_fused_resample(ctx.handle, 1 if fused else 0))
    N.check(N.lib.szb_ctx_set_l2_ring(ctx.handle, mb, st))
    for _ in range(2):
        step()
    ctx.sync()
    ctx.kernel_timing(True); ctx.kernel_timing_read(reset=True)
    ctx.timer_start()
    for _ in range(steps):
        step()
    ms = ctx.timer_stop() / steps
    k_ms, k_n = ctx.kernel_timing_read(reset=True); ctx.kernel_timing(False)
    cs = float(feats[::997].double().sum().item())
    if ref is None:
        ref = feats.clone()
    same = bool(torch.equal(ref, feats))
    print(f"{'FUSED (FIR in the extraction kernel)' if fused else f'two kernels, chunk {mb:3d} MB streams {st}'}: {ms:7.3f} ms/step  ({n_clips * 10 / ms / 1e3:.3f} M audio-s/s)  extract kernels {k_ms / steps:7.3f} ms in {k_n // steps} launches/step  identical_to_first={same} checksum {cs:.6f}", flush=True)
