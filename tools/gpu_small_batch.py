"""Small-batch training epochs (batch 8, the reference's default): per-step time with and without the captured two-step graph,
   for the tensor-core (3xTF32) and the FP32 CUDA-core arithmetic.   python tools/gpu_small_batch.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import streamz_b200 as sz
from streamz_b200 import _native as N
r = np.random.default_rng(0)
n, batch = 8 * 600, int(sys.argv[1]) if len(sys.argv) > 1 else 8
feats = r.standard_normal((n, 60)).astype(np.float32)
labels = r.integers(0, 2, n).astype(np.uint32)
for graphs in (True, False):
    if graphs: os.environ.pop("SZB_NO_GRAPHS", None)
    else: os.environ["SZB_NO_GRAPHS"] = "1"
    ctx = sz.Context(0)
    for mode in ("3xtf32", "fp32"):
        net = sz.SimpleNeuralNet(60, 512, 256, 2, seed=1, ctx=ctx).set_precision(mode)
        data = sz.DeviceFeatures(ctx, feats, labels)
        perm = r.permutation(n).astype(np.uint32)
        sz.train_epoch(net, data, perm, batch, 0.01, dropout=0.2, seed=1, stream=0)
        ctx.sync(); t0 = time.perf_counter()
        for e in range(3):
            sz.train_epoch(net, data, perm, batch, 0.01, dropout=0.2, seed=1, stream=1 + e)
        ctx.sync(); dt = time.perf_counter() - t0
        steps = 3 * ((n + batch - 1) // batch)
        print(f"graphs={graphs} {mode}: batch {batch}: {dt / steps * 1e6:.1f} us per step (wall), graph replays so far {int(N.lib.szb_ctx_graph_launch_count(ctx.handle))}", flush=True)
        data.close(); net.close()
    ctx.close()
