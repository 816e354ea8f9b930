"""A few MLP training steps at the configs[2] shape (batch 4096, 60-512-256-100) for profiling."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import streamz_b200 as sz
mode = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = sz.Context(0)
r = np.random.default_rng(0)
n = 4096 * steps
feats = r.standard_normal((n, 60)).astype(np.float32)
labels = r.integers(0, 100, n).astype(np.uint32)
net = sz.SimpleNeuralNet(60, 512, 256, 100, seed=1, ctx=ctx).set_precision(mode)
data = sz.DeviceFeatures(ctx, feats, labels)
perm = r.permutation(n).astype(np.uint32)
sz.train_epoch(net, data, perm, 4096, 0.01, dropout=0.2, seed=1, stream=0)
ctx.timer_start()
loss, used = sz.train_epoch(net, data, perm, 4096, 0.01, dropout=0.2, seed=1, stream=1)
ms = ctx.timer_stop()
print(mode, "steps", steps, "ms/step", ms / steps, "loss", loss / used)
