"""Ad-hoc GPU check used during development (not a test): smoke + quick extraction timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import __graft_entry__ as g
try:
    g.smoke()
except Exception as e:
    import traceback; traceback.print_exc()
import streamz_b200 as sz, streamz_oracle as orc
ctx = sz.Context(0); ex = sz.FeatureExtractor(ctx)
for L in (0, 799, 800, 1199, 1200, 1201, 2000, 13200, 13201, 44100 * 3 + 17):
    clip = orc.synth_clip(1, L, max(L, 1) / 44100.0)[:L]
    got = ex.extract(clip); want = orc.extract(clip)
    print("len", L, got.shape, "err", float(np.abs(got - want).max()) if len(want) else 0.0)
# odd offsets in a packed batch
clips = [orc.synth_clip(i % 3, i, 0.3 + 0.01 * i)[: 13000 + 37 * i + (i % 2)] for i in range(7)]
outs = ex.extract_batch(clips)
print("batch err", max(float(np.abs(o - orc.extract(c)).max()) for o, c in zip(outs, clips)))
# throughput: device-resident 44.1k, 2000 clips x 10 s
n_clips, L = 2000, 441000
base = orc.synth_clip(2, 1, 10.0)
pcm = np.tile(base, 8)
d_pcm = ctx.dev_alloc(n_clips * L * 2)
for i in range(n_clips // 8):
    ctx.h2d(d_pcm + i * 8 * L * 2, pcm)
off = (np.arange(n_clips + 1, dtype=np.uint64) * L)
import ctypes as C
from streamz_b200 import _native as N
total = int(N.lib.szb_extract_batch_windows(N.ptr(off), n_clips, 44100))
d_out = ctx.dev_alloc(total * 240)
woff = np.zeros(n_clips + 1, np.uint64)
for it in range(3):
    ctx.timer_start()
    N.check(N.lib.szb_extract_batch_dev(ctx.handle, C.c_void_p(d_pcm), N.ptr(off), n_clips, 44100, C.c_void_p(d_out), total, N.ptr(woff)))
    ms = ctx.timer_stop()
    print(f"extract 44.1k: {n_clips} clips x 10 s: {ms:.2f} ms -> {n_clips*10/ms*1e3:.3e} audio-s/s, {total/ms*1e3:.3e} win/s, {total*1040/ms/1e6:.1f} GB/s algorithmic")
out = np.empty((1101, 60), np.float32); ctx.d2h(out, d_out + (n_clips - 1) * 1101 * 240)
print("last clip err", float(np.abs(out - orc.extract(base)).max()))
