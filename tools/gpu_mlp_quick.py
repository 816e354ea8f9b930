import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import streamz_b200 as sz, streamz_oracle as orc
ctx = sz.Context(0)
for dims, B in (((60, 512, 256, 100), 300), ((60, 512, 256, 7), 4096), ((4, 3, 2, 2), 5), ((17, 33, 65, 9), 130)):
    onet = orc.Net.init(*dims, seed=1)
    onet.b1[:] = 0.05; onet.b2[:] = -0.03; onet.b3[:] = 0.01
    x = np.random.default_rng(0).standard_normal((B, dims[0])).astype(np.float32)
    q = orc.forward(onet.copy(np.float64), x)
    for mode in ("fp32", "3xtf32", "tf32"):
        net = sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx).set_precision(mode)
        p = net.forward(x)
        print(dims, B, mode, "fwd max err", float(np.abs(p - q).max()))
        o2 = onet.copy()
        t = np.zeros(dims[3], np.float32); t[0] = 1
        net.train_batch(x, t, 0.01); orc.train_batch(o2, x, t, 0.01)
        print("   train max werr", max(float(np.abs(a - b).max()) for a, b in zip(net.weights(), o2.params())))
