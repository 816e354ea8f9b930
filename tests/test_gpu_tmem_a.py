"""The three generations of the tcgen05 GEMM against one another and against the oracle:
gemm_tma_kernel (default: operands by TMA, A operand in tensor memory, two producer groups), gemm_tc_ta_kernel
(SZB_GEMM_TMA=0: cp.async operands, A operand in tensor memory) and gemm_tc_async_kernel (SZB_GEMM_TMA=0 SZB_GEMM_TA=0: both
operands from shared memory).  The switches are read when a context is created, so the test makes a context per kernel.  The
3xTF32 split and the MMA order are the same in all three: forward results must be equal bit for bit; a training step may
differ by the order of the split-K reductions only (and by the grouped weight-gradient launch of the default step)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dims,B", [((60, 512, 256, 100), 700), ((60, 200, 72, 10), 130), ((60, 512, 256, 1000), 260)])
def test_gemm_generations_agree(sz, ctx, oracle, monkeypatch, dims, B):
    others = []
    for env in ({"SZB_GEMM_TMA": "0"}, {"SZB_GEMM_TMA": "0", "SZB_GEMM_TA": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        others.append(sz.Context(0))
        for k in env:
            monkeypatch.delenv(k)
    onet = oracle.Net.init(*dims, seed=B)
    onet.b1[:] = np.random.default_rng(B).uniform(-.1, .1, dims[1])
    net = sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx)
    nets = [sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=c) for c in others]
    try:
        x = np.random.default_rng(B + 1).standard_normal((B, dims[0])).astype(np.float32)
        p = net.forward(x)
        for n in nets:
            q = n.forward(x)
            assert np.array_equal(p, q), f"forward differs: max {np.abs(p - q).max()}"
        assert np.abs(p - oracle.forward(onet.copy(np.float64), x)).max() <= 5e-5        # the 3xTF32 bar of test_gpu_mlp.py
        t = np.zeros(dims[3], np.float32)
        t[1] = 1
        net.train_batch(x, t, 0.01)
        oracle.train_batch(onet, x, t, 0.01)
        for n in nets:
            n.train_batch(x, t, 0.01)
            for a, b in zip(net.weights(), n.weights()):
                assert np.abs(a - b).max() <= 1e-6      # same products, reductions in a different order
        for a, o in zip(net.weights(), onet.params()):
            assert np.abs(a - o).max() <= 1e-5          # STEP_TOL of the default mode
    finally:
        for n in nets:
            n.close()
        for c in others:
            c.close()
