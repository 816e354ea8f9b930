"""The opt-in tcgen05 GEMM with the A operand in tensor memory (gemm_tc_ta_kernel) against the default
shared-memory-operand kernel and against the oracle.  SZB_GEMM_TA is read when a context is created, so the test makes a
second context for it.  The 3xTF32 split and the MMA order are the same in both kernels: forward results must be equal
bit for bit; a training step may differ by the order of the split-K atomics only."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dims,B", [((60, 512, 256, 100), 700), ((60, 200, 72, 10), 130)])
def test_tmem_a_kernel_equals_default_kernel(sz, ctx, oracle, monkeypatch, dims, B):
    monkeypatch.setenv("SZB_GEMM_TA", "1")
    ctx_ta = sz.Context(0)
    monkeypatch.delenv("SZB_GEMM_TA")
    onet = oracle.Net.init(*dims, seed=B)
    onet.b1[:] = np.random.default_rng(B).uniform(-.1, .1, dims[1])
    net = sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx)
    net_ta = sz.SimpleNeuralNet.from_weights(*onet.params(), ctx=ctx_ta)
    try:
        x = np.random.default_rng(B + 1).standard_normal((B, dims[0])).astype(np.float32)
        p, p_ta = net.forward(x), net_ta.forward(x)
        assert np.array_equal(p, p_ta), f"forward differs: max {np.abs(p - p_ta).max()}"
        assert np.abs(p_ta - oracle.forward(onet.copy(np.float64), x)).max() <= 5e-5        # the 3xTF32 bar of test_gpu_mlp.py
        t = np.zeros(dims[3], np.float32)
        t[1] = 1
        net.train_batch(x, t, 0.01)
        net_ta.train_batch(x, t, 0.01)
        oracle.train_batch(onet, x, t, 0.01)
        for a, b, o in zip(net.weights(), net_ta.weights(), onet.params()):
            assert np.abs(a - b).max() <= 1e-6          # same products, atomics in a different order
            assert np.abs(b - o).max() <= 1e-5          # STEP_TOL of the default mode
    finally:
        net_ta.close()
        ctx_ta.close()
