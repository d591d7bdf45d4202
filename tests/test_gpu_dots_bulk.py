"""The history pass of the two-loop recursion for LONG vectors: lbfgs_dots_bulk_kernel (lbfgs_kernels.cu, "(1b)").

It replaces, like lbfgs_dots_kernel, the 2k cublasSdot calls of the reference's two-loop recursion (src/cuda/lbfgs.cuh:206-261) and the
forming of the new pair (lbfgs.cuh:143-150) by one pass over the ring; it takes over above B200_DOTS_BULK_MIN elements (default 2^20:
BASELINE configs[4] has 2·10^7), where the history no longer fits L2. Every ring row's chunk of a 256-element tile arrives in a
shared-memory ring by bulk async copies; the partials it writes have the layout of the register-staged kernel's, so everything after
it is shared. Here B200_DOTS_BULK_MIN=0 sends small problems through it and the results are compared with the fp64 oracle and with
the register-staged kernel (B200_DOTS_BULK=0): fp32 outputs within 1e-6 relative L2 of the oracle (the two kernels sum the same
fp64 products in different orders)."""
import os

import numpy as np
import pytest

import lbfgs_ffnn_b200 as P
from lbfgs_ffnn_b200 import api
from conftest import rel_l2
from helpers import make_gpu_net, upload

pytestmark = pytest.mark.gpu


class _Env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.saved = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        api.reload_env()

    def __exit__(self, *exc):
        for k, v in self.saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        api.reload_env()


# the forms of the bulk kernel (eight consumer warps with two 16-byte positions per lane; sixteen with one; sixteen with the shared
# conversion of the tile's vectors) and the register-staged one
_VARIANTS = (("bulk", dict(B200_DOTS_BULK="1", B200_DOTS_BULK_MIN="0")), ("bulk16", dict(B200_DOTS_BULK="2", B200_DOTS_BULK_MIN="0")),
             ("bulk16s", dict(B200_DOTS_BULK="3", B200_DOTS_BULK_MIN="0")),  # g / s_new / y_new converted once per CTA
             ("staged", dict(B200_DOTS_BULK="0", B200_DOTS_BULK_MIN=None)))  # (bulk16s is the default form)


def _history(rs, k, n):
    S = rs.randn(k, n).astype(np.float32) * 0.01
    Y = (0.7 * S + 0.02 * rs.randn(k, n).astype(np.float32) * 0.01).astype(np.float32)
    rho = np.array([1.0 / np.dot(S[i].astype(np.float64), Y[i].astype(np.float64)) for i in range(k)], dtype=np.float32)
    return S, Y, rho


# rows per consumer warp 1 … 4 (k + 1 ring rows over 8 warps); n a multiple of 4 (the direct API's rows are n apart), a short last
# tile (n % 256 != 0), fewer tiles than CTAs (n = 1000), more tiles than one round of CTAs (n = 400 000: 1563 tiles over 296 CTAs)
@pytest.mark.parametrize("policy", ["cpu", "cuda", "slbfgs"])
@pytest.mark.parametrize("k,n", [(0, 1000), (1, 8), (5, 1000), (10, 101772), (20, 250004), (20, 400000), (30, 65536), (31, 5004)])
def test_bulk_dots_direction_matches_the_oracle(handle, oracle, policy, k, n):
    rs = np.random.RandomState(k * 7 + 1)
    S, Y, rho = _history(rs, k, n)
    g = rs.randn(n).astype(np.float32)
    p_o = oracle.direction(S, Y, rho, g, policy)
    gdp_o = float(np.dot(g.astype(np.float64), p_o))
    dS, dY, dg = (upload(S) if k else None), (upload(Y) if k else None), upload(g)
    res = {}
    for name, env in _VARIANTS:
        with _Env(**env):
            dp = P.DeviceBuffer(n)
            gdp = api.lbfgs_direction(handle, dS, dY, rho, dg, n, k, dp, policy)
            res[name] = (dp.copy_to_host(), gdp)
    for name, (p, gdp) in res.items():
        assert rel_l2(p, p_o) <= 1e-6, (name, rel_l2(p, p_o))
        assert abs(gdp - gdp_o) <= 1e-5 * abs(gdp_o), (name, gdp, gdp_o)
    for name in ("bulk", "bulk16", "bulk16s"):
        assert rel_l2(res[name][0], res["staged"][0]) <= 1e-6
        assert abs(res[name][1] - res["staged"][1]) <= 1e-9 * abs(gdp_o)


# the solver's own use: the pair is FORMED inside the pass (s = x - x_prev, y = g - g_prev written to the ring slot), the ring wraps
# (m = 5, 25 iterations), and n % 4 = 2 for both nets (the two elements past the last multiple of four take the scalar path)
@pytest.mark.parametrize("dims,acts", [([784, 128, 10], ["relu", "linear"]), ([784, 128, 64, 10], ["relu", "relu", "linear"]),
                                       ([784, 128, 12], ["tanh", "linear"])])
def test_lbfgs_with_the_bulk_dots_kernel_follows_the_staged_one_and_the_oracle(handle, oracle, dims, acts):
    B, iters, m = 1500, 25, 5
    X, _ = P.synthetic_mnist(B)
    rs = np.random.RandomState(2)
    T = np.zeros((B, dims[-1]), dtype=np.float32)
    T[np.arange(B), rs.randint(0, dims[-1], B)] = 1
    onet = oracle.OracleNet(dims, acts)
    w = onet.init_params_cuda_rule(123).astype(np.float32)
    dx, dt = upload(X), upload(T)
    runs = {}
    for name, env in _VARIANTS:
        with _Env(B200_NO_FUSED_DIRECTION="1", **env):
            net = make_gpu_net(handle, dims, acts, w, precision="fp32")
            s = P.CudaLBFGS(handle)
            s.setMemory(m); s.setMaxIterations(iters); s.setTolerance(0.0)
            rec = P.IterationRecorder(); rec.init(iters); s.setRecorder(rec)
            s.solve(net.params_size(), net.params_data(), dx, dt, B, net)
            runs[name] = (rec.copy_to_host()[0].astype(np.float64), net.get_params().copy())
    lb, ls = runs["bulk"][0], runs["staged"][0]
    assert len(lb) == len(ls) == len(runs["bulk16"][0]) == iters
    assert np.allclose(lb, ls, rtol=2e-5), np.max(np.abs(lb - ls) / ls)
    assert np.allclose(runs["bulk16"][0], ls, rtol=2e-5) and np.allclose(runs["bulk16s"][0], ls, rtol=2e-5)
    assert len(runs["bulk16s"][0]) == iters
    ref = onet.lbfgs(w, X, T, m=m, max_iters=iters, tol=0.0, policy="cuda")
    # (fp32 and fp64 trajectories part company slowly: 2e-3 over the first ten iterations, 1e-2 over all 25)
    assert np.allclose(lb[:10], ref["loss"][:10], rtol=2e-3) and np.allclose(lb, ref["loss"], rtol=1e-2), (lb, ref["loss"])
    assert lb[-1] < 0.7 * lb[0]
