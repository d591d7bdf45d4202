"""The fp16 tensor-core path of layer 0 (8-bit-pixel inputs), the one-pass last layer and the fused direction kernel.

Kernels under test (all reached through the C ABI):
  fwd16   gemm_fwd16.cu  persistent forward of layer 0: exact fp16 X  x  scaled fp16 {hi, lo} weights (22 mantissa bits)
  dw16    gemm_dw16.cu   split-K [dW_0; db_0]: X^T x [delta_hi | delta_lo], bias gradient from the ones column of the fp16 X copy
  tail    tail_layer.cu  last layer forward + loss + both deltas + [dW_L; db_L] in two passes over A_{L-1} (fp32 FFMA); the forward
          pass has a sample-per-lane form (default) and a feature-per-lane form (B200_TAIL_FWD=1)
  lbfgs_direction_kernel  dots -> grid barrier -> solve -> apply in one launch
Each can be switched off with an environment variable read per call, so the same network is evaluated through the generic
kernels and through the new ones and both are compared with the fp64 oracle.

Stated tolerances (relative L2 vs the fp64 oracle): loss / gradient 2e-5 in the fp32-accurate mode (tf32x3; measured 2e-7 .. 4e-6),
2e-3 / 2e-2 in the single-pass mode (tf32), 1e-5 for the fp32 FFMA mode (the tail serves it too)."""
import os

import numpy as np
import pytest

import lbfgs_ffnn_b200 as P
from conftest import rel_l2
from helpers import make_gpu_net, make_problem, relu_pattern_of, upload

pytestmark = pytest.mark.gpu

TOGGLES = ("B200_FWD16", "B200_TAIL", "B200_DW16", "B200_TAIL_FWD", "B200_MID16", "B200_PAIR", "B200_PDL", "B200_RING")


def _eval(handle, dims, acts, w, X, T, prec, env=None, quantize=True, want_pattern=False):
    saved = {k: os.environ.get(k) for k in TOGGLES}
    try:
        for k in TOGGLES:
            os.environ.pop(k, None)
        for k, v in (env or {}).items():
            os.environ[k] = v
        P.api.reload_env()
        B = X.shape[0]
        net = make_gpu_net(handle, dims, acts, w, precision=prec)
        dx, dt = upload(X), upload(T)
        if quantize and prec != "fp32":
            assert net.quantize_input(dx, B) is True
        loss = net.compute_loss_and_grad(dx, dt, B)
        g = net.get_grads()
        pattern = relu_pattern_of(net, acts) if want_pattern else None
        net.forward_only(dx, B)
        out = net.copy_output_to_host().reshape(B, dims[-1])
        return (loss, g, out, pattern) if want_pattern else (loss, g, out)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        P.api.reload_env()


def problem8(oracle, dims, acts, batch, seed=123):
    """8-bit-pixel inputs (x == float(u)/255.0f) for any output width"""
    onet, w, X, T = make_problem(oracle, dims, acts, batch)
    if dims[-1] != 10:
        X, _ = P.synthetic_mnist(batch, seed=seed)
        rs = np.random.RandomState(seed)
        T = np.zeros((batch, dims[-1]), dtype=np.float32)
        T[np.arange(batch), rs.randint(0, dims[-1], batch)] = 1
    return onet, w, X, T


NETS = [([784, 128, 10], ["relu", "linear"]), ([784, 64, 10], ["tanh", "linear"]), ([784, 32, 10], ["sigmoid", "linear"]),
        ([784, 128, 12], ["relu", "sigmoid"]), ([784, 64, 7], ["relu", "tanh"]), ([784, 128, 64, 10], ["relu", "relu", "linear"])]


@pytest.mark.parametrize("dims,acts", NETS)
@pytest.mark.parametrize("batch", [1, 31, 128, 129, 1000, 4099])
def test_fp16_path_parity(handle, oracle, dims, acts, batch):
    onet, w, X, T = problem8(oracle, dims, acts, batch)
    lo, go = onet.loss_grad(w, X, T)
    fo = onet.forward(w, X)
    loss, g, out = _eval(handle, dims, acts, w, X, T, "tf32x3")
    assert abs(loss - lo) <= 2e-5 * abs(lo), (loss, lo)
    assert rel_l2(g, go) <= 2e-5, rel_l2(g, go)
    assert rel_l2(out, fo) <= 2e-5, rel_l2(out, fo)


@pytest.mark.parametrize("env", [{"B200_DW16": "0"}, {"B200_TAIL": "0"}, {"B200_FWD16": "0"}, {"B200_TAIL_FWD": "1"}, {"B200_MID16": "0"},
                                 {"B200_FWD16": "0", "B200_TAIL": "0", "B200_DW16": "0"}])
@pytest.mark.parametrize("which", [0, 5])
def test_each_new_kernel_against_the_generic_path(handle, oracle, env, which):
    """the same evaluation with one (or all) of the new kernels replaced by the generic tcgen05 / FFMA kernels"""
    dims, acts, batch = NETS[which][0], NETS[which][1], 2500
    onet, w, X, T = make_problem(oracle, dims, acts, batch)
    lo, go = onet.loss_grad(w, X, T)
    l1, g1, o1 = _eval(handle, dims, acts, w, X, T, "tf32x3")
    l0, g0, o0 = _eval(handle, dims, acts, w, X, T, "tf32x3", env)
    for loss, g in ((l1, g1), (l0, g0)):
        assert abs(loss - lo) <= 2e-5 * abs(lo)
        assert rel_l2(g, go) <= 2e-5
    assert rel_l2(g1, g0) <= 2e-5 and rel_l2(o1, o0) <= 2e-5


@pytest.mark.parametrize("prec,tol_l,tol_g", [("fp32", 1e-5, 1e-5), ("tf32", 2e-3, 2e-2)])
def test_tail_serves_the_other_precision_modes(handle, oracle, prec, tol_l, tol_g):
    for dims, acts in (NETS[0], NETS[5]):
        onet, w, X, T = make_problem(oracle, dims, acts, 1500)
        lo, go = onet.loss_grad(w, X, T)
        loss, g, _ = _eval(handle, dims, acts, w, X, T, prec)
        assert abs(loss - lo) <= tol_l * abs(lo), (prec, dims, loss, lo)
        assert rel_l2(g, go) <= tol_g, (prec, dims, rel_l2(g, go))


@pytest.mark.parametrize("which", [0, 5])
def test_full_size_fp16_path(handle, oracle, which):
    """BASELINE configs[1] / configs[2] size: 60 000 samples, 469 tiles over 148 persistent CTAs, 37-way split-K in dw16"""
    dims, acts = NETS[which]
    onet, w, X, T = make_problem(oracle, dims, acts, 60000)
    loss, g, _, pattern = _eval(handle, dims, acts, w, X, T, "tf32x3", want_pattern=True)
    # 7.7 M ReLU units: a few sit within fp32 rounding of zero and land on the other side of it than in fp64; the fp64 objective is
    # evaluated on the pattern the GPU took (helpers.oracle_on_gpu_pattern) and the bound is the one of every other size
    lo, go = onet.loss_grad_masked(w, X, T, pattern)
    assert abs(loss - lo) <= 2e-5 * abs(lo) and rel_l2(g, go) <= 2e-5, (loss, lo, rel_l2(g, go))


@pytest.mark.parametrize("which", [0, 5])
def test_scaled_fp16_operands_survive_extreme_weights(handle, oracle, which):
    """per-neuron power-of-two scales: weights spanning 1e-6 .. 1e+3 across neurons, and a huge / tiny delta. In the deep net
    delta_0 is scaled by a bound chained through the weight matrices of the later layers (tail_layer.cu)"""
    dims, acts, batch = NETS[which][0], NETS[which][1], 700
    onet, w, X, T = make_problem(oracle, dims, acts, batch)
    rs = np.random.RandomState(5)
    w = w.copy()
    W0 = w[:784 * 128].reshape(784, 128)
    W0 *= (10.0 ** rs.uniform(-6, 3, size=128)).astype(np.float32)[None, :]
    for tscale in (1.0, 1e4, 1e-4):
        Ts = (T * np.float32(tscale)).astype(np.float32)
        lo, go = onet.loss_grad(w, X, Ts)
        loss, g, _ = _eval(handle, dims, acts, w, X, Ts, "tf32x3")
        assert np.isfinite(loss) and np.all(np.isfinite(g))
        assert abs(loss - lo) <= 2e-5 * abs(lo), (tscale, loss, lo)
        assert rel_l2(g, go) <= 2e-5, (tscale, rel_l2(g, go))


def test_input_copy_follows_the_buffer_contents(handle, oracle):
    """a minimisation re-derives the 8-bit copy from what the device buffer holds when it starts"""
    dims, acts, batch, iters = NETS[0][0], NETS[0][1], 600, 6
    onet, w, X, T = make_problem(oracle, dims, acts, batch)
    X2, T2 = P.synthetic_mnist(batch, seed=77)
    net = make_gpu_net(handle, dims, acts, w, precision="tf32x3")
    dx, dt = upload(X), upload(T)

    def solve():
        net.set_params(w)
        s = P.CudaLBFGS(handle)
        s.setMemory(10); s.setMaxIterations(iters); s.setTolerance(0.0)
        rec = P.IterationRecorder(); rec.init(iters); s.setRecorder(rec)
        s.solve(net.params_size(), net.params_data(), dx, dt, batch, net)
        return rec.copy_to_host()[0]

    la = solve()
    dx.copy_from_host(X2); dt.copy_from_host(T2)  # same device buffers, new data
    lb = solve()
    ra = onet.lbfgs(w, X, T, m=10, max_iters=iters, tol=0.0, policy="cuda")["loss"]
    rb = onet.lbfgs(w, X2, T2, m=10, max_iters=iters, tol=0.0, policy="cuda")["loss"]
    assert np.allclose(la, ra, rtol=1e-3) and np.allclose(lb, rb, rtol=1e-3), (la, ra, lb, rb)


def test_pooled_solver_and_graphs_are_reused_safely(handle, oracle):
    """solver objects are parked and reused (work space + CUDA graphs): repeated and interleaved solves stay correct"""
    dims, acts, iters = NETS[0][0], NETS[0][1], 8
    onet, w, X, T = make_problem(oracle, dims, acts, 900)
    net = make_gpu_net(handle, dims, acts, w, precision="tf32x3")
    dx, dt = upload(X), upload(T)

    def solve(batch, prec):
        net.set_precision(prec)
        net.set_params(w)
        s = P.CudaLBFGS(handle)
        s.setMemory(10); s.setMaxIterations(iters); s.setTolerance(0.0)
        rec = P.IterationRecorder(); rec.init(iters); s.setRecorder(rec)
        s.solve(net.params_size(), net.params_data(), dx, dt, batch, net)
        return rec.copy_to_host()[0]

    ref900 = onet.lbfgs(w, X, T, m=10, max_iters=iters, tol=0.0, policy="cuda")["loss"]
    ref500 = onet.lbfgs(w, X[:500], T[:500], m=10, max_iters=iters, tol=0.0, policy="cuda")["loss"]
    a = solve(900, "tf32x3")
    b = solve(900, "tf32x3")     # same shape: the parked solver and its graphs
    c = solve(500, "tf32x3")     # other batch: the graphs must be dropped
    d = solve(900, "fp32")       # other precision: other kernels behind the same pointers
    e = solve(900, "tf32x3")
    assert np.array_equal(a, b) and np.array_equal(a, e)
    assert np.allclose(a, ref900, rtol=1e-3) and np.allclose(d, ref900, rtol=1e-3) and np.allclose(c, ref500, rtol=1e-3)


@pytest.mark.parametrize("m", [0, 1, 3, 10, 20])
def test_fused_direction_kernel_matches_the_three_kernel_path(handle, oracle, m):
    dims, acts, batch, iters = NETS[0][0], NETS[0][1], 800, 14
    onet, w, X, T = make_problem(oracle, dims, acts, batch)
    dx, dt = upload(X), upload(T)
    out = {}
    for fused in (True, False):
        if fused:
            os.environ.pop("B200_NO_FUSED_DIRECTION", None)
        else:
            os.environ["B200_NO_FUSED_DIRECTION"] = "1"
        P.api.reload_env()
        try:
            net = make_gpu_net(handle, dims, acts, w, precision="fp32")
            s = P.CudaLBFGS(handle)
            s.setMemory(m); s.setMaxIterations(iters); s.setTolerance(0.0)
            rec = P.IterationRecorder(); rec.init(iters); s.setRecorder(rec)
            s.solve(net.params_size(), net.params_data(), dx, dt, batch, net)
            out[fused] = (rec.copy_to_host()[0], net.get_params())
        finally:
            os.environ.pop("B200_NO_FUSED_DIRECTION", None)
            P.api.reload_env()
    # identical arithmetic in the same order: bit-identical trajectories
    assert np.array_equal(out[True][0], out[False][0])
    assert np.array_equal(out[True][1], out[False][1])
    ref = onet.lbfgs(w, X, T, m=m, max_iters=iters, tol=0.0, policy="cuda")
    assert np.allclose(out[True][0], ref["loss"], rtol=2e-3)


@pytest.mark.parametrize("dims", [[784, 64, 32, 10], [784, 128, 128, 10], [784, 128, 64, 32, 10], [784, 64, 64, 64, 5],
                                  [784, 128, 64, 64, 64, 10], [256, 128, 64, 10]])
def test_deeper_nets_chain_the_fp16_delta_scale(handle, oracle, dims):
    """three to six layers: the layer-1 dX kernel emits the fp16 delta_0 with the scale bound chained through every later
    weight matrix, the middle layers run the two-stage generic kernel (tools/shape_sweep.py sweeps more shapes)"""
    acts = ["relu"] * (len(dims) - 2) + ["linear"]
    for batch in (33, 1500):
        X, _ = P.synthetic_mnist(batch, seed=5)
        X = np.ascontiguousarray(X[:, :dims[0]])
        rs = np.random.RandomState(1)
        T = np.zeros((batch, dims[-1]), dtype=np.float32)
        T[np.arange(batch), rs.randint(0, dims[-1], batch)] = 1
        onet = oracle.OracleNet(dims, acts)
        w = onet.init_params_cuda_rule(123).astype(np.float32)
        lo, go = onet.loss_grad(w, X, T)
        loss, g, _ = _eval(handle, dims, acts, w, X, T, "tf32x3")
        assert abs(loss - lo) <= 2e-5 * abs(lo), (dims, batch, loss, lo)
        assert rel_l2(g, go) <= 2e-5, (dims, batch, rel_l2(g, go))


@pytest.mark.parametrize("which", [0, 5])
@pytest.mark.parametrize("batch", [1000, 60000])
def test_evaluation_is_bit_reproducible(handle, oracle, which, batch):
    """the same evaluation 12 times: every kernel combines its partial results in a fixed order, so loss and gradient must be
    BIT-identical run to run. (compute-sanitizer is closed on this pool — profiles/r02_sanitizer.md — so this is the race detector
    the suite has for the mbarrier / TMA / TMEM pipelines: a missing wait or fence shows up as a run that differs.)"""
    dims, acts = NETS[which]
    onet, w, X, T = make_problem(oracle, dims, acts, batch)
    net = make_gpu_net(handle, dims, acts, w, precision="tf32x3")
    dx, dt = upload(X), upload(T)
    assert net.quantize_input(dx, batch)
    ref_loss, ref_g = None, None
    for rep in range(12):
        loss = net.compute_loss_and_grad(dx, dt, batch)
        g = net.get_grads()
        if rep == 0:
            ref_loss, ref_g = loss, g
        else:
            assert loss == ref_loss and np.array_equal(g, ref_g), rep


@pytest.mark.parametrize("dims,acts", [([784, 128, 64, 10], ["relu", "relu", "linear"]), ([784, 128, 10], ["relu", "linear"])])
@pytest.mark.parametrize("batch", [700, 5000])
def test_cta_pair_forward_is_bit_identical_to_the_one_cta_form(handle, oracle, dims, acts, batch):
    """cta_group::2 (two SMs share one M = 256 MMA, the weights split between them) runs the same MMA shapes per SM in the same
    order as the one-CTA kernel: loss, gradient and outputs must agree bit for bit; so must the deeper X ring (B200_RING=1)"""
    onet, w, X, T = problem8(oracle, dims, acts, batch)
    base = _eval(handle, dims, acts, w, X, T, "tf32x3", env={"B200_PAIR": "0"})
    for env in ({"B200_PAIR": "1"}, {"B200_PAIR": "0", "B200_RING": "3"}):
        got = _eval(handle, dims, acts, w, X, T, "tf32x3", env=env)
        assert got[0] == base[0], env
        assert np.array_equal(got[1], base[1]), env
        assert np.array_equal(got[2], base[2]), env
    lo, go = onet.loss_grad(w, X, T)
    assert abs(base[0] - lo) <= 2e-5 * abs(lo) and rel_l2(base[1], go) <= 2e-5


def test_programmatic_dependent_launch_changes_no_result(handle, oracle):
    """B200_PDL=0 (plain stream order) against the default (every kernel launched programmatically, griddepcontrol.wait at its
    top): a 12-iteration L-BFGS run through the captured graph must produce the same losses and parameters bit for bit"""
    dims, acts, batch, iters = [784, 128, 64, 10], ["relu", "relu", "linear"], 3000, 12
    onet, w, X, T = problem8(oracle, dims, acts, batch)
    dx, dt = upload(X), upload(T)
    out = {}
    saved = os.environ.get("B200_PDL")
    try:
        for pdl in ("1", "0"):
            os.environ["B200_PDL"] = pdl
            P.api.reload_env()
            net = make_gpu_net(handle, dims, acts, w, precision="tf32x3")
            s = P.CudaLBFGS(handle)
            s.setMemory(10); s.setMaxIterations(iters); s.setTolerance(0.0)
            rec = P.IterationRecorder(); rec.init(iters); s.setRecorder(rec)
            s.solve(net.params_size(), net.params_data(), dx, dt, batch, net)
            out[pdl] = (rec.copy_to_host()[0], net.get_params())
    finally:
        if saved is None:
            os.environ.pop("B200_PDL", None)
        else:
            os.environ["B200_PDL"] = saved
        P.api.reload_env()
    assert np.array_equal(out["1"][0], out["0"][0])
    assert np.array_equal(out["1"][1], out["0"][1])
