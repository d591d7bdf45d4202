"""tcgen05 / TMEM / TMA path of the objective (precision modes tf32x3 and tf32), through the C ABI.

Stated tolerances (relative L2 vs the fp64 oracle):
  fp32   (FFMA, tests/test_gpu_parity.py)  1e-5  — the north-star bar; measured ~6e-8.
  tf32x3 (hi/lo split on the tensor cores) 2e-5  — products are fp32-accurate; what remains is the tensor core's
         truncating fp32 accumulation (a ~1e-6 systematic bias on same-signed sums), amplified in the gradient
         by ReLU units flipping near zero. Measured 1e-7 (dW, dX alone) .. 6e-6 (forward on the tensor cores).
  tf32   (single pass, 10-bit mantissa operands) 2e-2 on the gradient, 2e-3 on the loss."""
import os

import numpy as np
import pytest

import lbfgs_ffnn_b200 as P
from conftest import rel_l2
from helpers import make_gpu_net, make_problem, oracle_on_gpu_pattern, upload

pytestmark = pytest.mark.gpu

TOL = {"tf32x3": 2e-5, "tf32": 2e-2}
NETS = [([784, 128, 10], ["relu", "linear"]), ([784, 128, 64, 10], ["relu", "relu", "linear"]),
        ([784, 256, 128, 64, 10], ["tanh", "sigmoid", "relu", "linear"]), ([96, 64, 32, 12], ["relu", "tanh", "linear"])]


def _check(handle, oracle, dims, acts, batch, prec, mask=None, impose_pattern=False):
    if mask is None:
        os.environ.pop("B200_TC_MASK", None)
    else:
        os.environ["B200_TC_MASK"] = str(mask)
    P.api.reload_env()
    try:
        onet, w, X, T = make_problem(oracle, dims, acts, batch)
        loss_o, g_o = onet.loss_grad(w, X, T)
        net = make_gpu_net(handle, dims, acts, w, precision=prec)
        dx, dt = upload(X), upload(T)
        loss = net.compute_loss_and_grad(dx, dt, batch)
        g = net.get_grads()
        if impose_pattern:  # the fp64 objective on the ReLU pattern this evaluation took (helpers.oracle_on_gpu_pattern)
            loss_o, g_o = oracle_on_gpu_pattern(onet, net, acts, w, X, T)
        net.forward_only(dx, batch)
        out = net.copy_output_to_host().reshape(batch, dims[-1])
        return abs(loss - loss_o) / abs(loss_o), rel_l2(g, g_o), rel_l2(out, onet.forward(w, X))
    finally:
        os.environ.pop("B200_TC_MASK", None)
        P.api.reload_env()


@pytest.mark.parametrize("mask,name", [(1, "fwd+fused-last"), (9, "fwd"), (2, "dx"), (4, "dw"), (15, "all-unfused"), (7, "all")])
@pytest.mark.parametrize("prec", ["tf32", "tf32x3"])
def test_tc_kernels_one_role_at_a_time(handle, oracle, prec, mask, name):
    """each GEMM role on the tensor-core path with the other two on the FFMA path"""
    dims, acts = NETS[1]
    el, eg, eo = _check(handle, oracle, dims, acts, 1000, prec, mask)
    assert el <= TOL[prec] and eg <= TOL[prec] and eo <= TOL[prec], (name, el, eg, eo)


@pytest.mark.parametrize("dims,acts", NETS)
@pytest.mark.parametrize("batch", [1, 37, 1000, 4099])
@pytest.mark.parametrize("prec", ["tf32", "tf32x3"])
def test_tc_loss_grad_parity(handle, oracle, dims, acts, batch, prec):
    el, eg, eo = _check(handle, oracle, dims, acts, batch, prec)
    if prec == "tf32" and batch < 100:
        # with a handful of samples one ReLU unit flipped by a 2^-11 operand rounding moves the whole gradient
        assert el <= 2e-2 and eo <= 2e-2, (el, eo)
        return
    assert el <= TOL[prec] and eg <= TOL[prec] and eo <= TOL[prec], (el, eg, eo)


@pytest.mark.parametrize("prec", ["tf32", "tf32x3"])
def test_tc_full_size(handle, oracle, prec):
    """BASELINE configs[1]/[2] size: 60 000 samples (K = 60 000 reduction in dW)"""
    for dims, acts in NETS[:2]:
        # 11.5 M ReLU units at 60 000 samples: a handful sit within fp32 rounding of zero and land on the other side of it than
        # in fp64 (any fp32 evaluation does that, the reference's CUDA backend included). The comparison is therefore made on the
        # activation pattern the GPU took — same bound as at every other size, no 60 000-sample exemption;
        # tests/test_gpu_parity_full_size.py holds the headline mode to the north-star 1e-5 there and counts the units.
        el, eg, eo = _check(handle, oracle, dims, acts, 60000, prec, impose_pattern=True)
        assert el <= TOL[prec] and eg <= TOL[prec] and eo <= TOL[prec], (dims, el, eg, eo)


def test_tc_lbfgs_trajectory(handle, oracle):
    dims, acts, B, m, iters = [784, 128, 10], ["relu", "linear"], 1000, 10, 20
    onet, w, X, T = make_problem(oracle, dims, acts, B)
    ref = onet.lbfgs(w, X, T, m=m, max_iters=iters, tol=0.0, policy="cuda")
    net = make_gpu_net(handle, dims, acts, w, precision="tf32x3")
    s = P.CudaLBFGS(handle)
    s.setMemory(m); s.setMaxIterations(iters); s.setTolerance(0.0)
    rec = P.IterationRecorder(); rec.init(iters); s.setRecorder(rec)
    s.solve(net.params_size(), net.params_data(), upload(X), upload(T), B, net)
    loss, _, _ = rec.copy_to_host()
    for k in range(iters):
        assert abs(loss[k] - ref["loss"][k]) <= 4e-5 * (1.6 ** min(k, 20)) * abs(ref["loss"][k]), (k, loss[k], ref["loss"][k])


@pytest.mark.parametrize("prec", ["tf32", "tf32x3"])
@pytest.mark.parametrize("dims,acts", NETS[:2])
@pytest.mark.parametrize("batch", [37, 1000, 60000])
def test_u8_input_path(handle, oracle, dims, acts, batch, prec):
    """8-bit image inputs (x == float(u)/255.0f exactly): the layer-0 GEMMs read a uint8 copy (b200_net_quantize_input)"""
    onet, w, X, T = make_problem(oracle, dims, acts, batch)
    loss_o, g_o = onet.loss_grad(w, X, T)
    net = make_gpu_net(handle, dims, acts, w, precision=prec)
    dx, dt = upload(X), upload(T)
    assert net.quantize_input(dx, batch) is True
    loss = net.compute_loss_and_grad(dx, dt, batch)
    g = net.get_grads()
    if batch >= 60000:  # on the ReLU pattern the GPU took (see test_tc_full_size): same bound as the smaller batches
        loss_o, g_o = oracle_on_gpu_pattern(onet, net, acts, w, X, T)
    tol_l = TOL[prec] if prec == "tf32x3" else 2e-3
    tol_g = TOL[prec]
    if prec == "tf32" and batch < 100:
        tol_g = 1.0
    assert abs(loss - loss_o) <= tol_l * abs(loss_o), (loss, loss_o)
    assert rel_l2(g, g_o) <= tol_g, rel_l2(g, g_o)
    # the fp32-array path on the same network gives the same answer to rounding
    net.clear_input_cache()
    loss_f = net.compute_loss_and_grad(dx, dt, batch)
    g_f = net.get_grads()
    if prec == "tf32x3":
        assert abs(loss - loss_f) <= 5e-6 * abs(loss_f)
        assert rel_l2(g, g_f) <= (5e-5 if batch >= 60000 else tol_g)  # two fp32 evaluations may each take their own pattern


def test_u8_input_rejected_when_not_quantised(handle, oracle):
    dims, acts, batch = NETS[0][0], NETS[0][1], 500
    onet, w, X, T = make_problem(oracle, dims, acts, batch)
    Xn = (X + np.float32(1e-3) * np.random.RandomState(0).rand(*X.shape).astype(np.float32)).astype(np.float32)
    net = make_gpu_net(handle, dims, acts, w, precision="tf32x3")
    dx, dt = upload(Xn), upload(T)
    assert net.quantize_input(dx, batch) is False
    loss_o, g_o = onet.loss_grad(w, Xn, T)
    loss = net.compute_loss_and_grad(dx, dt, batch)
    assert abs(loss - loss_o) <= 2e-5 * abs(loss_o) and rel_l2(net.get_grads(), g_o) <= 2e-5


def test_sgd_tensorcore_subranges(handle, oracle):
    """CudaSGD slices the input by pointer offset (src/cuda/sgd.cuh:104-107): the uint8 cache must follow the slices"""
    dims, acts, B = [784, 128, 10], ["relu", "linear"], 1000
    onet, w, X, T = make_problem(oracle, dims, acts, B)
    ref = onet.sgd_cuda_policy(w, X, T, batch_size=96, lr=0.02, momentum=0.9, decay_rate=0.5, decay_step=2, max_iters=3,
                               tol=0.0, record=True)
    net = make_gpu_net(handle, dims, acts, w, precision="tf32x3")
    s = P.CudaSGD(handle)
    s.setLearningRate(0.02); s.setMomentum(0.9); s.setBatchSize(96); s.setLearningRateDecay(0.5, 2)
    s.setMaxIterations(3); s.setTolerance(0.0); s.setDimensions(784, 10)
    rec = P.IterationRecorder(); rec.init(4); s.setRecorder(rec)
    s.solve(net.params_size(), net.params_data(), upload(X), upload(T), B, net)
    loss, _, _ = rec.copy_to_host()
    assert np.allclose(loss, ref["loss"], rtol=3e-4)
    assert rel_l2(net.get_params(), ref["params"]) <= 3e-4


@pytest.mark.parametrize("prec", ["tf32", "tf32x3"])
@pytest.mark.parametrize("dims,acts,batch", [([784, 512, 256, 10], ["relu", "tanh", "linear"], 777),
                                             ([784, 1024, 1024, 10], ["relu", "relu", "linear"], 1500),
                                             ([256, 384, 96, 32, 8], ["sigmoid", "relu", "relu", "linear"], 300)])
def test_tc_multi_tile_shapes(handle, oracle, dims, acts, batch, prec):
    """hidden widths above one 128-wide tile: several N tiles in FWD/DX, several M tiles and N tiles in DW
    (the shape class of BASELINE configs[4], 784-4096-4096-10, at test size)"""
    el, eg, eo = _check(handle, oracle, dims, acts, batch, prec)
    tol_g = max(TOL[prec], 5e-5)  # ~3 M ReLU units: a few flip against the fp64 oracle (see test_tc_full_size)
    assert el <= TOL[prec] and eg <= tol_g and eo <= TOL[prec], (el, eg, eo)
