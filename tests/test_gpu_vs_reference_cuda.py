"""GPU tests against the REFERENCE'S OWN CUDA backend (oracle/_ref/libref_cuda.so: cuBLAS SGEMM + src/cuda/*.cuh
compiled unmodified from /root/reference by oracle/ref_cuda/). Same inputs, same initial parameters (the
reference's bindParams rule, which this library reproduces), same algorithm (Armijo L-BFGS, GD, SGD)."""
import numpy as np
import pytest

import lbfgs_ffnn_b200 as P
from conftest import rel_l2
from helpers import make_gpu_net, upload
from oracle import ref_cuda_binding as rc

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rc.available(), reason="oracle/_ref/libref_cuda.so not built")]

ACT = {"linear": 0, "tanh": 1, "relu": 2, "sigmoid": 3}
NETS = [([784, 128, 10], ["relu", "linear"]), ([784, 128, 64, 10], ["relu", "relu", "linear"])]


@pytest.mark.parametrize("dims,acts", NETS)
def test_bind_params_identical_to_reference(handle, dims, acts):
    ref = rc.RefCudaNet(dims, [ACT[a] for a in acts])
    ref.bind_params(7)
    net = make_gpu_net(handle, dims, acts)
    net.bindParams(7)
    assert np.array_equal(net.get_params(), ref.get_params())  # bit-identical initial parameters


@pytest.mark.parametrize("dims,acts", NETS)
@pytest.mark.parametrize("prec,tol", [("fp32", 2e-6), ("tf32x3", 2e-5)])
def test_loss_grad_vs_reference_cuda(handle, dims, acts, prec, tol):
    B = 2000
    X, T = P.synthetic_mnist(B)
    dx, dt = upload(X), upload(T)
    ref = rc.RefCudaNet(dims, [ACT[a] for a in acts])
    ref.bind_params(123)
    w = ref.get_params()
    lr, gr = ref.loss_grad(dx.data(), dt.data(), B)
    net = make_gpu_net(handle, dims, acts, w, precision=prec)
    loss = net.compute_loss_and_grad(dx, dt, B)
    assert abs(loss - lr) <= tol * abs(lr)
    assert rel_l2(net.get_grads(), gr) <= tol
    net.forward_only(dx, B)
    assert rel_l2(net.copy_output_to_host().reshape(B, -1), ref.forward(dx.data(), B)) <= tol


@pytest.mark.parametrize("dims,acts", NETS)
def test_lbfgs_trajectory_vs_reference_cuda(handle, dims, acts):
    """CudaLBFGS::solve of the reference vs this library, both fp32, 30 iterations, m = 10"""
    B, iters = 1000, 30
    X, T = P.synthetic_mnist(B)
    dx, dt = upload(X), upload(T)
    ref = rc.RefCudaNet(dims, [ACT[a] for a in acts])
    ref.bind_params(123)
    w = ref.get_params()
    r = ref.solve("lbfgs", dx.data(), dt.data(), B, iters, tol=0.0, memory=10)
    net = make_gpu_net(handle, dims, acts, w)
    s = P.CudaLBFGS(handle)
    s.setMemory(10); s.setMaxIterations(iters); s.setTolerance(0.0)
    rec = P.IterationRecorder(); rec.init(iters); s.setRecorder(rec)
    s.solve(net.params_size(), net.params_data(), dx, dt, B, net)
    loss, gn, _ = rec.copy_to_host()
    assert r["iters"] == s.iterations() == iters
    for k in range(iters):  # two fp32 trajectories: agreement degrades geometrically with the iteration count
        assert abs(loss[k] - r["loss"][k]) <= 5e-5 * (1.6 ** min(k, 20)) * abs(r["loss"][k]), (k, loss[k], r["loss"][k])


def test_gd_sgd_vs_reference_cuda(handle):
    dims, acts, B = NETS[0][0], NETS[0][1], 1000
    X, T = P.synthetic_mnist(B)
    dx, dt = upload(X), upload(T)
    ref = rc.RefCudaNet(dims, [ACT[a] for a in acts])
    ref.bind_params(123)
    w = ref.get_params()
    r = ref.solve("gd", dx.data(), dt.data(), B, 20, tol=0.0, lr=0.05, momentum=0.9)
    net = make_gpu_net(handle, dims, acts, w)
    s = P.CudaGD(handle)
    s.setLearningRate(0.05); s.setMomentum(0.9); s.setMaxIterations(20); s.setTolerance(0.0)
    rec = P.IterationRecorder(); rec.init(20); s.setRecorder(rec)
    s.solve(net.params_size(), net.params_data(), dx, dt, B, net)
    loss, _, _ = rec.copy_to_host()
    assert np.allclose(loss, r["loss"], rtol=1e-4)
    assert rel_l2(net.get_params(), ref.get_params()) <= 1e-4

    ref.set_params(w)
    r = ref.solve("sgd", dx.data(), dt.data(), B, 4, tol=0.0, lr=0.02, momentum=0.9, sgd_batch=96, decay_rate=0.5, decay_step=2)
    net.set_params(w)
    s = P.CudaSGD(handle)
    s.setLearningRate(0.02); s.setMomentum(0.9); s.setBatchSize(96); s.setLearningRateDecay(0.5, 2)
    s.setMaxIterations(4); s.setTolerance(0.0); s.setDimensions(784, 10)
    rec = P.IterationRecorder(); rec.init(5); s.setRecorder(rec)
    s.solve(net.params_size(), net.params_data(), dx, dt, B, net)
    loss, _, _ = rec.copy_to_host()
    assert loss.size == r["loss"].size == 5
    assert np.allclose(loss, r["loss"], rtol=2e-4)
    assert rel_l2(net.get_params(), ref.get_params()) <= 2e-4
