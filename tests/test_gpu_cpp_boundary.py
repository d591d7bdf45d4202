"""The C++ side of the drop-in boundary, EXECUTED on the GPU: examples/main_gpu_synthetic.cpp is the reference's runner
(tests/mnist/main-gpu.cpp) written against include/unified + include/cuda_mlp — the headers a reference maintainer would
compile against — linked with libb200lbfgs.so by the host compiler alone. It must run, and the `<name>_history.csv` files it
writes (the reference's schema, src/unified_optimization.hpp:446-465) must carry the same numbers as the Python mirror of the
same interface driving the same C ABI with the same seeds (6 significant digits: the CSV's precision)."""
import csv
import os
import shutil
import subprocess

import numpy as np
import pytest

import lbfgs_ffnn_b200 as P
from helpers import upload

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, ITERS = 2000, 20


def _csv(path):
    with open(path) as f:
        rows = list(csv.DictReader(f))
    assert rows and list(rows[0].keys()) == ["Iteration", "Loss", "GradNorm", "TimeMs"]
    return np.array([float(r["Loss"]) for r in rows]), np.array([float(r["GradNorm"]) for r in rows])


@pytest.fixture(scope="module")
def runner_output(tmp_path_factory):
    if shutil.which("g++") is None and not os.path.exists(os.path.join(ROOT, "examples", "main_gpu_synthetic")):
        pytest.skip("no host compiler and no prebuilt runner")
    if shutil.which("g++") is not None:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "-s"])
    cwd = tmp_path_factory.mktemp("runner")
    r = subprocess.run([os.path.join(ROOT, "examples", "main_gpu_synthetic"), str(N), str(ITERS)], cwd=cwd, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return cwd, r.stdout


def _python_mirror(handle, kind, prec="fp32"):
    X, T = P.synthetic_mnist(N)  # the same mt19937(123) stream as the runner's generator
    dx, dt = upload(X), upload(T)
    net = P.CudaNetwork(handle)
    net.addLayer(784, 128, "relu"); net.addLayer(128, 10, "linear")
    net.bindParams(123)
    net.set_precision(prec)
    iters = ITERS
    if kind == "gd":
        s = P.CudaGD(handle); s.setLearningRate(0.02); s.setMomentum(0.9)
    elif kind == "lbfgs":
        s = P.CudaLBFGS(handle); s.setMemory(10)
    elif kind == "sgd":
        iters = max(1, ITERS // 20)
        s = P.CudaSGD(handle); s.setLearningRate(0.01); s.setMomentum(0.0); s.setBatchSize(256); s.setLearningRateDecay(0.80, 40)
        s.setDimensions(784, 10)
    else:
        iters = max(1, ITERS // 20)
        s = P.CudaSLBFGS(handle); s.setStepSize(0.02); s.setBatchSize(1000); s.setMemory(10); s.setUpdateInterval(10)
        s.setHessianBatchSize(5000)
    s.setMaxIterations(iters); s.setTolerance(1e-4 if kind == "slbfgs" else 1e-3)
    rec = P.IterationRecorder(); rec.init(iters + 1); s.setRecorder(rec)
    s.solve(net.params_size(), net.params_data(), dx, dt, N, net)
    loss, gn, _ = rec.copy_to_host()
    return loss.astype(np.float64), gn.astype(np.float64)


@pytest.mark.parametrize("name,kind,prec", [("SYN_GD", "gd", "fp32"), ("SYN_LBFGS_m10", "lbfgs", "fp32"),
                                            ("SYN_LBFGS_m10_tf32x3", "lbfgs", "tf32x3"), ("SYN_SGD", "sgd", "fp32"),
                                            ("SYN_SLBFGS", "slbfgs", "fp32")])
def test_cpp_runner_history_matches_python_mirror(handle, runner_output, name, kind, prec):
    cwd, stdout = runner_output
    assert f">>> Running CUDA Experiment: {name}" in stdout
    loss_c, gn_c = _csv(os.path.join(cwd, f"{name}_history.csv"))
    loss_p, gn_p = _python_mirror(handle, kind, prec)
    assert len(loss_c) == len(loss_p) and len(loss_c) >= 1
    # same kernels, same order, deterministic reductions: equal to the CSV's 6 digits
    np.testing.assert_allclose(loss_c, loss_p, rtol=2e-5)
    np.testing.assert_allclose(gn_c, gn_p, rtol=2e-5)


def test_cpp_runner_reports_train_and_test_metrics(runner_output):
    _, stdout = runner_output
    assert stdout.count("Training Results: MSE=") == 5 and stdout.count("Test Results: MSE=") == 5
    assert "Data Uploaded to GPU. Train: 2000 samples." in stdout
