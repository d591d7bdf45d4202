"""A small invocation of every kernel family for compute-sanitizer (racecheck / memcheck): the tcgen05 / TMA / mbarrier pipelines
(fwd16 with its three epilogues, dw16 plain and folded, the generic gemm_tc roles), the one-pass last layer, the fused direction
kernel with its grid barrier, finalize, the pair network of S-LBFGS. Sizes are tiny: the tools slow kernels down 10-100x.
usage (GPU box): compute-sanitizer --tool racecheck python tools/sanitize_case.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lbfgs_ffnn_b200 as P
from helpers import make_gpu_net, upload

h = P.CublasHandle(0)
B = 300
X, T = P.synthetic_mnist(B)
dx, dt = upload(X), upload(T)
for dims, acts in (([784, 128, 64, 10], ["relu", "relu", "linear"]), ([784, 128, 10], ["relu", "linear"])):
    for prec in ("tf32x3", "fp32"):
        net = make_gpu_net(h, dims, acts, None, precision=prec)
        if prec != "fp32":
            assert net.quantize_input(dx, B)
        loss = net.compute_loss_and_grad(dx, dt, B)
        s = P.CudaLBFGS(h)
        s.setMemory(4); s.setMaxIterations(4); s.setTolerance(0.0)
        s.solve(net.params_size(), net.params_data(), dx, dt, B, net)
        print(dims, prec, "loss", loss, "iterations", s.iterations(), flush=True)
        net.close()
net = make_gpu_net(h, [784, 128, 64, 10], ["relu", "relu", "linear"], None, precision="tf32x3")
s = P.CudaSLBFGS(h)
s.setMaxIterations(1); s.setTolerance(0.0); s.setStepSize(0.02); s.setBatchSize(100); s.setMemory(3); s.setUpdateInterval(1)
s.setHessianBatchSize(100)
rec = P.IterationRecorder(); rec.init(1); s.setRecorder(rec)
s.solve(net.params_size(), net.params_data(), dx, dt, B, net)
print("slbfgs epoch loss", rec.copy_to_host()[0], flush=True)
net.close(); h.close()
print("sanitize case done")
