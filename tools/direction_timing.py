"""Phases of the fused direction kernel (B200_TC_TIMING) during a short L-BFGS run of the deep net, in steady state."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lbfgs_ffnn_b200 as P
from helpers import make_gpu_net, upload
B = 60000
h = P.CublasHandle(0)
X, T = P.synthetic_mnist(B)
dx, dt = upload(X), upload(T)
net = make_gpu_net(h, [784, 128, 64, 10], ["relu", "relu", "linear"], None, precision="tf32x3")
s = P.CudaLBFGS(h); s.setMemory(10); s.setMaxIterations(150); s.setTolerance(0.0)
s.solve(net.params_size(), net.params_data(), dx, dt, B, net)
os.environ["B200_TC_TIMING"] = "1"; P.api.reload_env()
s.setMaxIterations(14)
s.solve(net.params_size(), net.params_data(), dx, dt, B, net)
