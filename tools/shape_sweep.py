"""Parity of one loss+gradient evaluation against the fp64 oracle over network shapes outside the test list (GPU box).
usage: python tools/shape_sweep.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lbfgs_ffnn_b200 as P
from oracle import binding as ob
from helpers import make_gpu_net, upload

def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
h = P.CublasHandle(0)
shapes = [[784, 64, 32, 10], [784, 128, 128, 10], [784, 128, 64, 32, 10], [784, 64, 64, 64, 5], [784, 128, 32, 12],
          [784, 128, 96, 10], [784, 128, 64, 64, 64, 10], [256, 128, 64, 10], [784, 128, 256, 10]]
worst = 0.0
for dims in shapes:
    acts = ["relu"] * (len(dims) - 2) + ["linear"]
    for B in (33, 1000, 5000):
        X, _ = P.synthetic_mnist(B, seed=5)
        X = np.ascontiguousarray(X[:, :dims[0]])
        rs = np.random.RandomState(1)
        T = np.zeros((B, dims[-1]), dtype=np.float32); T[np.arange(B), rs.randint(0, dims[-1], B)] = 1
        onet = ob.OracleNet(dims, acts)
        w = onet.init_params_cuda_rule(123).astype(np.float32)
        lo, go = onet.loss_grad(w, X, T)
        for prec in ("tf32x3", "fp32"):
            net = make_gpu_net(h, dims, acts, w, precision=prec)
            dx, dt = upload(X), upload(T)
            q = net.quantize_input(dx, B) if prec != "fp32" else None
            loss = net.compute_loss_and_grad(dx, dt, B)
            g = net.get_grads()
            e = (abs(loss - lo) / abs(lo), rel(g, go))
            worst = max(worst, e[1])
            flag = "" if e[0] < 2e-5 and e[1] < 5e-5 else "  <-- CHECK"
            print(json.dumps({"net": "-".join(map(str, dims)), "B": B, "prec": prec, "quantized": q, "loss_err": f"{e[0]:.1e}", "grad_err": f"{e[1]:.1e}"}) + flag, flush=True)
print("worst grad rel-L2", worst)
