"""GPU check of the persistent fp16 forward kernel (gemm_fwd16.cu) against the fp64 oracle and the generic tcgen05 kernel.
usage: python tools/fwd16_check.py [batches...]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lbfgs_ffnn_b200 as P
from oracle import binding as ob
from helpers import make_problem, make_gpu_net, upload

def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
batches = [int(a) for a in sys.argv[1:]] or [128, 1000, 4099, 60000]
h = P.CublasHandle(0)
for dims, acts in (([784, 128, 10], ["relu", "linear"]), ([784, 128, 64, 10], ["relu", "relu", "linear"]), ([784, 64, 10], ["tanh", "linear"])):
    for B in batches:
        onet, w, X, T = make_problem(ob, dims, acts, B)
        lo, go = onet.loss_grad(w, X, T)
        dx, dt = upload(X), upload(T)
        for prec in ("tf32x3", "tf32", "fp32"):
            row = {"net": "-".join(map(str, dims)), "B": B, "prec": prec}
            for mode in ("1", "0"):
                os.environ["B200_DW16"] = mode; P.api.reload_env()
                net = make_gpu_net(h, dims, acts, w, precision=prec)
                assert net.quantize_input(dx, B)
                loss = net.compute_loss_and_grad(dx, dt, B)
                g = net.get_grads()
                row["dw16" if mode == "1" else "nodw16"] = (f"{abs(loss-lo)/abs(lo):.2e}", f"{rel(g, go):.2e}")
            print(json.dumps(row), flush=True)
os.environ["B200_FWD16"] = "1"
B = 60000
X, T = P.synthetic_mnist(B)
dx, dt = upload(X), upload(T)
for dims, acts in (([784, 128, 10], ["relu", "linear"]), ([784, 128, 64, 10], ["relu", "relu", "linear"])):
    for prec in ("tf32x3", "tf32"):
        for mode in ("111", "110", "100"):
            os.environ["B200_FWD16"] = mode[0]; os.environ["B200_TAIL"] = mode[1]; os.environ["B200_DW16"] = mode[2]; P.api.reload_env()
            net = make_gpu_net(h, dims, acts, None, precision=prec)
            net.quantize_input(dx, B)
            for _ in range(3): net.compute_loss_and_grad(dx, dt, B)
            h.profile(True)
            for _ in range(20): net.loss_grad_async(dx, dt, B)
            rep = h.profile_report(); h.profile(False)
            print(json.dumps({"net": "-".join(map(str, dims)), "prec": prec, "fwd16,tail,dw16": mode,
                              "us": {k: round(1e3 * v[1] / v[0], 1) for k, v in rep.items()}}), flush=True)
