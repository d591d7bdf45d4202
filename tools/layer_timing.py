"""Per-kernel event timing of one loss+gradient evaluation (library profiler scopes).
usage: python tools/layer_timing.py [dims like 784-128-64-10] [batch] [precision]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lbfgs_ffnn_b200 as P
from helpers import make_gpu_net, upload

dims = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "784-128-64-10").split("-")]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
prec = sys.argv[3] if len(sys.argv) > 3 else "tf32x3"
acts = ["relu"] * (len(dims) - 2) + ["linear"]
h = P.CublasHandle(0)
X, T = P.synthetic_mnist(B)
dx, dt = upload(X), upload(T)
net = make_gpu_net(h, dims, acts, None, precision=prec)
net.quantize_input(dx, B)
for _ in range(3): net.compute_loss_and_grad(dx, dt, B)
h.profile(True)
for _ in range(20): net.loss_grad_async(dx, dt, B)
rep = h.profile_report(); h.profile(False)
us = {k: round(1e3 * v[1] / v[0], 1) for k, v in rep.items()}
print(json.dumps({"net": "-".join(map(str, dims)), "B": B, "prec": prec, "us": us, "sum_us": round(sum(us.values()), 1)}))
