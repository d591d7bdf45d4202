import os, sys, json
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
os.environ.setdefault("B200_WIDE16_MIN", "0")  # small shapes too
import lbfgs_ffnn_b200 as P
from oracle import binding as ob
from helpers import make_gpu_net, upload
def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
def blocks(dims):
    o, out = 0, []
    for i in range(len(dims) - 1):
        n = (dims[i] + 1) * dims[i + 1]; out.append((o, o + n)); o += n
    return out
h = P.CublasHandle(0)
for dims, B in (([320,256,256,256,12], 4100), ([320,256,256,256,12], 4096), ([320,256,256,256,12], 2100), ([320,256,256,12], 4100), ([784,256,256,256,10], 4100)):
    acts = ["relu"] * (len(dims) - 2) + ["linear"]
    X, _ = P.synthetic_mnist(B, seed=5); X = np.ascontiguousarray(X[:, :dims[0]])
    rs = np.random.RandomState(1)
    T = np.zeros((B, dims[-1]), dtype=np.float32); T[np.arange(B), rs.randint(0, dims[-1], B)] = 1
    onet = ob.OracleNet(dims, acts); w = onet.init_params_cuda_rule(123).astype(np.float32)
    lo, go = onet.loss_grad(w, X, T)
    for mask in (7, 6, 5, 3):
        os.environ["B200_TC_MASK"] = str(mask); P.api.reload_env()
        net = make_gpu_net(h, dims, acts, w, precision="tf32x3")
        dx, dt = upload(X), upload(T); net.quantize_input(dx, B)
        loss = net.compute_loss_and_grad(dx, dt, B); g = net.get_grads()
        print(dims, B, "mask", mask, "loss_err %.1e" % (abs(loss-lo)/abs(lo)), [f"{rel(g[a:b], go[a:b]):.1e}" for a, b in blocks(dims)], flush=True)
