"""Wide-layer path (gemm_fwd16.cu, wide16) over shapes outside the test list: loss / gradient per layer block against the fp64 oracle
evaluated on the GPU's ReLU pattern. usage (GPU box): python tools/wide16_sweep.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
os.environ.setdefault("B200_WIDE16_MIN", "0")
import lbfgs_ffnn_b200 as P
from oracle import binding as ob
from helpers import make_gpu_net, relu_pattern_of, upload

def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
def blocks(dims):
    o, out = 0, []
    for i in range(len(dims) - 1):
        n = (dims[i] + 1) * dims[i + 1]; out.append((o, o + n)); o += n
    return out
h = P.CublasHandle(0)
cases = [([1000, 640, 10], ["relu", "linear"]), ([300, 256, 128, 10], ["relu", "relu", "linear"]), ([784, 1024, 512, 256, 10], ["relu", "sigmoid", "tanh", "linear"]),
         ([260, 384, 384, 7], ["relu", "relu", "linear"]), ([784, 256, 256, 64, 10], ["relu", "relu", "relu", "linear"]), ([512, 2048, 12], ["tanh", "linear"]),
         ([784, 128, 256, 256, 10], ["relu", "relu", "relu", "linear"])]
worst = 0.0
for dims, acts in cases:
    for B in (7, 129, 3000):
        rs = np.random.RandomState(3)
        if dims[0] <= 784:
            X, _ = P.synthetic_mnist(B, seed=5); X = np.ascontiguousarray(X[:, :dims[0]])
        else:
            X = rs.rand(B, dims[0]).astype(np.float32)
        T = np.zeros((B, dims[-1]), dtype=np.float32); T[np.arange(B), rs.randint(0, dims[-1], B)] = 1
        onet = ob.OracleNet(dims, acts); w = onet.init_params_cuda_rule(123).astype(np.float32)
        net = make_gpu_net(h, dims, acts, w, precision="tf32x3")
        dx, dt = upload(X), upload(T); net.quantize_input(dx, B)
        n0 = P.api.launch_count()
        loss = net.compute_loss_and_grad(dx, dt, B); g = net.get_grads()
        n1 = P.api.launch_count()
        lo, go = onet.loss_grad_masked(w, X, T, relu_pattern_of(net, acts))
        e = (abs(loss - lo) / abs(lo), rel(g, go)); worst = max(worst, e[1])
        flag = "" if e[0] < 5e-6 and e[1] < 1e-5 else "  <-- CHECK"
        print(json.dumps({"net": "-".join(map(str, dims)), "acts": "".join(a[0] for a in acts), "B": B, "launches": n1 - n0, "loss_err": f"{e[0]:.1e}", "grad_err": f"{e[1]:.1e}",
                          "per_layer": [f"{rel(g[a:b], go[a:b]):.1e}" for a, b in blocks(dims)]}) + flag, flush=True)
print("worst grad rel-L2", worst)
