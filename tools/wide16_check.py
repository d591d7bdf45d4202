"""Wide hidden layers on the fp16 pair kernels (B200_WIDE16, gemm_fwd16.cu): one loss + gradient evaluation against the fp64
oracle, with the path on and off, per layer block of the gradient; then event timings of one evaluation at a larger size.
usage (GPU box): python tools/wide16_check.py [timing-only]"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
os.environ.setdefault("B200_WIDE16_MIN", "0")  # small shapes too
import lbfgs_ffnn_b200 as P
from oracle import binding as ob
from helpers import make_gpu_net, upload


def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def layer_blocks(dims):
    o, out = 0, []
    for i in range(len(dims) - 1):
        n = (dims[i] + 1) * dims[i + 1]
        out.append((o, o + n)); o += n
    return out


def run(h, dims, acts, w, X, T, B, wide):
    os.environ["B200_WIDE16"] = "1" if wide else "0"
    P.api.reload_env()
    net = make_gpu_net(h, dims, acts, w, precision="tf32x3")
    dx, dt = upload(X), upload(T)
    net.quantize_input(dx, B)
    loss = net.compute_loss_and_grad(dx, dt, B)
    return loss, net.get_grads()


h = P.CublasHandle(0)
if len(sys.argv) < 2:
    worst = 0.0
    for dims, acts in (([784, 512, 256, 10], None), ([784, 256, 384, 10], ["tanh", "relu", "linear"]), ([320, 256, 256, 256, 12], None)):
        acts = acts or ["relu"] * (len(dims) - 2) + ["linear"]
        for B in (33, 1000, 4100):
            X, _ = P.synthetic_mnist(B, seed=5)
            X = np.ascontiguousarray(X[:, :dims[0]])
            rs = np.random.RandomState(1)
            T = np.zeros((B, dims[-1]), dtype=np.float32); T[np.arange(B), rs.randint(0, dims[-1], B)] = 1
            onet = ob.OracleNet(dims, acts)
            w = onet.init_params_cuda_rule(123).astype(np.float32)
            lo, go = onet.loss_grad(w, X, T)
            for wide in (1, 0):
                loss, g = run(h, dims, acts, w, X, T, B, wide)
                per = [f"{rel(g[a:b], go[a:b]):.1e}" for a, b in layer_blocks(dims)]
                e = (abs(loss - lo) / abs(lo), rel(g, go))
                if wide: worst = max(worst, e[1])
                flag = "" if e[0] < 1e-5 and e[1] < 1e-5 else "  <-- CHECK"
                print(json.dumps({"net": "-".join(map(str, dims)), "B": B, "wide16": wide, "loss_err": f"{e[0]:.1e}", "grad_err": f"{e[1]:.1e}",
                                  "per_layer": per}) + flag, flush=True)
    print("worst grad rel-L2 (wide16 on)", worst)

# timing of one evaluation, 784-4096-4096-10 at 16 384 samples (1/8 of the per-GPU share of configs[4])
dims, acts, B = [784, 4096, 4096, 10], ["relu", "relu", "linear"], 16384
X, _ = P.synthetic_mnist(B, seed=5)
rs = np.random.RandomState(1)
T = np.zeros((B, 10), dtype=np.float32); T[np.arange(B), rs.randint(0, 10, B)] = 1
onet = ob.OracleNet(dims, acts)
w = onet.init_params_cuda_rule(123).astype(np.float32)
res = {}
for wide in (1, 0):
    os.environ["B200_WIDE16"] = "1" if wide else "0"
    P.api.reload_env()
    net = make_gpu_net(h, dims, acts, w, precision="tf32x3")
    dx, dt = upload(X), upload(T)
    net.quantize_input(dx, B)
    for _ in range(2): loss = net.compute_loss_and_grad(dx, dt, B)
    h.profile(True)
    for _ in range(3): loss = net.compute_loss_and_grad(dx, dt, B)
    prof = h.profile_report(); h.profile(False)
    res[wide] = (loss, net.get_grads())
    print(json.dumps({"net": "784-4096-4096-10", "B": B, "wide16": wide, "loss": loss,
                      "us": {k: round(1e3 * v[1] / max(v[0], 1), 1) for k, v in prof.items()}}), flush=True)
print("wide16 on vs off at 16384 samples: loss rel diff %.1e, grad rel-L2 %.1e" % (abs(res[1][0] - res[0][0]) / abs(res[0][0]), rel(res[1][1], res[0][1])))
