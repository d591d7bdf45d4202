"""mid16 (hidden layer 1 of a three-layer net on the fp16 kernels) against the generic path and the fp64 oracle, plus per-kernel
CUDA-event times. usage: python tools/mid16_check.py [batches...]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lbfgs_ffnn_b200 as P
from oracle import binding as ob
from helpers import make_problem, make_gpu_net, upload, relu_pattern_of

def rel(a, b): return float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))
batches = [int(v) for v in sys.argv[1:]] or [128, 1000, 4099, 60000]
h = P.CublasHandle(0)
for dims, acts in (([784, 128, 64, 10], ["relu", "relu", "linear"]), ([784, 128, 128, 10], ["relu", "tanh", "linear"]),
                   ([784, 128, 64, 7], ["linear", "relu", "sigmoid"])):
    for B in batches:
        onet, w, X, T = make_problem(ob, dims, acts, B)
        if dims[-1] != 10:
            X, _ = P.synthetic_mnist(B)
            rs = np.random.RandomState(1); T = np.zeros((B, dims[-1]), np.float32); T[np.arange(B), rs.randint(0, dims[-1], B)] = 1
        dx, dt = upload(X), upload(T)
        row = {"net": "-".join(map(str, dims)), "acts": acts, "B": B}
        for mode in ("1", "0"):
            os.environ["B200_MID16"] = mode; P.api.reload_env()
            net = make_gpu_net(h, dims, acts, w, precision="tf32x3")
            assert net.quantize_input(dx, B)
            loss = net.compute_loss_and_grad(dx, dt, B)
            g = net.get_grads()
            pat = relu_pattern_of(net, acts)
            lm, gm = onet.loss_grad_masked(w, X, T, pat)
            net.forward_only(dx, B)
            out = net.copy_output_to_host().reshape(B, dims[-1])
            # per-layer gradient errors
            offs = np.cumsum([0] + [a * b + b for a, b in zip(dims[:-1], dims[1:])])
            per = [rel(g[offs[i]:offs[i + 1]], gm[offs[i]:offs[i + 1]]) for i in range(len(acts))]
            row["mid16" if mode == "1" else "generic"] = dict(loss=f"{abs(loss - lm) / abs(lm):.2e}", grad=f"{rel(g, gm):.2e}",
                                                             per_layer=[f"{v:.1e}" for v in per], fwd=f"{rel(out, onet.forward(w, X)):.2e}")
            if B >= 60000:
                for _ in range(3): net.compute_loss_and_grad(dx, dt, B)
                h.profile(True)
                for _ in range(10): net.compute_loss_and_grad(dx, dt, B)
                rep = h.profile_report(); h.profile(False)
                row["us_" + ("mid16" if mode == "1" else "generic")] = {k: round(1e3 * v[1] / v[0], 1) for k, v in rep.items()}
                row["eval_us_" + ("mid16" if mode == "1" else "generic")] = round(sum(1e3 * v[1] / v[0] for v in rep.values()), 1)
            net.close()
        print(json.dumps(row), flush=True)
