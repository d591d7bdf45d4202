"""Generates tests/golden/golden_small.json from the CPU oracle (oracle/oracle.cpp).

The reference's CPU path cannot be built in this container (Eigen is absent) and its own
tests hold no vectors for the MLP path, so these fixtures are regression anchors of the oracle — which is
itself pinned by the reference's analytic KATs and an independent NumPy restatement (tests/test_oracle.py).
Run from the repo root:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402
import lbfgs_ffnn_b200 as P  # noqa: E402

out = {"objective": [], "lbfgs": []}
for dims, acts, B in [([784, 128, 10], [2, 0], 64), ([784, 128, 64, 10], [2, 2, 0], 100), ([784, 16, 10], [1, 3], 33)]:
    net = ob.OracleNet(dims, acts)
    w = net.init_params_cpu_rule(123).astype(np.float32)
    X, T = P.synthetic_mnist(B, in_dim=dims[0], n_classes=dims[-1], seed=123)
    loss, g = net.loss_grad(w, X, T)
    probe = [0, 1, 1000, w.size // 2, w.size - 11, w.size - 1]
    out["objective"].append(dict(dims=dims, acts=acts, batch=B, seed=123, data_seed=123, loss=loss,
                                 gnorm=float(np.linalg.norm(g)), probe_idx=probe, probe_grad=[float(g[i]) for i in probe]))
for policy in ("cpu", "cuda"):
    dims, acts, B = [784, 128, 10], [2, 0], 1000  # BASELINE.json configs[0]: m=10, 1000 samples
    net = ob.OracleNet(dims, acts)
    w = net.init_params_cpu_rule(123).astype(np.float32)
    X, T = P.synthetic_mnist(B, seed=123)
    r = net.lbfgs(w, X, T, m=10, max_iters=25, tol=0.0, policy=policy)
    out["lbfgs"].append(dict(dims=dims, acts=acts, batch=B, seed=123, data_seed=123, m=10, policy=policy,
                             loss=[float(v) for v in r["loss"]], gnorm=[float(v) for v in r["gnorm"]],
                             alpha=[float(v) for v in r["alpha"]]))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.json"), "w") as f:
    json.dump(out, f, indent=1)
print("written")
