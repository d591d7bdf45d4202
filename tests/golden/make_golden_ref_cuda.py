"""Golden vectors produced by the REFERENCE ITSELF: its CUDA backend (cuBLAS SGEMM + src/cuda/*.cuh, compiled
unmodified by oracle/ref_cuda/) run on a B200 on the synthetic inputs of lbfgs_ffnn_b200/data.py.

Run on the GPU box from the repo root:   python tests/golden/make_golden_ref_cuda.py gpurun_out/golden_ref_cuda.json
then copy the file to tests/golden/golden_ref_cuda.json. tests/test_oracle.py::test_oracle_vs_reference_cuda_goldens
checks the fp64 oracle against these numbers (fp32-level agreement), which pins the restatement of the objective and
of the CUDA-flavoured L-BFGS / GD / SGD to outputs of the reference's own code."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lbfgs_ffnn_b200 as P  # noqa: E402  (device buffers only)
from oracle import ref_cuda_binding as rc  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else "golden_ref_cuda.json"
ACT = {"linear": 0, "tanh": 1, "relu": 2, "sigmoid": 3}
G = {"objective": [], "lbfgs": [], "gd": [], "sgd": [], "init": []}
cases = [([784, 128, 10], ["relu", "linear"], 1000), ([784, 128, 64, 10], ["relu", "relu", "linear"], 1000),
         ([784, 32, 10], ["tanh", "sigmoid"], 257)]
for dims, acts, B in cases:
    X, T = P.synthetic_mnist(B, in_dim=dims[0], n_classes=dims[-1], seed=123)
    dx, dt = P.DeviceBuffer(), P.DeviceBuffer()
    dx.copy_from_host(X); dt.copy_from_host(T)
    net = rc.RefCudaNet(dims, [ACT[a] for a in acts])
    net.bind_params(123)  # the reference's own initialiser (src/cuda/network.cuh:37-59)
    w = net.get_params()
    G["init"].append(dict(dims=dims, acts=acts, seed=123, sum=float(w.astype(np.float64).sum()), sumsq=float((w.astype(np.float64) ** 2).sum()),
                          probe=[float(w[i]) for i in (0, 1, 1000, w.size // 2, w.size - 11, w.size - 1)]))
    loss, g = net.loss_grad(dx.data(), dt.data(), B)
    probe = [0, 1, 1000, w.size // 2, w.size - 11, w.size - 1]
    G["objective"].append(dict(dims=dims, acts=acts, batch=B, seed=123, loss=float(loss), gnorm=float(np.linalg.norm(g.astype(np.float64))),
                               probe_idx=probe, probe_grad=[float(g[i]) for i in probe]))
    for kind, kw in (("lbfgs", dict(memory=10)), ("gd", dict(lr=0.05, momentum=0.9)),
                     ("sgd", dict(lr=0.02, momentum=0.9, sgd_batch=96, decay_rate=0.5, decay_step=2))):
        net.bind_params(123)
        iters = 25 if kind != "sgd" else 4
        r = net.solve(kind, dx.data(), dt.data(), B, iters, tol=0.0, **kw)
        G[kind].append(dict(dims=dims, acts=acts, batch=B, seed=123, iters=int(r["iters"]), opts=kw,
                            loss=[float(v) for v in r["loss"]], gnorm=[float(v) for v in r["gnorm"]]))
    net.close()
with open(out_path, "w") as f:
    json.dump(G, f, indent=1)
print("written", out_path)
