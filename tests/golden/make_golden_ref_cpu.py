"""Golden vectors produced by the REFERENCE'S OWN CPU PATH: /root/reference/src/unified_launcher.hpp and everything it includes
(network.hpp, layer.hpp, unified_optimization.hpp, minimizer/{lbfgs,full_batch_minimizer,ring_buffer,gd,s_gd,s_lbfgs}.hpp),
compiled unmodified by oracle/ref_cpu/ against an Eigen-API stand-in (Eigen is absent from this image) and driven through
UnifiedLauncher<CpuBackend>::train / run_full_batch_cpu.

Run in the authoring container (where /root/reference exists) from the repo root:
    make -C oracle/ref_cpu && python tests/golden/make_golden_ref_cpu.py
tests/test_oracle_vs_reference_cpu.py checks oracle/oracle.cpp (the fp64 restatement the GPU parity tests are held to) against
these numbers: the L-BFGS (weak Wolfe), GD, SGD (random mini-batches) and S-LBFGS restatements are thereby pinned to outputs of
the reference's own code. Inputs are regenerated from seeds (lbfgs_ffnn_b200/data.py, numpy RandomState), not stored."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_cpu_binding as rc  # noqa: E402
from golden_cases import CPU_CASES, case_problem, digest  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "golden_ref_cpu.json")
G = []
for case in CPU_CASES:
    dims, N = case["dims"], case["N"]
    net = rc.RefCpuNet(dims)
    w_bind = net.bind_params(123)
    w0, X, T = case_problem(case)
    rec = dict(case=case, bind_params=digest(w_bind))
    loss, g = net.loss_grad(w0, X, T)
    rec["objective"] = dict(loss=loss, grad=digest(g))
    r = net.full_batch("lbfgs", w0, X, T, max_iters=case["lbfgs_iters"], tolerance=0.0, m_param=case["m"])
    rec["lbfgs"] = dict(iters=r["iters"], loss=list(map(float, r["loss"])), gnorm=list(map(float, r["gnorm"])), params=digest(r["params"]))
    r = net.full_batch("gd", w0, X, T, max_iters=case["gd_iters"], tolerance=0.0, learning_rate=case["gd_lr"])
    rec["gd"] = dict(iters=r["iters"], loss=list(map(float, r["loss"])), gnorm=list(map(float, r["gnorm"])), params=digest(r["params"]))
    r = net.train("sgd", w0, X, T, max_iters=case["sgd_epochs"], learning_rate=case["sgd_lr"], batch_size=case["sgd_batch"], log_interval=1)
    rec["sgd"] = dict(loss_csv=list(map(float, r["loss"])), gnorm_csv=list(map(float, r["gnorm"])), params=digest(r["params"]))
    rec["slbfgs"] = []
    for sl in case["slbfgs"]:
        # one run per epoch count: the strategy's recorder only reaches a 6-digit CSV, the parameters come back in full
        per_epoch = []
        for ep in range(1, sl["epochs"] + 1):
            r = net.train("slbfgs", w0, X, T, max_iters=ep, tolerance=0.0, learning_rate=sl["step"], batch_size=sl["batch"],
                          m_param=sl["M"], L_param=sl["L"], b_H_param=sl["b_H"], log_interval=1)
            per_epoch.append(digest(r["params"]))
        rec["slbfgs"].append(dict(opts=sl, loss_csv=list(map(float, r["loss"])), gnorm_csv=list(map(float, r["gnorm"])),
                                  params_after_epoch=per_epoch))
    G.append(rec)
with open(out_path, "w") as f:
    json.dump(dict(source="reference CPU sources (unmodified) + Eigen-API stand-in, oracle/ref_cpu", cases=G), f, indent=1)
print("written", out_path)
