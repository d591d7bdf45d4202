"""Cases shared by tests/golden/make_golden_ref_cpu.py (which runs the reference's own CPU code) and
tests/test_oracle_vs_reference_cpu.py (which holds oracle/oracle.cpp to its outputs)."""
import numpy as np

import lbfgs_ffnn_b200 as P

ACTS = {(784, 128, 10): ["relu", "linear"], (784, 128, 64, 10): ["relu", "relu", "linear"],
        (20, 16, 8, 4): ["tanh", "sigmoid", "linear"], (784, 256, 128, 64, 10): ["relu", "relu", "relu", "linear"]}

CPU_CASES = [
    dict(dims=[20, 16, 8, 4], N=300, m=10, lbfgs_iters=40, gd_iters=30, gd_lr=0.1, sgd_epochs=5, sgd_lr=0.05, sgd_batch=32,
         slbfgs=[dict(batch=50, M=10, L=2, b_H=0, step=0.02, epochs=4), dict(batch=30, M=0, L=3, b_H=60, step=0.05, epochs=3),
                 dict(batch=25, M=3, L=2, b_H=100, step=0.03, epochs=5)]),
    dict(dims=[784, 128, 10], N=600, m=20, lbfgs_iters=25, gd_iters=10, gd_lr=0.05, sgd_epochs=3, sgd_lr=0.02, sgd_batch=128,
         slbfgs=[dict(batch=100, M=10, L=2, b_H=300, step=0.02, epochs=3)]),
    dict(dims=[784, 128, 64, 10], N=1000, m=10, lbfgs_iters=25, gd_iters=10, gd_lr=0.05, sgd_epochs=3, sgd_lr=0.05, sgd_batch=128,
         slbfgs=[dict(batch=100, M=10, L=3, b_H=0, step=0.02, epochs=3), dict(batch=100, M=0, L=10, b_H=0, step=0.02, epochs=2)]),
]


def case_problem(case):
    """(w0 float64, X, T): the initial parameters are the oracle-independent numpy draw below, so neither side's initialiser is
    involved; X, T are fp32-representable."""
    dims, N = case["dims"], case["N"]
    n = sum(a * b + b for a, b in zip(dims[:-1], dims[1:]))
    rs = np.random.RandomState(20240 + len(dims))
    w0 = np.concatenate([np.concatenate([(rs.standard_normal(a * b) * np.sqrt(2.0 / a)), rs.standard_normal(b) * 0.01])
                         for a, b in zip(dims[:-1], dims[1:])]).astype(np.float32).astype(np.float64)
    assert w0.size == n
    if dims[0] == 784:
        X, T = P.synthetic_mnist(N, seed=123)
    else:
        rs = np.random.RandomState(7)
        X = rs.rand(N, dims[0]).astype(np.float32)
        T = np.zeros((N, dims[-1]), dtype=np.float32)
        T[np.arange(N), rs.randint(0, dims[-1], N)] = 1
    return w0, X, T


def digest(v):
    """a vector as a few full-precision numbers: enough to pin it, small enough to commit"""
    v = np.asarray(v, dtype=np.float64)
    idx = sorted(set([0, 1, 2, v.size // 3, v.size // 2, v.size - 3, v.size - 2, v.size - 1]))
    return dict(n=int(v.size), norm=float(np.linalg.norm(v)), sum=float(v.sum()), idx=idx, val=[float(v[i]) for i in idx],
                wsum=float(np.dot(v, np.cos(np.arange(v.size) * 0.37))))


def digest_close(v, d, rtol):
    """max relative deviation of vector v from a stored digest (relative to the vector's norm for the probes)"""
    g = digest(v)
    assert g["n"] == d["n"] and g["idx"] == d["idx"]
    nrm = max(d["norm"], 1e-300)
    errs = [abs(g["norm"] - d["norm"]) / nrm, abs(g["sum"] - d["sum"]) / (nrm * np.sqrt(d["n"])),
            abs(g["wsum"] - d["wsum"]) / (nrm * np.sqrt(d["n"]))]
    errs += [abs(a - b) / max(abs(b), nrm / np.sqrt(d["n"])) for a, b in zip(g["val"], d["val"])]
    return max(errs), rtol
