"""ctypes binding of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: import from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package (lbfgs_ffnn_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

ACT = {"linear": 0, "tanh": 1, "relu": 2, "sigmoid": 3}


def build(force=False):
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


def _declare(L):
    dp, fp, ip, lp, up = (C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_long),
                          C.POINTER(C.c_uint32))
    L.oracle_num_threads.restype = C.c_int
    L.oracle_set_num_threads.argtypes = [C.c_int]
    L.oracle_net_create.restype = C.c_void_p
    L.oracle_net_create.argtypes = [C.c_int, ip, ip]
    L.oracle_net_destroy.argtypes = [C.c_void_p]
    L.oracle_net_params_size.restype = C.c_long
    L.oracle_net_params_size.argtypes = [C.c_void_p]
    L.oracle_init_params_cpu_rule.argtypes = [C.c_void_p, C.c_uint, dp]
    L.oracle_init_params_cuda_rule.argtypes = [C.c_void_p, C.c_uint, fp]
    L.oracle_loss.restype = C.c_double
    L.oracle_loss.argtypes = [C.c_void_p, dp, dp, dp, C.c_long]
    L.oracle_loss_grad.restype = C.c_double
    L.oracle_loss_grad.argtypes = [C.c_void_p, dp, dp, dp, C.c_long, dp]
    L.oracle_forward.argtypes = [C.c_void_p, dp, dp, C.c_long, dp]
    L.oracle_loss_grad_masked.restype = C.c_double
    L.oracle_loss_grad_masked.argtypes = [C.c_void_p, dp, dp, dp, C.c_long, C.POINTER(C.c_uint8), dp]
    L.oracle_slbfgs_batch_grad.restype = C.c_double
    L.oracle_slbfgs_batch_grad.argtypes = [C.c_void_p, dp, dp, dp, C.c_long, up, C.c_long, C.c_double, dp]
    L.oracle_direction.argtypes = [C.c_long, C.c_int, dp, dp, dp, dp, C.c_int, dp]
    L.oracle_lbfgs_mlp.restype = C.c_int
    L.oracle_lbfgs_mlp.argtypes = [C.c_void_p, dp, dp, dp, C.c_long, C.c_int, C.c_int, C.c_double, C.c_int, dp, dp, dp,
                                   dp, lp]
    L.oracle_lbfgs_analytic.restype = C.c_int
    L.oracle_lbfgs_analytic.argtypes = [C.c_int, C.c_long, dp, C.c_int, C.c_int, C.c_double, C.c_int, dp, dp]
    L.oracle_analytic_eval.restype = C.c_double
    L.oracle_analytic_eval.argtypes = [C.c_int, C.c_long, dp, dp]
    L.oracle_gd_mlp.restype = C.c_int
    L.oracle_gd_mlp.argtypes = [C.c_void_p, dp, dp, dp, C.c_long, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int,
                                dp, dp]
    L.oracle_sgd_mlp_cpu_policy.restype = C.c_int
    L.oracle_sgd_mlp_cpu_policy.argtypes = [C.c_void_p, dp, dp, dp, C.c_long, C.c_int, C.c_double, C.c_int, C.c_uint, dp, dp]
    L.oracle_sgd_mlp_cuda_policy.restype = C.c_int
    L.oracle_sgd_mlp_cuda_policy.argtypes = [C.c_void_p, dp, dp, dp, C.c_long, C.c_int, C.c_double, C.c_double,
                                             C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, dp, dp]
    L.oracle_slbfgs_mlp.restype = C.c_int
    L.oracle_slbfgs_mlp.argtypes = [C.c_void_p, dp, dp, dp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                    C.c_int, C.c_double, C.c_uint, dp, dp, up, C.c_long, ip, ip]
    L.oracle_sample_stream.argtypes = [C.c_uint, C.c_long, C.c_long, C.c_int, up]


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleNet:
    """cpu_mlp::Network restatement. dims = [in, h1, ..., out]; acts per layer."""

    def __init__(self, dims, acts):
        self.dims = list(dims)
        self.acts = [ACT[a] if isinstance(a, str) else int(a) for a in acts]
        d = np.asarray(self.dims, dtype=np.int32)
        a = np.asarray(self.acts, dtype=np.int32)
        self.h = lib().oracle_net_create(len(self.acts), _p(d, C.c_int), _p(a, C.c_int))
        self.n = lib().oracle_net_params_size(self.h)

    def __del__(self):
        try:
            lib().oracle_net_destroy(self.h)
        except Exception:
            pass

    def init_params_cpu_rule(self, seed=123):
        out = np.empty(self.n, dtype=np.float64)
        lib().oracle_init_params_cpu_rule(self.h, seed, _p(out, C.c_double))
        return out

    def init_params_cuda_rule(self, seed=123):
        out = np.empty(self.n, dtype=np.float32)
        lib().oracle_init_params_cuda_rule(self.h, seed, _p(out, C.c_float))
        return out

    def loss(self, params, X, T):
        params, X, T = _f64(params), _f64(X), _f64(T)
        B = X.size // self.dims[0]
        return lib().oracle_loss(self.h, _p(params, C.c_double), _p(X, C.c_double), _p(T, C.c_double), B)

    def loss_grad(self, params, X, T):
        params, X, T = _f64(params), _f64(X), _f64(T)
        B = X.size // self.dims[0]
        g = np.empty(self.n, dtype=np.float64)
        loss = lib().oracle_loss_grad(self.h, _p(params, C.c_double), _p(X, C.c_double), _p(T, C.c_double), B,
                                      _p(g, C.c_double))
        return loss, g

    def loss_grad_masked(self, params, X, T, masks):
        """loss / gradient with the ReLU pattern of the hidden layers imposed: masks = [uint8 (B, out_l) for each hidden layer]"""
        params, X, T = _f64(params), _f64(X), _f64(T)
        B = X.size // self.dims[0]
        flat = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.uint8).ravel() for m in masks]))
        assert flat.size == B * sum(self.dims[1:-1])
        g = np.empty(self.n, dtype=np.float64)
        loss = lib().oracle_loss_grad_masked(self.h, _p(params, C.c_double), _p(X, C.c_double), _p(T, C.c_double), B,
                                             _p(flat, C.c_uint8), _p(g, C.c_double))
        return loss, g

    def forward(self, params, X):
        params, X = _f64(params), _f64(X)
        B = X.size // self.dims[0]
        out = np.empty((B, self.dims[-1]), dtype=np.float64)
        lib().oracle_forward(self.h, _p(params, C.c_double), _p(X, C.c_double), B, _p(out, C.c_double))
        return out

    def slbfgs_batch_grad(self, params, X, T, idx, lam=1e-4):
        params, X, T = _f64(params), _f64(X), _f64(T)
        N = X.size // self.dims[0]
        g = np.empty(self.n, dtype=np.float64)
        if idx is None:
            ip, bs = None, N
        else:
            idx = np.ascontiguousarray(idx, dtype=np.uint32)
            ip, bs = _p(idx, C.c_uint32), idx.size
        f = lib().oracle_slbfgs_batch_grad(self.h, _p(params, C.c_double), _p(X, C.c_double), _p(T, C.c_double), N, ip,
                                           bs, lam, _p(g, C.c_double))
        return f, g

    def lbfgs(self, params, X, T, m=10, max_iters=100, tol=1e-4, policy="cpu"):
        """policy 'cpu' = reference CPU algorithm (weak Wolfe); 'cuda' = reference CUDA algorithm (Armijo), fp64."""
        x = _f64(params).copy()
        X, T = _f64(X), _f64(T)
        B = X.size // self.dims[0]
        hl, hg, hm, ha = (np.zeros(max_iters) for _ in range(4))
        counts = np.zeros(3, dtype=np.int64)
        it = lib().oracle_lbfgs_mlp(self.h, _p(x, C.c_double), _p(X, C.c_double), _p(T, C.c_double), B, m, max_iters,
                                    tol, 0 if policy == "cpu" else 1, _p(hl, C.c_double), _p(hg, C.c_double),
                                    _p(hm, C.c_double), _p(ha, C.c_double), _p(counts, C.c_long))
        return dict(params=x, iters=it, loss=hl[:it], gnorm=hg[:it], ms=hm[:it], alpha=ha[:it],
                    n_f=int(counts[1]), n_g=int(counts[2]))

    def gd(self, params, X, T, lr, momentum=0.0, max_iters=100, tol=1e-4, policy="cuda"):
        x = _f64(params).copy()
        X, T = _f64(X), _f64(T)
        B = X.size // self.dims[0]
        hl, hg = np.zeros(max_iters), np.zeros(max_iters)
        it = lib().oracle_gd_mlp(self.h, _p(x, C.c_double), _p(X, C.c_double), _p(T, C.c_double), B, lr, momentum,
                                 max_iters, tol, 0 if policy == "cpu" else 1, _p(hl, C.c_double), _p(hg, C.c_double))
        return dict(params=x, iters=it, loss=hl[:it], gnorm=hg[:it])

    def sgd_cuda_policy(self, params, X, T, batch_size, lr, momentum=0.0, decay_rate=1.0, decay_step=0, max_iters=10,
                        tol=0.0, record=True):
        x = _f64(params).copy()
        X, T = _f64(X), _f64(T)
        B = X.size // self.dims[0]
        hl, hg = np.zeros(max_iters + 1), np.zeros(max_iters + 1)
        it = lib().oracle_sgd_mlp_cuda_policy(self.h, _p(x, C.c_double), _p(X, C.c_double), _p(T, C.c_double), B,
                                              batch_size, lr, momentum, decay_rate, decay_step, max_iters, tol,
                                              1 if record else 0, _p(hl, C.c_double), _p(hg, C.c_double))
        return dict(params=x, iters=it, loss=hl[:it], gnorm=hg[:it])

    def sgd_cpu_policy(self, params, X, T, batch_size, lr, max_iters=10, seed=123):
        """reference CPU SGD (src/minimizer/s_gd.hpp:63-170): random mini-batches, no momentum"""
        x = _f64(params).copy()
        X, T = _f64(X), _f64(T)
        N = X.size // self.dims[0]
        hl, hg = np.zeros(max_iters), np.zeros(max_iters)
        it = lib().oracle_sgd_mlp_cpu_policy(self.h, _p(x, C.c_double), _p(X, C.c_double), _p(T, C.c_double), N, batch_size,
                                             lr, max_iters, seed, _p(hl, C.c_double), _p(hg, C.c_double))
        return dict(params=x, iters=it, loss=hl[:it], gnorm=hg[:it])

    def slbfgs(self, params, X, T, batch_size, M=10, L=10, b_H=0, step=0.02, max_iters=5, tol=1e-4, seed=123):
        x = _f64(params).copy()
        X, T = _f64(X), _f64(T)
        N = X.size // self.dims[0]
        hl, hg = np.zeros(max_iters), np.zeros(max_iters)
        m = max(1, N // batch_size)
        cap = m * batch_size
        tr = np.zeros(cap, dtype=np.uint32)
        ap = np.full(max_iters, -1, dtype=np.int32)
        pa = np.full(max_iters, -1, dtype=np.int32)
        it = lib().oracle_slbfgs_mlp(self.h, _p(x, C.c_double), _p(X, C.c_double), _p(T, C.c_double), N, batch_size, M,
                                     L, b_H, step, max_iters, tol, seed, _p(hl, C.c_double), _p(hg, C.c_double),
                                     _p(tr, C.c_uint32), cap, _p(ap, C.c_int), _p(pa, C.c_int))
        return dict(params=x, iters=it, loss=hl[:it], gnorm=hg[:it], first_epoch_idx=tr, anchor_picks=ap[:it],
                    pairs_after=pa[:it])


def direction(S, Y, rho, g, policy="cpu"):
    """Two-loop recursion. S, Y: (k, n) logical order (row 0 oldest)."""
    S, Y, rho, g = _f64(S), _f64(Y), _f64(rho), _f64(g)
    k = S.shape[0] if S.size else 0
    n = g.size
    out = np.empty(n, dtype=np.float64)
    pol = {"cpu": 0, "cuda": 1, "slbfgs": 2}[policy]
    lib().oracle_direction(n, k, _p(S, C.c_double), _p(Y, C.c_double), _p(rho, C.c_double), _p(g, C.c_double), pol,
                           _p(out, C.c_double))
    return out


def lbfgs_analytic(fn, x0, m=16, max_iters=4000, tol=1e-12, policy="cpu"):
    x = _f64(x0).copy()
    gn, fv = C.c_double(), C.c_double()
    fid = {"rosenbrock": 0, "ackley": 1, "rastrigin": 2}[fn]
    it = lib().oracle_lbfgs_analytic(fid, x.size, _p(x, C.c_double), m, max_iters, tol, 0 if policy == "cpu" else 1,
                                     C.byref(gn), C.byref(fv))
    return dict(x=x, iters=it, gnorm=gn.value, f=fv.value)


def analytic_eval(fn, x):
    x = _f64(x)
    g = np.empty_like(x)
    fid = {"rosenbrock": 0, "ackley": 1, "rastrigin": 2}[fn]
    f = lib().oracle_analytic_eval(fid, x.size, _p(x, C.c_double), _p(g, C.c_double))
    return f, g


def sample_stream(seed, N, b, count):
    out = np.empty(count * b, dtype=np.uint32)
    lib().oracle_sample_stream(seed, N, b, count, _p(out, C.c_uint32))
    return out.reshape(count, b)


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(t):
    lib().oracle_set_num_threads(int(t))
