"""ctypes binding of oracle/_ref/libref_cpu.so — the reference's OWN CPU path (src/unified_launcher.hpp and everything it
includes) compiled unmodified from /root/reference against an Eigen-API stand-in (oracle/ref_cpu/).

TEST INFRASTRUCTURE ONLY (tests/, tools/, bench.py's reference arm): never imported by the product package.
The networks are a fixed list because the reference's layer sizes are template parameters (oracle/ref_cpu/ref_cpu.cpp)."""
import ctypes as C
import csv
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_cpu.so")
_lib = None

NETS = {  # id -> (dims, activations)
    0: ([784, 128, 10], ["relu", "linear"]),
    1: ([784, 128, 64, 10], ["relu", "relu", "linear"]),
    2: ([20, 16, 8, 4], ["tanh", "sigmoid", "linear"]),
    3: ([784, 256, 128, 64, 10], ["relu", "relu", "relu", "linear"]),
}
KIND = {"gd": 0, "lbfgs": 1, "sgd": 2, "slbfgs": 3}


def net_id(dims):
    for k, (d, _) in NETS.items():
        if list(d) == list(dims):
            return k
    raise KeyError(f"the reference CPU build has no instantiation of {dims}")


class Cfg(C.Structure):
    _fields_ = [("max_iters", C.c_int), ("tolerance", C.c_double), ("learning_rate", C.c_double), ("momentum", C.c_double),
                ("batch_size", C.c_int), ("m_param", C.c_int), ("L_param", C.c_int), ("b_H_param", C.c_int),
                ("log_interval", C.c_int)]


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.ref_cpu_num_threads.restype = C.c_int
        L.ref_cpu_set_num_threads.argtypes = [C.c_int]
        L.ref_cpu_params_size.restype = C.c_long
        L.ref_cpu_params_size.argtypes = [C.c_int]
        L.ref_cpu_bind_params.argtypes = [C.c_int, C.c_uint, dp]
        L.ref_cpu_loss_grad.argtypes = [C.c_int, dp, dp, dp, C.c_long, C.c_long, C.c_long, dp, dp]
        L.ref_cpu_full_batch.argtypes = [C.c_int, C.c_int, dp, dp, dp, C.c_long, C.c_long, C.c_long, C.POINTER(Cfg), dp, dp, dp,
                                         dp, ip, ip]
        L.ref_cpu_train.argtypes = [C.c_int, C.c_int, dp, dp, dp, C.c_long, C.c_long, C.c_long, C.POINTER(Cfg), C.c_char_p, dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def num_threads():
    return lib().ref_cpu_num_threads()


def set_num_threads(t):
    lib().ref_cpu_set_num_threads(int(t))


def _cfg(max_iters=100, tolerance=1e-4, learning_rate=0.01, momentum=0.0, batch_size=128, m_param=10, L_param=10, b_H_param=0,
         log_interval=1):
    return Cfg(max_iters, tolerance, learning_rate, momentum, batch_size, m_param, L_param, b_H_param, log_interval)


class RefCpuNet:
    """cpu_mlp::Network behind UnifiedLauncher<CpuBackend>, unmodified. X is [N][in], T is [N][out] (= the reference's
    column-major in x N / out x N matrices)."""

    def __init__(self, dims):
        self.id = net_id(dims)
        self.dims = list(dims)
        self.n = lib().ref_cpu_params_size(self.id)

    def bind_params(self, seed=123):
        w = np.empty(self.n)
        assert lib().ref_cpu_bind_params(self.id, seed, _d(w)) == 0
        return w

    def loss_grad(self, w, X, T):
        w, X, T = _f64(w), _f64(X), _f64(T)
        N = X.size // self.dims[0]
        loss = C.c_double(0)
        g = np.empty(self.n)
        assert lib().ref_cpu_loss_grad(self.id, _d(w), _d(X), _d(T), self.dims[0], self.dims[-1], N, C.byref(loss), _d(g)) == 0
        return loss.value, g

    def full_batch(self, kind, w, X, T, **kw):
        """cpu_mlp::LBFGS / GradientDescent + run_full_batch_cpu with the recorder kept (full-precision history)"""
        w, X, T = _f64(w), _f64(X), _f64(T)
        N = X.size // self.dims[0]
        cfg = _cfg(**kw)
        out = np.empty(self.n)
        hl, hg, hm = (np.zeros(cfg.max_iters) for _ in range(3))
        size, iters = C.c_int(0), C.c_int(0)
        assert lib().ref_cpu_full_batch(self.id, KIND[kind], _d(w), _d(X), _d(T), self.dims[0], self.dims[-1], N, C.byref(cfg),
                                        _d(out), _d(hl), _d(hg), _d(hm), C.byref(size), C.byref(iters)) == 0
        s = size.value
        return dict(params=out, iters=iters.value, loss=hl[:s].copy(), gnorm=hg[:s].copy(), ms=hm[:s].copy())

    def train(self, kind, w, X, T, **kw):
        """UnifiedLauncher<CpuBackend>::train with the Unified{GD,LBFGS,SGD,SLBFGS} strategy; history from the CSV it writes"""
        w, X, T = _f64(w), _f64(X), _f64(T)
        N = X.size // self.dims[0]
        cfg = _cfg(**kw)
        out = np.empty(self.n)
        with tempfile.TemporaryDirectory() as d:
            name = os.path.join(d, "run")
            assert lib().ref_cpu_train(self.id, KIND[kind], _d(w), _d(X), _d(T), self.dims[0], self.dims[-1], N, C.byref(cfg),
                                       name.encode(), _d(out)) == 0
            rows = []
            path = name + "_history.csv"
            if os.path.exists(path):
                with open(path) as f:
                    rows = list(csv.DictReader(f))
        return dict(params=out, iteration=np.array([int(r["Iteration"]) for r in rows]),
                    loss=np.array([float(r["Loss"]) for r in rows]), gnorm=np.array([float(r["GradNorm"]) for r in rows]),
                    ms=np.array([float(r["TimeMs"]) for r in rows]))
