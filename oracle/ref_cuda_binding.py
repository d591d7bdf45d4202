"""ctypes binding of oracle/_ref/libref_cuda.so — the reference's own CUDA backend compiled from /root/reference
(oracle/ref_cuda/). TEST INFRASTRUCTURE ONLY (tests/, tools/): never imported by the product package."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_cuda.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
        L.ref_net_create.restype = C.c_void_p
        L.ref_net_create.argtypes = [C.c_int, ip, ip]
        L.ref_net_destroy.argtypes = [C.c_void_p]
        L.ref_net_params_size.restype = C.c_long
        L.ref_net_params_size.argtypes = [C.c_void_p]
        L.ref_net_bind_params.argtypes = [C.c_void_p, C.c_uint]
        L.ref_net_set_params.argtypes = [C.c_void_p, fp]
        L.ref_net_get_params.argtypes = [C.c_void_p, fp]
        L.ref_net_loss_grad.restype = C.c_float
        L.ref_net_loss_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, fp]
        L.ref_net_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, fp]
        L.ref_solve.restype = C.c_int
        L.ref_solve.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float,
                                C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, fp, ip, fp]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class RefCudaNet:
    """cuda_mlp::CudaNetwork + the reference minimizers, unmodified."""

    def __init__(self, dims, acts):
        d = np.asarray(dims, dtype=np.int32)
        a = np.asarray(acts, dtype=np.int32)
        self.dims = list(dims)
        self.h = lib().ref_net_create(len(acts), d.ctypes.data_as(C.POINTER(C.c_int)), a.ctypes.data_as(C.POINTER(C.c_int)))
        self.n = lib().ref_net_params_size(self.h)

    def close(self):
        if self.h:
            lib().ref_net_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bind_params(self, seed):
        lib().ref_net_bind_params(self.h, seed)

    def set_params(self, w):
        w = np.ascontiguousarray(w, dtype=np.float32)
        lib().ref_net_set_params(self.h, _fp(w))

    def get_params(self):
        w = np.empty(self.n, dtype=np.float32)
        lib().ref_net_get_params(self.h, _fp(w))
        return w

    def loss_grad(self, x_dev, t_dev, batch):
        g = np.empty(self.n, dtype=np.float32)
        loss = lib().ref_net_loss_grad(self.h, C.c_void_p(x_dev), C.c_void_p(t_dev), batch, _fp(g))
        return loss, g

    def forward(self, x_dev, batch):
        out = np.empty(batch * self.dims[-1], dtype=np.float32)
        lib().ref_net_forward(self.h, C.c_void_p(x_dev), batch, _fp(out))
        return out.reshape(batch, self.dims[-1])

    def solve(self, kind, x_dev, t_dev, batch, max_iters, tol=0.0, memory=10, lr=0.01, momentum=0.9, sgd_batch=64,
              decay_rate=1.0, decay_step=0, record=True):
        cap = max_iters + 1 if record else 0
        l, g, t = (np.zeros(max(cap, 1), dtype=np.float32) for _ in range(3))
        size, ms = C.c_int(0), C.c_float(0)
        k = {"lbfgs": 0, "gd": 1, "sgd": 2}[kind]
        it = lib().ref_solve(self.h, k, C.c_void_p(x_dev), C.c_void_p(t_dev), batch, max_iters, tol, memory, lr, momentum,
                             sgd_batch, decay_rate, decay_step, self.dims[0], self.dims[-1], cap, _fp(l), _fp(g), _fp(t),
                             C.byref(size), C.byref(ms))
        s = size.value
        return dict(iters=it, loss=l[:s].copy(), gnorm=g[:s].copy(), time_ms=t[:s].copy(), total_ms=ms.value)
