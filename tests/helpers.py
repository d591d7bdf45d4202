"""Shared builders for the parity tests: identical fp32-representable inputs for the CUDA path and the oracle."""
import numpy as np

import lbfgs_ffnn_b200 as P

ACTS = {"linear": 0, "tanh": 1, "relu": 2, "sigmoid": 3}


def make_problem(oracle, dims, acts, batch, seed=123, data_seed=123):
    """returns (oracle_net, params32, X32, T32): every array float32 (so both sides see identical values)."""
    net = oracle.OracleNet(dims, acts)
    params = net.init_params_cpu_rule(seed).astype(np.float32)  # CPU rule in double, cast to float
    if dims[0] == 784 and dims[-1] == 10:
        X, T = P.synthetic_mnist(batch, seed=data_seed)
    else:
        rs = np.random.RandomState(data_seed)
        X = rs.rand(batch, dims[0]).astype(np.float32)
        T = np.zeros((batch, dims[-1]), dtype=np.float32)
        T[np.arange(batch), rs.randint(0, dims[-1], batch)] = 1
    return net, params, X, T


def make_gpu_net(handle, dims, acts, params=None, precision="fp32"):
    net = P.CudaNetwork(handle)
    for i, a in enumerate(acts):
        net.addLayer(dims[i], dims[i + 1], a)
    net.bindParams(123)
    net.set_precision(precision)
    if params is not None:
        net.set_params(params)
    return net


def upload(arr):
    buf = P.DeviceBuffer()
    buf.copy_from_host(np.ascontiguousarray(arr, dtype=np.float32))
    return buf
