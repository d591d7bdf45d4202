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


def relu_pattern_of(net, acts):
    """activation pattern of the hidden layers as the GPU evaluation took it (a > 0), one uint8 (batch, out) array per hidden
    layer; zeros for hidden layers that are not ReLU (the oracle ignores those)"""
    masks = []
    for l, a in enumerate(acts[:-1]):
        act = net.copy_activation_to_host(l)
        masks.append((act > 0).astype(np.uint8) if a in ("relu", 2) else np.zeros(act.shape, dtype=np.uint8))
    return masks


def oracle_on_gpu_pattern(onet, net, acts, w, X, T):
    """fp64 oracle loss / gradient with the ReLU pattern imposed that the last GPU evaluation of `net` took.

    Why: a pre-activation within fp32 rounding of zero lands on either side in fp32; at 60 000 samples x 192 hidden units a
    handful do, and each such unit moves the gradient by ~1e-5 of its norm (one sample's whole back-propagated signal through that
    unit). That is a property of evaluating max(z, 0) in fp32 at all — the reference's own CUDA backend has it too — not of the
    kernels' arithmetic; imposing the pattern removes it from the comparison and leaves the arithmetic error alone."""
    return onet.loss_grad_masked(w, X, T, relu_pattern_of(net, acts))


def upload(arr):
    buf = P.DeviceBuffer()
    buf.copy_from_host(np.ascontiguousarray(arr, dtype=np.float32))
    return buf
