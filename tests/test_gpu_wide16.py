"""Wide hidden layers (hundreds to thousands of columns: BASELINE configs[4], 784-4096-4096-10) on the fp16 pair kernels.

Kernels under test (reached through the C ABI; gemm_fwd16.cu, "wide16"):
  wide_amax_kernel / wide_split_kernel   an fp32 matrix -> the scaled fp16 pair [rows][hi | lo] and its transpose (+ row of ones)
  fwd16_kernel<..., EPI_WIDE, ..., WIDE> forward, dX and dW of a wide layer: CTA pairs (cta_group::2), M = 256, N = 256, K = 16 MMAs,
                                         the contraction accumulated in TMEM chunks that the epilogue warps sum in fp32 registers
They replace cublasSgemm + the element-wise kernels of CudaDenseLayer (src/cuda/layer.cuh:48-58 forward, :81-86 dW + db, :89-103 dX)
for layers with >= 256 inputs and a multiple of 128 (>= 256) outputs in the fp32-accurate mode; B200_WIDE16=0 runs the generic 3xTF32
kernels instead, and the two are compared with each other and with the fp64 oracle.

Stated tolerances (fp32-accurate mode): loss 5e-6, gradient 1e-5 relative L2 against the fp64 oracle evaluated on the ReLU pattern the
GPU took (helpers.relu_pattern_of: a pre-activation within fp32 rounding of zero lands on either side, in any fp32 evaluation)."""
import os

import numpy as np
import pytest

import lbfgs_ffnn_b200 as P
from conftest import rel_l2
from helpers import make_gpu_net, relu_pattern_of, upload

pytestmark = pytest.mark.gpu


def _problem(oracle, dims, acts, batch, seed=5):
    X, _ = P.synthetic_mnist(batch, seed=seed)
    X = np.ascontiguousarray(X[:, :dims[0]])
    rs = np.random.RandomState(1)
    T = np.zeros((batch, dims[-1]), dtype=np.float32)
    T[np.arange(batch), rs.randint(0, dims[-1], batch)] = 1
    onet = oracle.OracleNet(dims, acts)
    w = onet.init_params_cuda_rule(123).astype(np.float32)
    return onet, w, X, T


def _eval(handle, dims, acts, w, X, T, env=None, launches=False):
    keys = ("B200_WIDE16", "B200_WIDE_CHUNK", "B200_WIDE16_MIN")
    saved = {k: os.environ.get(k) for k in keys}
    try:
        for k in keys:
            os.environ.pop(k, None)
        os.environ["B200_WIDE16_MIN"] = "0"  # (by default GEMMs below 2^30 multiply-adds stay on the generic kernels)
        for k, v in (env or {}).items():
            os.environ[k] = v
        P.api.reload_env()
        B = X.shape[0]
        net = make_gpu_net(handle, dims, acts, w, precision="tf32x3")
        dx, dt = upload(X), upload(T)
        net.quantize_input(dx, B)
        n0 = P.api.launch_count()
        loss = net.compute_loss_and_grad(dx, dt, B)
        n1 = P.api.launch_count()
        g = net.get_grads()
        pattern = relu_pattern_of(net, acts)
        net.forward_only(dx, B)
        out = net.copy_output_to_host().reshape(B, dims[-1])
        return (loss, g, out, pattern, n1 - n0) if launches else (loss, g, out, pattern)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        P.api.reload_env()


NETS = [
    ([784, 512, 256, 10], ["relu", "relu", "linear"]),
    ([784, 256, 384, 10], ["tanh", "relu", "linear"]),       # dX through a tanh layer: act' from the fp32 activations
    ([320, 256, 256, 256, 12], ["relu", "relu", "relu", "linear"]),
    ([784, 384, 10], ["relu", "linear"]),                    # one wide layer (forward and dW only)
]


@pytest.mark.parametrize("which", range(len(NETS)))
@pytest.mark.parametrize("batch", [1, 33, 1000, 4100])
def test_wide_layers_match_the_oracle(handle, oracle, which, batch):
    """ragged sizes: 1 and 33 samples (one tile, the pair's second tile wholly past the batch), 4100 (an odd tile count and a last K
    block of the dW contraction with 4 valid samples)"""
    dims, acts = NETS[which]
    onet, w, X, T = _problem(oracle, dims, acts, batch)
    loss, g, out, pattern, launches = _eval(handle, dims, acts, w, X, T, launches=True)
    lo, go = onet.loss_grad_masked(w, X, T, pattern)
    assert abs(loss - lo) <= 5e-6 * abs(lo), (dims, batch, loss, lo)
    assert rel_l2(g, go) <= 1e-5, (dims, batch, rel_l2(g, go))
    # every wide layer block of the gradient on its own (a wrong bias row would hide in the norm of the whole)
    o = 0
    for l in range(len(dims) - 1):
        K, N = dims[l], dims[l + 1]
        assert rel_l2(g[o:o + K * N], go[o:o + K * N]) <= 2e-5, (dims, batch, "dW", l)
        assert rel_l2(g[o + K * N:o + (K + 1) * N], go[o + K * N:o + (K + 1) * N]) <= 2e-5, (dims, batch, "db", l)
        o += (K + 1) * N
    # the path under test did run: the generic kernels need fewer launches (no operand splits)
    _, g0, out0, _, launches0 = _eval(handle, dims, acts, w, X, T, env={"B200_WIDE16": "0"}, launches=True)
    assert launches > launches0, (launches, launches0)
    assert rel_l2(out, out0) <= 1e-5, (dims, batch, rel_l2(out, out0))


def test_accumulation_chunks_bound_the_truncation_bias(handle, oracle):
    """the tensor core's fp32 accumulate truncates; with the contraction accumulated in one piece (a chunk longer than K) the loss of a
    4096-deep contraction carries a bias several times the chunked form's"""
    dims, acts = [320, 2048, 256, 10], ["relu", "relu", "linear"]
    onet, w, X, T = _problem(oracle, dims, acts, 512)
    errs = {}
    for chunk in ("4", "1000"):
        loss, g, _, pattern = _eval(handle, dims, acts, w, X, T, env={"B200_WIDE_CHUNK": chunk})
        lo, go = onet.loss_grad_masked(w, X, T, pattern)
        errs[chunk] = (abs(loss - lo) / abs(lo), rel_l2(g, go))
    assert errs["4"][0] <= 5e-6 and errs["4"][1] <= 1e-5, errs
    assert errs["1000"][1] <= 1e-4, errs          # still a correct GEMM
    assert errs["4"][0] <= errs["1000"][0], errs  # and the chunks do what they are for


def test_wide_evaluation_is_bit_reproducible(handle, oracle):
    """no atomics and a fixed reduction order anywhere on the path (per-CTA maxima, one slice per dW): the same evaluation gives the same bits"""
    dims, acts = NETS[0]
    onet, w, X, T = _problem(oracle, dims, acts, 2500)
    ref = None
    for _ in range(4):
        loss, g, _, _ = _eval(handle, dims, acts, w, X, T)
        if ref is None:
            ref = (loss, g.copy())
        assert loss == ref[0] and np.array_equal(g, ref[1])


def test_small_gemms_stay_on_the_generic_kernels(handle, oracle):
    """the default threshold (B200_WIDE16_MIN = 2^30 multiply-adds): a 1000-sample evaluation of a 256-wide net launches exactly what
    it launches with the wide path switched off"""
    dims, acts = NETS[0]
    onet, w, X, T = _problem(oracle, dims, acts, 1000)
    a = _eval(handle, dims, acts, w, X, T, env={"B200_WIDE16_MIN": str(1 << 30)}, launches=True)
    b = _eval(handle, dims, acts, w, X, T, env={"B200_WIDE16": "0"}, launches=True)
    assert a[4] == b[4] and np.array_equal(a[1], b[1])


def test_lbfgs_on_a_wide_net_follows_the_oracle(handle, oracle):
    """five L-BFGS iterations (reference CUDA policy: Armijo) on a wide net: the loss trajectory of the fp64 oracle within 1e-3"""
    dims, acts = [784, 512, 256, 10], ["relu", "relu", "linear"]
    onet, w, X, T = _problem(oracle, dims, acts, 2000)
    os.environ["B200_WIDE16_MIN"] = "0"
    P.api.reload_env()
    try:
        net = make_gpu_net(handle, dims, acts, w, precision="tf32x3")
        dx, dt = upload(X), upload(T)
        net.quantize_input(dx, 2000)
        s = P.CudaLBFGS(handle)
        s.setMemory(10); s.setMaxIterations(5); s.setTolerance(0.0)
        rec = P.IterationRecorder(); rec.init(5); s.setRecorder(rec)
        n0 = P.api.launch_count()
        s.solve(net.params_size(), net.params_data(), dx, dt, 2000, net)
        launches = P.api.launch_count() - n0
        l, _, _ = rec.copy_to_host()
    finally:
        os.environ.pop("B200_WIDE16_MIN", None)
        P.api.reload_env()
    ref = onet.lbfgs(w, X, T, m=10, max_iters=5, tol=0.0, policy="cuda")
    assert np.allclose(l, ref["loss"], rtol=1e-3), (l, ref["loss"])
    assert launches > 0


def _float_problem(oracle, dims, acts, batch):
    """MNIST-shaped images plus sub-quantum noise: not 8-bit pixels, so the exact-fp16 kernels of layer 0 do not apply"""
    onet, w, X, T = _problem(oracle, dims, acts, batch)
    X = (X + np.float32(1e-3) * np.random.RandomState(7).rand(*X.shape).astype(np.float32)).astype(np.float32)
    return onet, w, X, T


@pytest.mark.parametrize("dims,acts", [([784, 128, 10], ["relu", "linear"]), ([784, 128, 64, 10], ["relu", "relu", "linear"])])
@pytest.mark.parametrize("batch", [33, 3000])
def test_float_input_takes_the_wide_kernels_for_a_128_wide_layer_0(handle, oracle, dims, acts, batch):
    """a 128-wide first layer on an input that is not 8-bit pixels: forward as a wide GEMM with one column tile, dW as a split-K wide
    GEMM (four output units: slices of the samples fill the machine, combined by the deterministic finalize kernel)"""
    onet, w, X, T = _float_problem(oracle, dims, acts, batch)
    loss, g, out, pattern, launches = _eval(handle, dims, acts, w, X, T, launches=True)
    lo, go = onet.loss_grad_masked(w, X, T, pattern)
    assert abs(loss - lo) <= 5e-6 * abs(lo), (dims, batch, loss, lo)
    assert rel_l2(g, go) <= 1e-5, (dims, batch, rel_l2(g, go))
    K, N = dims[0], dims[1]
    assert rel_l2(g[:K * N], go[:K * N]) <= 1e-5 and rel_l2(g[K * N:(K + 1) * N], go[K * N:(K + 1) * N]) <= 1e-5
    _, g0, _, _, launches0 = _eval(handle, dims, acts, w, X, T, env={"B200_WIDE16": "0"}, launches=True)
    assert launches > launches0 and rel_l2(g, g0) <= 1e-5


def test_held_float_input_is_split_once(handle, oracle):
    """b200_net_quantize_input on an input that is not 8-bit pixels: the caller holds it constant, so layer 0's operand split is made by
    the first evaluation only; an evaluation on other data in between invalidates it"""
    dims, acts = [784, 128, 10], ["relu", "linear"]
    onet, w, X, T = _float_problem(oracle, dims, acts, 2000)
    os.environ["B200_WIDE16_MIN"] = "0"
    P.api.reload_env()
    try:
        net = make_gpu_net(handle, dims, acts, w, precision="tf32x3")
        dx, dt = upload(X), upload(T)
        assert net.quantize_input(dx, 2000) is False
        counts, grads = [], []
        for _ in range(3):
            n0 = P.api.launch_count()
            net.compute_loss_and_grad(dx, dt, 2000)
            counts.append(P.api.launch_count() - n0)
            grads.append(net.get_grads().copy())
        assert counts[1] == counts[2] == counts[0] - 2, counts  # (maximum + split of X: first evaluation only)
        assert np.array_equal(grads[0], grads[1]) and np.array_equal(grads[0], grads[2])
        X2 = np.ascontiguousarray(X[::-1])
        dx2, dt2 = upload(X2), upload(np.ascontiguousarray(T[::-1]))
        net.compute_loss_and_grad(dx2, dt2, 2000)  # other data through the same buffers
        n0 = P.api.launch_count()
        net.compute_loss_and_grad(dx, dt, 2000)
        assert P.api.launch_count() - n0 == counts[0]  # split again
        assert np.array_equal(net.get_grads(), grads[0])
    finally:
        os.environ.pop("B200_WIDE16_MIN", None)
        P.api.reload_env()
