"""Host-side mirror of the reference's CUDA operator interface, over the C ABI (include/b200_lbfgs.h).

Same names, argument meaning and error behaviour as the reference classes so the parity tests read like the
reference's own call sites:

  CublasHandle          src/cuda/cublas_handle.cuh:22-39   (here: an opaque context, no cuBLAS — SURVEY.md D7)
  DeviceBuffer          src/cuda/device_buffer.cuh:7-96
  CudaNetwork           src/cuda/network.cuh:16-156
  IterationRecorder     src/iteration_recorder.hpp:81-146
  CudaMinimizerBase     src/cuda/minimizer_base.cuh:12-67
  CudaLBFGS/GD/SGD      src/cuda/lbfgs.cuh, gd.cuh, sgd.cuh
  CudaSLBFGS            new on the GPU; semantics of src/minimizer/s_lbfgs.hpp:165-290

The C++ twin of this file is include/cuda_mlp/*.cuh. Device pointers are plain integers.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ACT, LS, PREC, B200Error, check, lib

Linear, Tanh, ReLU, Sigmoid = 0, 1, 2, 3  # ActivationType, src/cuda/kernels.cuh:53-58


def _ptr(x):
    """device pointer of a DeviceBuffer / torch tensor / int"""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "data"):
        d = x.data
        return d() if callable(d) else d
    raise TypeError(f"cannot take a device pointer of {type(x)}")


class CublasHandle:
    """Opaque context shim (device, stream, optional NCCL communicator). Name kept for call-site compatibility."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().b200_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)

    def close(self):
        if self._h:
            lib().b200_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(lib().b200_ctx_synchronize(self._h))

    @property
    def stream(self):
        return lib().b200_ctx_stream(self._h) or 0

    def set_stream(self, cuda_stream):
        check(lib().b200_ctx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def profile(self, enable=True):
        check(lib().b200_ctx_profile(self._h, 1 if enable else 0))

    def profile_report(self):
        """{kernel class: (launches, total_ms)} measured with CUDA events around each launch"""
        import json
        buf = C.create_string_buffer(1 << 16)
        check(lib().b200_ctx_profile_report(self._h, buf, len(buf)))
        return {k: (int(v[0]), float(v[1])) for k, v in json.loads(buf.value.decode()).items()}

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        check(lib().b200_comm_unique_id(buf))
        return bytes(buf)

    def init_comm(self, unique_id, rank, world):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        check(lib().b200_ctx_init_comm(self._h, buf, int(rank), int(world)))

    @property
    def rank(self):
        return lib().b200_ctx_rank(self._h)

    @property
    def world(self):
        return lib().b200_ctx_world(self._h)

    def allreduce_f32(self, dev, n):
        check(lib().b200_ctx_allreduce_f32(self._h, C.c_void_p(_ptr(dev)), n))


class DeviceBuffer:
    """Move-only RAII device array (float32 unless dtype is given)."""

    def __init__(self, count=0, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self._p = C.c_void_p()
        self._n = 0
        if count:
            self.resize(count)

    def resize(self, count):  # device_buffer.cuh:52-62: realloc, contents not preserved
        if count == self._n:
            return
        self.free()
        if count:
            check(lib().b200_malloc(C.byref(self._p), count * self.dtype.itemsize))
        self._n = count

    def free(self):
        if self._p:
            lib().b200_free(self._p)
        self._p = C.c_void_p()
        self._n = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def data(self):
        return self._p.value or 0

    def size(self):
        return self._n

    def copy_from_host(self, host):
        a = np.ascontiguousarray(host, dtype=self.dtype).ravel()
        if a.size != self._n:
            self.resize(a.size)
        check(lib().b200_memcpy_h2d(self._p, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def copy_to_host(self, count=None):
        n = self._n if count is None else count
        out = np.empty(n, dtype=self.dtype)
        check(lib().b200_memcpy_d2h(out.ctypes.data_as(C.c_void_p), self._p, out.nbytes))
        return out


def _d2h(ptr, n, dtype=np.float32):
    out = np.empty(n, dtype=dtype)
    check(lib().b200_memcpy_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), out.nbytes))
    return out


def _h2d(ptr, host, dtype=np.float32):
    a = np.ascontiguousarray(host, dtype=dtype).ravel()
    check(lib().b200_memcpy_h2d(C.c_void_p(ptr), a.ctypes.data_as(C.c_void_p), a.nbytes))


class CudaNetwork:
    def __init__(self, handle):
        self.handle = handle
        self._layers = []
        self._h = C.c_void_p()

    def addLayer(self, in_dim, out_dim, act):
        if self._h:
            raise B200Error("addLayer after bindParams")
        self._layers.append((int(in_dim), int(out_dim), ACT[act] if isinstance(act, str) else int(act)))

    def _create(self):
        if self._h:
            return
        if not self._layers:
            raise B200Error("network has no layers")
        dims = [self._layers[0][0]] + [l[1] for l in self._layers]
        for (i, o, _), d in zip(self._layers, dims[:-1]):
            if i != d:
                raise B200Error("layer dimensions do not chain")
        acts = [l[2] for l in self._layers]
        cd = (C.c_int * len(dims))(*dims)
        ca = (C.c_int * len(acts))(*acts)
        check(lib().b200_net_create(self.handle._h, len(acts), cd, ca, C.byref(self._h)))
        self.dims, self.acts = dims, acts

    def bindParams(self, seed=123):
        self._create()
        check(lib().b200_net_bind_params(self._h, int(seed)))

    def close(self):
        if self._h:
            lib().b200_net_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def params_size(self):
        self._create()
        return lib().b200_net_params_size(self._h)

    def output_size(self):
        self._create()
        return lib().b200_net_output_size(self._h)

    def params_data(self):
        return lib().b200_net_params_data(self._h) or 0

    def grads_data(self):
        return lib().b200_net_grads_data(self._h) or 0

    def zeroGrads(self):
        check(lib().b200_net_zero_grads(self._h))

    def forward_only(self, input_dev, batch):
        check(lib().b200_net_forward(self._h, C.c_void_p(_ptr(input_dev)), int(batch)))

    def compute_loss_and_grad(self, input_dev, target_dev, batch):
        loss = C.c_float()
        check(lib().b200_net_loss_grad(self._h, C.c_void_p(_ptr(input_dev)), C.c_void_p(_ptr(target_dev)), int(batch),
                                       C.byref(loss)))
        return loss.value

    def loss_grad_async(self, input_dev, target_dev, batch, params_dev=None, grad_dev=None, loss_dev=None):
        check(lib().b200_net_loss_grad_async(self._h, C.c_void_p(_ptr(params_dev) or 0), C.c_void_p(_ptr(input_dev)),
                                             C.c_void_p(_ptr(target_dev)), int(batch), C.c_void_p(_ptr(grad_dev) or 0),
                                             C.c_void_p(_ptr(loss_dev) or 0)))

    def copy_output_to_host(self, count=None):
        n = self.output_size() * self.last_batch() if count is None else count
        out = np.empty(n, dtype=np.float32)
        check(lib().b200_net_copy_output_to_host(self._h, out.ctypes.data_as(C.c_void_p), n))
        return out

    def last_batch(self):
        return lib().b200_net_last_batch(self._h)

    def copy_activation_to_host(self, layer):
        """activations of `layer` (0 = first hidden layer) left by the last forward / loss_grad call, shape (batch, out)"""
        b, out = self.last_batch(), self.dims[layer + 1]
        a = np.empty(b * out, dtype=np.float32)
        check(lib().b200_net_copy_activation_to_host(self._h, int(layer), a.ctypes.data_as(C.c_void_p), a.size))
        return a.reshape(b, out)

    def evaluate(self, input_dev, target_dev, batch):
        mse, acc = C.c_double(), C.c_double()
        check(lib().b200_net_evaluate(self._h, C.c_void_p(_ptr(input_dev)), C.c_void_p(_ptr(target_dev)), int(batch),
                                      C.byref(mse), C.byref(acc)))
        return mse.value, acc.value

    # ---- additions over the reference -------------------------------------------------------------
    def set_precision(self, mode):
        check(lib().b200_net_set_precision(self._h, PREC[mode] if isinstance(mode, str) else int(mode)))

    def set_l2(self, lam):
        check(lib().b200_net_set_l2(self._h, float(lam)))

    def set_global_batch(self, batch_global):
        check(lib().b200_net_set_global_batch(self._h, int(batch_global)))

    def quantize_input(self, input_dev, batch):
        """cache a uint8 copy of an input that is exactly u/255 (see include/b200_lbfgs.h); returns True if used"""
        ok = C.c_int(0)
        check(lib().b200_net_quantize_input(self._h, C.c_void_p(_ptr(input_dev)), int(batch), C.byref(ok)))
        return bool(ok.value)

    def clear_input_cache(self):
        check(lib().b200_net_clear_input_cache(self._h))

    def set_params(self, host):
        """inject an explicit parameter vector (the two reference backends initialise differently)"""
        a = np.asarray(host)
        if a.size != self.params_size():
            raise B200Error("parameter vector has the wrong size")
        _h2d(self.params_data(), a)

    def get_params(self):
        return _d2h(self.params_data(), self.params_size())

    def get_grads(self):
        return _d2h(self.grads_data(), self.params_size())


class IterationRecorder:
    """IterationRecorder<CudaBackend>: loss / grad-norm / cumulative-ms per iteration."""

    def __init__(self):
        self.init(0)

    def init(self, capacity):
        self._cap = int(capacity)
        self._loss = np.zeros(max(self._cap, 1), dtype=np.float32)
        self._grad = np.zeros(max(self._cap, 1), dtype=np.float32)
        self._time = np.zeros(max(self._cap, 1), dtype=np.float32)
        self._size = 0
        self.evaluations = 0
        self.launches = 0

    def reset(self):
        self._size = 0

    def size(self):
        return self._size

    def copy_to_host(self):
        s = self._size
        return self._loss[:s].copy(), self._grad[:s].copy(), self._time[:s].copy()

    def _c_history(self):
        h = _lib.History()
        h.capacity = self._cap
        h.loss = self._loss.ctypes.data_as(C.POINTER(C.c_float))
        h.grad_norm = self._grad.ctypes.data_as(C.POINTER(C.c_float))
        h.time_ms = self._time.ctypes.data_as(C.POINTER(C.c_float))
        return h

    def _absorb(self, h):
        self._size = h.size
        self.evaluations = h.evaluations
        self.launches = h.launches


def write_cuda_history_csv(filename, recorder, log_interval):
    """src/unified_optimization.hpp:446-465 — same columns, same stride, default ostream float formatting (%g)."""
    if log_interval <= 0:
        return
    loss, grad, time_ms = recorder.copy_to_host()
    if loss.size == 0:
        return
    with open(filename, "w") as f:
        f.write("Iteration,Loss,GradNorm,TimeMs\n")
        for i in range(0, loss.size, max(1, log_interval)):
            f.write(f"{i},{float(loss[i]):g},{float(grad[i]):g},{float(time_ms[i]):g}\n")


class CudaMinimizerBase:
    def __init__(self, handle):
        self.handle = handle
        self.max_iters_ = 200  # minimizer_base.cuh:61-65
        self.max_line_iters_ = 20
        self.tol_ = 1e-6
        self.c1_ = 1e-4
        self.rho_ = 0.5
        self.recorder_ = None
        self.last_iterations_ = 0
        self.last_evaluations_ = 0
        self.last_launches_ = 0

    def setMaxIterations(self, iters):
        self.max_iters_ = int(iters)

    def setTolerance(self, tol):
        self.tol_ = float(tol)

    def setLineSearchParams(self, max_iters, c1, rho):
        self.max_line_iters_, self.c1_, self.rho_ = max(1, int(max_iters)), float(c1), float(rho)  # minimizer_base.cuh:38-45

    def setRecorder(self, recorder):
        self.recorder_ = recorder

    def iterations(self):
        return self.last_iterations_

    def _callback(self, loss_grad):
        if loss_grad is None or isinstance(loss_grad, CudaNetwork):
            return _lib.LOSS_GRAD_FN(), (loss_grad._h if loss_grad is not None else None)

        def tramp(user, params, grad, inp, tgt, batch):
            return float(loss_grad(params, grad, inp, tgt, batch))

        return _lib.LOSS_GRAD_FN(tramp), None

    def _history(self):
        if self.recorder_ is None:
            h = _lib.History()
            return h
        self.recorder_.reset()
        return self.recorder_._c_history()

    def _finish(self, h):
        self.last_iterations_ = h.iterations
        self.last_evaluations_ = h.evaluations
        self.last_launches_ = h.launches
        if self.recorder_ is not None:
            self.recorder_._absorb(h)


class CudaLBFGS(CudaMinimizerBase):
    def __init__(self, handle):
        super().__init__(handle)
        self.m_ = 16  # lbfgs.cuh:263
        self.linesearch_ = "armijo"
        self.c2_ = 0.9

    def setMemory(self, m):
        self.m_ = int(m)

    def setShardHistory(self, mode):
        """multi-GPU: 0 replicated, 1 sharded by parameter index, -1 auto (see include/b200_lbfgs.h)"""
        self.shard_history_ = int(mode)

    def setLineSearchPolicy(self, policy, c2=0.9):
        """'armijo' = reference CUDA backend; 'wolfe' = reference CPU backend"""
        self.linesearch_, self.c2_ = policy, float(c2)

    def _opts(self):
        o = _lib.LbfgsOpts()
        lib().b200_lbfgs_default_opts(C.byref(o))
        o.max_iters, o.tol, o.memory = self.max_iters_, self.tol_, self.m_
        o.max_line_iters, o.c1, o.rho, o.c2 = self.max_line_iters_, self.c1_, self.rho_, self.c2_
        o.linesearch = LS[self.linesearch_]
        o.record_timing = 1
        o.shard_history = getattr(self, "shard_history_", -1)
        return o

    # resumable form (b200_lbfgs_create / run / destroy): the same minimisation continued in slices
    def begin(self, n):
        self.end()
        self._solver = C.c_void_p()
        o = self._opts()
        check(lib().b200_lbfgs_create(self.handle._h, int(n), C.byref(o), C.byref(self._solver)))

    def run(self, params, input_dev, target_dev, batch, iters, loss_grad):
        cb, net_h = self._callback(loss_grad)
        h = self._history()
        check(lib().b200_lbfgs_run(self._solver, net_h, cb, None, C.c_void_p(_ptr(params)), C.c_void_p(_ptr(input_dev)),
                                   C.c_void_p(_ptr(target_dev)), int(batch), int(iters), C.byref(h)))
        self._finish(h)

    def end(self):
        if getattr(self, "_solver", None):
            lib().b200_lbfgs_destroy(self._solver)
        self._solver = None

    def solve(self, n, params, input_dev, target_dev, batch, loss_grad):
        """loss_grad: a CudaNetwork (fast path, the network's own objective) or a Python callable
        f(params_ptr, grad_ptr, input_ptr, target_ptr, batch) -> float (the reference's LossGradFun)."""
        o = self._opts()
        cb, net_h = self._callback(loss_grad)
        h = self._history()
        check(lib().b200_lbfgs_solve(self.handle._h, net_h, cb, None, int(n), C.c_void_p(_ptr(params) or 0),
                                     C.c_void_p(_ptr(input_dev)), C.c_void_p(_ptr(target_dev)), int(batch),
                                     C.byref(o), C.byref(h)))
        self._finish(h)


class CudaGD(CudaMinimizerBase):
    def __init__(self, handle):
        super().__init__(handle)
        self.lr_, self.momentum_ = 0.01, 0.9  # gd.cuh:108-110

    def setLearningRate(self, lr):
        self.lr_ = float(lr)

    def setMomentum(self, m):
        self.momentum_ = float(m)

    def solve(self, n, params, input_dev, target_dev, batch, loss_grad):
        o = _lib.GdOpts()
        lib().b200_gd_default_opts(C.byref(o))
        o.max_iters, o.tol, o.lr, o.momentum = self.max_iters_, self.tol_, self.lr_, self.momentum_
        cb, net_h = self._callback(loss_grad)
        h = self._history()
        check(lib().b200_gd_solve(self.handle._h, net_h, cb, None, int(n), C.c_void_p(_ptr(params) or 0),
                                  C.c_void_p(_ptr(input_dev)), C.c_void_p(_ptr(target_dev)), int(batch), C.byref(o),
                                  C.byref(h)))
        self._finish(h)


class CudaSGD(CudaMinimizerBase):
    def __init__(self, handle):
        super().__init__(handle)
        self.lr_, self.momentum_, self.decay_rate_, self.decay_step_ = 0.01, 0.9, 1.0, 0  # sgd.cuh:156-163
        self.batch_size_, self.input_dim_, self.output_dim_ = 64, 0, 0
        self.sampling_, self.seed_ = 0, 123

    def setLearningRate(self, lr):
        self.lr_ = float(lr)

    def setMomentum(self, m):
        self.momentum_ = float(m)

    def setBatchSize(self, b):
        self.batch_size_ = int(b)

    def setLearningRateDecay(self, rate, step):
        self.decay_rate_, self.decay_step_ = float(rate), int(step)

    def setDimensions(self, in_dim, out_dim):
        self.input_dim_, self.output_dim_ = int(in_dim), int(out_dim)

    def setSampling(self, mode, seed=123):
        """'sequential' (the CUDA backend's CudaSGD) or 'random' (the CPU backend's s_gd.hpp mini-batches, on the GPU)"""
        self.sampling_, self.seed_ = {"sequential": 0, "random": 1}[mode], int(seed)

    def solve(self, n, params, input_dev, target_dev, total_samples, loss_grad):
        o = _lib.SgdOpts()
        lib().b200_sgd_default_opts(C.byref(o))
        o.max_iters, o.tol, o.lr, o.momentum = self.max_iters_, self.tol_, self.lr_, self.momentum_
        o.decay_rate, o.decay_step, o.batch_size = self.decay_rate_, self.decay_step_, self.batch_size_
        o.input_dim, o.output_dim = self.input_dim_, self.output_dim_
        o.sampling, o.seed = self.sampling_, self.seed_
        cb, net_h = self._callback(loss_grad)
        h = self._history()
        check(lib().b200_sgd_solve(self.handle._h, net_h, cb, None, int(n), C.c_void_p(_ptr(params) or 0),
                                   C.c_void_p(_ptr(input_dev)), C.c_void_p(_ptr(target_dev)), int(total_samples),
                                   C.byref(o), C.byref(h)))
        self._finish(h)


class CudaSLBFGS(CudaMinimizerBase):
    """S-LBFGS on the GPU (SLBFGS::stochastic_solve + the UnifiedSLBFGS_CPU objective, lambda = 1e-4)."""

    def __init__(self, handle):
        super().__init__(handle)
        self.tol_ = 1e-4
        self.step_size_, self.batch_size_, self.M_, self.L_, self.b_H_ = 0.01, 128, 10, 10, 0
        self.lambda_, self.epsilon_, self.seed_ = 1e-4, 1e-4, 123
        self.hvp_step_scale_ = 256.0  # include/b200_lbfgs.h: b200_slbfgs_opts::hvp_step_scale
        self.pair_eval_ = 1          # b200_slbfgs_opts::pair_eval

    def setStepSize(self, s):
        self.step_size_ = float(s)

    def setBatchSize(self, b):
        self.batch_size_ = int(b)

    def setMemory(self, m):
        self.M_ = int(m)

    def setUpdateInterval(self, L):
        self.L_ = int(L)

    def setHessianBatchSize(self, b_H):
        self.b_H_ = int(b_H)

    def setSeed(self, seed):
        self.seed_ = int(seed)

    def setPairEvaluation(self, on):
        """True (default): both evaluations of a step in one forward/backward of the stacked pair network"""
        self.pair_eval_ = 1 if on else 0

    def setHvpStepScale(self, scale):
        """1 = the reference's finite-difference step verbatim; default 256 (fp32 resolution, see include/b200_lbfgs.h)"""
        self.hvp_step_scale_ = float(scale)

    def solve(self, n, params, input_dev, target_dev, total_samples, net):
        o = _lib.SlbfgsOpts()
        lib().b200_slbfgs_default_opts(C.byref(o))
        o.max_iters, o.tol, o.step_size, o.batch_size = self.max_iters_, self.tol_, self.step_size_, self.batch_size_
        o.memory, o.L, o.b_H, o.lam, o.epsilon, o.seed = self.M_, self.L_, self.b_H_, self.lambda_, self.epsilon_, self.seed_
        o.record = 1 if self.recorder_ is not None else 0
        o.hvp_step_scale = self.hvp_step_scale_
        o.pair_eval = self.pair_eval_
        h = self._history()
        check(lib().b200_slbfgs_solve(self.handle._h, net._h, int(n), C.c_void_p(_ptr(params) or 0),
                                      C.c_void_p(_ptr(input_dev)), C.c_void_p(_ptr(target_dev)), int(total_samples),
                                      C.byref(o), C.byref(h)))
        self._finish(h)


# ---- building blocks exposed for parity tests -------------------------------------------------------
def lbfgs_direction(handle, S_dev, Y_dev, rho, g_dev, n, k, p_dev, policy="armijo"):
    pol = {"armijo": 0, "cuda": 0, "wolfe": 1, "cpu": 1, "slbfgs": 2}[policy]
    rho = np.ascontiguousarray(rho, dtype=np.float32)
    gdp = C.c_double()
    check(lib().b200_lbfgs_direction(handle._h, int(n), int(k), C.c_void_p(_ptr(S_dev) or 0), C.c_void_p(_ptr(Y_dev) or 0),
                                     rho.ctypes.data_as(C.POINTER(C.c_float)), C.c_void_p(_ptr(g_dev)), pol,
                                     C.c_void_p(_ptr(p_dev)), C.byref(gdp)))
    return gdp.value


def device_dot(handle, x, y, n):
    out = C.c_double()
    check(lib().b200_vec_dot(handle._h, C.c_void_p(_ptr(x)), C.c_void_p(_ptr(y)), int(n), C.byref(out)))
    return out.value


def device_nrm2(handle, x, n):
    out = C.c_double()
    check(lib().b200_vec_nrm2(handle._h, C.c_void_p(_ptr(x)), int(n), C.byref(out)))
    return out.value


def device_axpy(handle, n, alpha, x, y):
    check(lib().b200_vec_axpy(handle._h, int(n), float(alpha), C.c_void_p(_ptr(x)), C.c_void_p(_ptr(y))))


def device_scal(handle, n, alpha, x):
    check(lib().b200_vec_scal(handle._h, int(n), float(alpha), C.c_void_p(_ptr(x))))


def slbfgs_sample_stream(seed, N, b, count):
    out = np.empty(count * min(b, N), dtype=np.uint32)
    check(lib().b200_slbfgs_sample_stream(int(seed), int(N), int(b), int(count), out.ctypes.data_as(C.c_void_p)))
    return out.reshape(count, min(b, N))


def launch_count():
    return lib().b200_launch_count()


def reload_env():
    """re-read the B200_* debugging switches (the library reads the environment once, at first use)"""
    check(lib().b200_debug_reload_env())
