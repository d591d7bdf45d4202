"""UnifiedLauncher<CudaBackend> / UnifiedOptimizer<CudaBackend> mirror (src/unified_launcher.hpp:83-205,
src/unified_optimization.hpp:26-48,410-643). Backend selection in the reference is a compile-time tag; this
package is the CudaBackend half only — there is no CPU backend here and nothing falls back to one."""
import os
from dataclasses import dataclass

import numpy as np

from . import api


@dataclass
class UnifiedConfig:  # src/unified_optimization.hpp:26-48 (+ new fields with defaults)
    name: str = "Experiment"
    max_iters: int = 100
    tolerance: float = 1e-4
    learning_rate: float = 0.01
    momentum: float = 0.0
    lr_decay: float = 0.0
    lr_decay_rate: int = 1
    batch_size: int = 128
    m_param: int = 10
    L_param: int = 10
    b_H_param: int = 0
    log_interval: int = 10
    reset_params: bool = True
    seed: int = 123
    # additions
    linesearch: str = "armijo"   # 'armijo' (reference CUDA backend) | 'wolfe' (reference CPU backend)
    precision: str = "fp32"      # 'fp32' | 'tf32x3' | 'tf32'
    log_dir: str = "."


@dataclass
class UnifiedDataset:  # column-major like the reference's Eigen matrices: x is (samples, in) C-contiguous
    train_x: np.ndarray
    train_y: np.ndarray
    test_x: np.ndarray = None
    test_y: np.ndarray = None


def cuda_log_filename(config):
    return (config.name if config.name else "run") + "_history.csv"


class UnifiedOptimizer:
    def optimize(self, handle, net, dataset, dx, dy, config):
        raise NotImplementedError

    last_solver = None

    def _run(self, make_solver, net, dx, dy, dataset, config, n_samples):
        """run_cuda_solver_once (src/unified_optimization.hpp:470-515)"""
        solver = make_solver()
        recorder = api.IterationRecorder()
        recorder.init(config.max_iters + 1)
        solver.setRecorder(recorder)
        net.set_precision(config.precision)
        solver.solve(net.params_size(), net.params_data(), dx, dy, n_samples, net)
        net.handle.synchronize()
        api.write_cuda_history_csv(os.path.join(config.log_dir, cuda_log_filename(config)), recorder, config.log_interval)
        self.last_solver, self.last_recorder = solver, recorder


class UnifiedGD(UnifiedOptimizer):
    def optimize(self, handle, net, dataset, dx, dy, c):
        def make():
            s = api.CudaGD(handle)
            s.setLearningRate(c.learning_rate); s.setMomentum(c.momentum)
            s.setMaxIterations(c.max_iters); s.setTolerance(c.tolerance)
            return s
        self._run(make, net, dx, dy, dataset, c, dataset.train_x.shape[0])


class UnifiedLBFGS(UnifiedOptimizer):
    def optimize(self, handle, net, dataset, dx, dy, c):
        def make():
            s = api.CudaLBFGS(handle)
            s.setMemory(c.m_param); s.setMaxIterations(c.max_iters); s.setTolerance(c.tolerance)
            s.setLineSearchPolicy(c.linesearch)
            if c.linesearch == "wolfe":
                s.setLineSearchParams(50, 1e-4, 0.5)  # full_batch_minimizer.hpp:113-116
            return s
        self._run(make, net, dx, dy, dataset, c, dataset.train_x.shape[0])


class UnifiedSGD(UnifiedOptimizer):
    def optimize(self, handle, net, dataset, dx, dy, c):
        def make():
            s = api.CudaSGD(handle)
            s.setLearningRate(c.learning_rate); s.setMomentum(c.momentum); s.setBatchSize(c.batch_size)
            s.setMaxIterations(c.max_iters); s.setLearningRateDecay(c.lr_decay, c.lr_decay_rate)
            s.setDimensions(dataset.train_x.shape[1], dataset.train_y.shape[1])
            return s
        self._run(make, net, dx, dy, dataset, c, dataset.train_x.shape[0])


class UnifiedSLBFGS(UnifiedOptimizer):
    """Available on the GPU here; the reference static_asserts (src/unified_optimization.hpp:639-641)."""

    def optimize(self, handle, net, dataset, dx, dy, c):
        def make():
            s = api.CudaSLBFGS(handle)
            s.setMaxIterations(c.max_iters); s.setTolerance(c.tolerance)
            s.setStepSize(c.learning_rate); s.setBatchSize(c.batch_size)
            s.setMemory(c.m_param); s.setUpdateInterval(c.L_param); s.setHessianBatchSize(c.b_H_param)
            s.setSeed(123)  # the reference seeds S-LBFGS with kDefaultSeed, not config.seed (s_lbfgs.hpp:183)
            return s
        self._run(make, net, dx, dy, dataset, c, dataset.train_x.shape[0])


class UnifiedLauncher:
    def __init__(self, device=0, verbose=True):
        self.handle_ = api.CublasHandle(device)
        self.net_ = api.CudaNetwork(self.handle_)
        self.dataset_ = None
        self.verbose = verbose
        self.d_train_x_, self.d_train_y_ = api.DeviceBuffer(), api.DeviceBuffer()
        self.d_test_x_, self.d_test_y_ = api.DeviceBuffer(), api.DeviceBuffer()

    def addLayer(self, in_dim, out_dim, act):
        self.net_.addLayer(in_dim, out_dim, act)

    def buildNetwork(self):
        self.net_.bindParams()

    def getNetwork(self):
        return self.net_

    def setData(self, data):
        self.dataset_ = data
        self.d_train_x_.copy_from_host(data.train_x)  # double -> float conversion as unified_launcher.hpp:109-121
        self.d_train_y_.copy_from_host(data.train_y)
        if data.test_x is not None:
            self.d_test_x_.copy_from_host(data.test_x)
            self.d_test_y_.copy_from_host(data.test_y)
        if self.verbose:
            print(f"Data Uploaded to GPU. Train: {data.train_x.shape[0]} samples.")

    def train(self, optimizer, config):
        if self.verbose:
            print(f">>> Running CUDA Experiment: {config.name}")
        if config.reset_params:
            self.net_.bindParams(config.seed)
        optimizer.optimize(self.handle_, self.net_, self.dataset_, self.d_train_x_, self.d_train_y_, config)
        return self._evaluate(self.d_train_x_, self.d_train_y_, self.dataset_.train_x.shape[0], "Training Results")

    def test(self):
        return self._evaluate(self.d_test_x_, self.d_test_y_, self.dataset_.test_x.shape[0], "Test Results")

    def _evaluate(self, dx, dy, batch, label):
        mse, acc = self.net_.evaluate(dx, dy, batch)
        if self.verbose:
            print(f"{label}: MSE={mse:g}, Accuracy={acc:g}%")
        return mse, acc
