"""lbfgs_ffnn_b200 — B200-native (sm_100a) backend of SignorB/lbfgs-FFNN's optimizer-plus-objective hot path.

Only what the path needs: csrc/ (CUDA kernels + the C ABI of include/b200_lbfgs.h) and the host-side mirror of
the reference's CUDA operator interface (api.py, launcher.py). No CPU fallback: importing works anywhere, but
every entry point raises if libb200lbfgs.so is missing or no sm_100 device is present."""
from . import _lib
from .api import (CublasHandle, CudaGD, CudaLBFGS, CudaNetwork, CudaSGD, CudaSLBFGS, DeviceBuffer, IterationRecorder,
                  Linear, ReLU, Sigmoid, Tanh, write_cuda_history_csv)
from .data import kDefaultSeed, synthetic_mnist
from .launcher import (UnifiedConfig, UnifiedDataset, UnifiedGD, UnifiedLauncher, UnifiedLBFGS, UnifiedSGD, UnifiedSLBFGS)

__all__ = [n for n in dir() if not n.startswith("_")]
