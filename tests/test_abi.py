"""CPU suite, part 2: the C-ABI library loads here (no GPU) and exports every symbol include/b200_lbfgs.h declares;
the product fails loudly — no CPU fallback — when no device is present."""
import ctypes
import os
import re

import pytest

from conftest import HAS_GPU, ROOT

HEADER = os.path.join(ROOT, "include", "b200_lbfgs.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if n != "b200_loss_grad_fn"))


def test_header_symbols_exported():
    from lbfgs_ffnn_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 45
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/b200_lbfgs.h but not exported"


def test_binding_covers_header():
    from lbfgs_ffnn_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared_functions()
    _lib.lib()  # sets restype/argtypes for every symbol; raises on a missing one


def test_no_cublas_linked():
    import subprocess
    from lbfgs_ffnn_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "cublas" not in out.lower()  # BASELINE.json north_star: no cuBLAS on the hot path


def test_abi_version_and_defaults():
    from lbfgs_ffnn_b200 import _lib
    L = _lib.lib()
    assert L.b200_abi_version() == 1
    o = _lib.LbfgsOpts()
    L.b200_lbfgs_default_opts(ctypes.byref(o))
    # src/cuda/minimizer_base.cuh:61-65, src/cuda/lbfgs.cuh:263
    assert (o.max_iters, o.memory, o.max_line_iters, o.linesearch) == (200, 16, 20, 0)
    assert abs(o.tol - 1e-6) < 1e-12 and abs(o.c1 - 1e-4) < 1e-10 and o.rho == 0.5
    g = _lib.GdOpts()
    L.b200_gd_default_opts(ctypes.byref(g))
    assert abs(g.lr - 0.01) < 1e-9 and abs(g.momentum - 0.9) < 1e-7  # src/cuda/gd.cuh:108-110
    s = _lib.SgdOpts()
    L.b200_sgd_default_opts(ctypes.byref(s))
    assert (s.batch_size, s.decay_step, s.input_dim) == (64, 0, 0)  # src/cuda/sgd.cuh:156-163


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device behaviour")
def test_fails_loudly_without_device():
    import lbfgs_ffnn_b200 as P
    with pytest.raises(P._lib.B200Error):
        P.CublasHandle(0)


def test_history_csv_schema(tmp_path):
    # scripts/plot_results.py:26-29 reads exactly these columns
    import numpy as np
    from lbfgs_ffnn_b200 import api
    r = api.IterationRecorder()
    r.init(5)
    r._loss[:5] = [1.0, 0.5, 0.25, 0.125, 0.0625]
    r._grad[:5] = [3, 2, 1, 0.5, 0.25]
    r._time[:5] = np.arange(5) * 1.5
    r._size = 5
    f = tmp_path / "x_history.csv"
    api.write_cuda_history_csv(str(f), r, 2)
    lines = f.read_text().strip().split("\n")
    assert lines[0] == "Iteration,Loss,GradNorm,TimeMs"
    assert [l.split(",")[0] for l in lines[1:]] == ["0", "2", "4"]
    assert lines[2] == "2,0.25,1,3"
