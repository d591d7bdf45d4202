"""2-GPU test (needs >= 2 devices): sample-sharded L-BFGS over NCCL matches the single-GPU run."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_rank_lbfgs_matches_single():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.strip().split("\n") if l.startswith("{")][-1]
    r = json.loads(line)
    # a different reduction order (shard partials + NCCL) perturbs the fp32 gradient in the last bits; the trajectories
    # then drift apart geometrically like any two fp32 runs (compare tests/test_gpu_vs_reference_cuda.py)
    # (measured over 25 iterations: identical for the first 5, 2.7e-3 in the loss and 1.1e-2 in the parameters at the end)
    assert r["max_rel_loss_diff_first5"] <= 2e-5, r
    assert r["max_rel_loss_diff"] <= 1e-2, r
    assert r["params_rel_l2"] <= 3e-2, r
    # history sharded by parameter index (reduce-scatter + all-reduce of the 5(m+1)+1 partial dots + all-gather)
    assert r["sharded_max_rel_loss_diff_first5"] <= 2e-5, r
    assert r["sharded_max_rel_loss_diff"] <= 1e-2, r
    assert r["sharded_params_rel_l2"] <= 3e-2, r
