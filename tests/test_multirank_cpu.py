"""World-size-2 tests on CPU (gloo): the host-side contract of the sample-sharded multi-GPU path.

The product's N>1 data path is: rank r evaluates its contiguous shard with the GLOBAL 1/B scaling
(b200_net_set_global_batch), then one sum-allreduce of [gradient | loss]; everything else (history, direction,
line search) is replicated. These tests check that contract with the oracle as the per-rank evaluator and gloo as
the collective, and the S-LBFGS index partitioning that b200_slbfgs_solve uses per rank."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as ob
    import lbfgs_ffnn_b200 as P
    from lbfgs_ffnn_b200 import api
    ob.set_num_threads(2)
    dims, acts, B = [784, 32, 10], ["relu", "linear"], 600
    X, T = P.synthetic_mnist(B)
    net = ob.OracleNet(dims, acts)
    w = net.init_params_cuda_rule(123)
    shard = B // world
    xs, ts = X[rank * shard:(rank + 1) * shard], T[rank * shard:(rank + 1) * shard]
    # per-rank evaluation scaled by the global batch: oracle loss/grad are means over the shard -> rescale
    l, g = net.loss_grad(w, xs, ts)
    buf = torch.from_numpy(np.concatenate([g * (shard / B), [l * (shard / B)]]))
    dist.all_reduce(buf)  # the one exchange step per evaluation
    lf, gf = net.loss_grad(w, X, T)
    ok_grad = np.linalg.norm(buf[:-1].numpy() - gf) <= 1e-12 * np.linalg.norm(gf)
    ok_loss = abs(buf[-1].item() - lf) <= 1e-12 * abs(lf)
    # S-LBFGS: every rank draws the SAME index stream (same seed) and takes its contiguous chunk of each batch
    idx = api.slbfgs_sample_stream(123, B, 100, 5)
    mine = idx[:, rank * (100 // world):(rank + 1) * (100 // world)]
    gathered = [torch.zeros_like(torch.from_numpy(mine.astype(np.int64))) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mine.astype(np.int64)))
    ok_idx = np.array_equal(np.concatenate([t.numpy() for t in gathered], axis=1), idx.astype(np.int64))
    if rank == 0:
        q.put((bool(ok_grad), bool(ok_loss), bool(ok_idx)))
    dist.destroy_process_group()


def test_sample_sharding_contract_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert res == (True, True, True), res
    assert all(p.exitcode == 0 for p in procs)


def test_examples_compile():
    """the C++ drop-in headers (include/cuda_mlp, include/unified) build the reference-style runner with g++ alone"""
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "-B", "-s"])
    assert os.path.exists(os.path.join(ROOT, "examples", "main_gpu_synthetic"))
