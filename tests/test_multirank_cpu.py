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


import subprocess  # noqa: E402


def test_reference_runner_mains_compile_unmodified():
    """tests/mnist/main-gpu.cpp, tests/fashion-mnist/main-gpu.cpp and main_gpu_deep.cpp of the reference, byte for byte as they
    lie in its checkout, compile and link against include/compat + libb200lbfgs.so (examples/build_reference_runners.sh): the
    literal meaning of "drop-in backend". Only where the reference checkout exists (this container)."""
    import pytest
    if not os.path.isdir("/root/reference/tests"):
        pytest.skip("no reference checkout on this machine")
    out = subprocess.run([os.path.join(ROOT, "examples", "build_reference_runners.sh")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    for name in ("mnist_main_gpu", "fashion_mnist_main_gpu", "fashion_mnist_main_gpu_deep"):
        exe = os.path.join(ROOT, "examples", "_ref", name)
        assert os.path.exists(exe) and os.access(exe, os.X_OK)
        assert f"built examples/_ref/{name}" in out.stdout


def test_idx_loader_reads_the_format_and_falls_back(tmp_path):
    """the drop-in MNISTLoader (include/compat/tests/mnist/mnist_loader.hpp): IDX files are read like the reference reads them
    (tests/mnist/mnist_loader.hpp:19-100), a missing blob gives synthetic data of the requested size, a non-IDX file throws"""
    import struct
    import numpy as np
    rs = np.random.RandomState(0)
    pix = rs.randint(0, 256, size=(5, 28 * 28), dtype=np.uint8)
    lab = np.array([3, 0, 9, 9, 1], dtype=np.uint8)
    (tmp_path / "img.idx3").write_bytes(struct.pack(">IIII", 2051, 5, 28, 28) + pix.tobytes())
    (tmp_path / "lab.idx1").write_bytes(struct.pack(">II", 2049, 5) + lab.tobytes())
    (tmp_path / "bad.idx3").write_bytes(b"not an idx file at all")
    src = tmp_path / "t.cpp"
    src.write_text('''
#include "compat/tests/mnist/mnist_loader.hpp"
#include <cstdio>
int main(int argc, char **argv) {
  Eigen::MatrixXd x = MNISTLoader::loadImages(argv[1], 4), y = MNISTLoader::loadLabels(argv[2], 4);
  std::printf("%ld %ld %ld %ld\\n", x.rows(), x.cols(), y.rows(), y.cols());
  for (int i = 0; i < 4; ++i) { int a = 0; for (int r = 0; r < 10; ++r) if (y(r, i) == 1.0) a = r; std::printf("%d ", a); }
  std::printf("\\n%.9g %.9g\\n", x(0, 0), x(783, 3));
  Eigen::MatrixXd s = MNISTLoader::loadImages(argv[3], 7);
  std::printf("%ld %ld\\n", s.rows(), s.cols());
  try { MNISTLoader::loadImages(argv[4], 1); std::printf("no throw\\n"); } catch (const std::exception &e) { std::printf("threw\\n"); }
  return 0;
}''')
    exe = tmp_path / "t"
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src),
                           "-L", os.path.join(ROOT, "lbfgs_ffnn_b200"), "-lb200lbfgs", "-Wl,-rpath," + os.path.join(ROOT, "lbfgs_ffnn_b200")])
    out = subprocess.run([str(exe), str(tmp_path / "img.idx3"), str(tmp_path / "lab.idx1"), str(tmp_path / "missing"),
                          str(tmp_path / "bad.idx3")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.strip().splitlines() if not l.startswith("Loading ")]  # the loader's own banners
    assert lines[0] == "784 4 10 4"
    assert lines[1].split() == ["3", "0", "9", "9"]
    a, b = map(float, lines[2].split())
    # printed with 9 significant digits: enough to identify a float32
    assert np.float32(a) == np.float32(pix[0, 0]) / np.float32(255.0) and np.float32(b) == np.float32(pix[3, 783]) / np.float32(255.0)
    assert lines[3] == "784 7" and lines[4] == "threw"
    assert "synthetic" in out.stderr
