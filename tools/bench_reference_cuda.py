"""On-box comparator: the REFERENCE's CUDA backend (oracle/_ref/libref_cuda.so, cuBLAS + its own kernels, compiled
unmodified for sm_100a) timed on the same workload as bench.py (784-128-10, 60 000 samples, L-BFGS m=10).
Usage (GPU box): python tools/bench_reference_cuda.py [iters] > gpurun_out/ref_cuda_bench.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lbfgs_ffnn_b200 as P
from oracle import ref_cuda_binding as rc

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
out = {}
for name, dims, acts in (("784-128-10", [784, 128, 10], [2, 0]), ("784-128-64-10", [784, 128, 64, 10], [2, 2, 0])):
    X, T = P.synthetic_mnist(60000)
    dx, dt = P.DeviceBuffer(), P.DeviceBuffer()
    dx.copy_from_host(X); dt.copy_from_host(T)
    net = rc.RefCudaNet(dims, acts)
    net.bind_params(123)
    net.solve("lbfgs", dx.data(), dt.data(), 60000, 5, memory=10, record=False)  # warm-up (cuBLAS init, allocations)
    net.bind_params(123)
    r = net.solve("lbfgs", dx.data(), dt.data(), 60000, iters, memory=10, record=False)
    out[name] = dict(iterations=r["iters"], total_ms=r["total_ms"], iters_per_sec=r["iters"] / (r["total_ms"] / 1e3))
    net.close()
print(json.dumps({"impl": "reference CUDA backend (cuBLAS SGEMM, src/cuda/*.cuh) on this B200", "workload": "L-BFGS m=10, 60000 samples", **out}))
