"""BASELINE configs[3]: S-LBFGS (SVRG + finite-difference HVP pairs) on 784-128-64-10, 60 000 samples, mini-batch 1000,
b_H = 5000, M = 10, L = 10. Prints GPU epochs/s (device-resident, CUDA events around b200_slbfgs_solve) and, with --cpu, the
oracle port's time for ONE epoch on the host cores.
usage: python tools/slbfgs_timing.py [epochs] [--cpu]"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import lbfgs_ffnn_b200 as P
from helpers import make_gpu_net, upload

epochs = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 10
dims, acts, N = [784, 128, 64, 10], ["relu", "relu", "linear"], 60000
h = P.CublasHandle(0)
stream = torch.cuda.Stream()
h.set_stream(stream.cuda_stream)
X, T = P.synthetic_mnist(N)
dx, dt = upload(X), upload(T)
net = make_gpu_net(h, dims, acts, None, precision="tf32x3")
w0 = net.get_params()

def run(ep):
    net.set_params(w0)
    s = P.CudaSLBFGS(h)
    s.setMaxIterations(ep); s.setTolerance(0.0); s.setStepSize(0.02); s.setBatchSize(1000)
    s.setMemory(10); s.setUpdateInterval(10); s.setHessianBatchSize(5000)
    rec = P.IterationRecorder(); rec.init(ep); s.setRecorder(rec)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = P.api.launch_count()
    e0.record(stream)
    s.solve(net.params_size(), net.params_data(), dx, dt, N, net)
    e1.record(stream)
    torch.cuda.synchronize()
    loss, gn, _ = rec.copy_to_host()
    return e0.elapsed_time(e1), loss, P.api.launch_count() - l0

run(2)
ms, loss, launches = run(epochs)
out = {"workload": "slbfgs_mlp784-128-64-10_N60000_b1000_bH5000_M10_L10", "epochs": epochs, "ms_per_epoch": ms / epochs,
       "epochs_per_s": 1e3 * epochs / ms, "kernel_launches_per_epoch": launches / epochs,
       "loss_first": float(loss[0]), "loss_last": float(loss[-1])}
if "--cpu" in sys.argv:
    from oracle import binding as ob
    onet = ob.OracleNet(dims, acts)
    t0 = time.time()
    ref = onet.slbfgs(w0, X, T, batch_size=1000, M=10, L=10, b_H=5000, step=0.02, max_iters=1, tol=0.0, seed=123)
    out["cpu_port_s_per_epoch"] = time.time() - t0
    out["cpu_loss_first"] = float(ref["loss"][0])
print(json.dumps(out))
