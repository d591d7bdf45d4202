"""S-LBFGS, BASELINE configs[3] size, M = 10: per-epoch loss against the fp64 oracle for several finite-difference step scales
(b200_slbfgs_opts::hvp_step_scale). usage: python tools/slbfgs_hvp_scale.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lbfgs_ffnn_b200 as P
from oracle import binding as ob
from helpers import make_problem, make_gpu_net, upload
dims, acts, N, E = [784, 128, 64, 10], ["relu", "relu", "linear"], 60000, 3
onet, w, X, T = make_problem(ob, dims, acts, N)
ref = onet.slbfgs(w, X, T, batch_size=1000, M=10, L=10, b_H=5000, step=0.02, max_iters=E, tol=0.0, seed=123)
print("oracle", ref["loss"], ref["gnorm"], flush=True)
h = P.CublasHandle(0)
dx, dt = upload(X), upload(T)
for prec in ("fp32", "tf32x3"):
    for scale in (1, 16, 64, 1024, 16384):
        net = make_gpu_net(h, dims, acts, w, precision=prec)
        s = P.CudaSLBFGS(h)
        s.setMaxIterations(E); s.setTolerance(0.0); s.setStepSize(0.02); s.setBatchSize(1000)
        s.setMemory(10); s.setUpdateInterval(10); s.setHessianBatchSize(5000); s.setHvpStepScale(scale)
        rec = P.IterationRecorder(); rec.init(E); s.setRecorder(rec)
        s.solve(net.params_size(), net.params_data(), dx, dt, N, net)
        loss, gn, _ = rec.copy_to_host()
        print(json.dumps(dict(prec=prec, scale=scale, loss=[float(v) for v in loss], rel=[float(abs(a - b) / b) for a, b in zip(loss, ref["loss"])])), flush=True)
        net.close()
