import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lbfgs_ffnn_b200 as P
from helpers import make_gpu_net, upload
dims, acts, N, E = [784, 128, 64, 10], ["relu", "relu", "linear"], 60000, 6
h = P.CublasHandle(0)
X, T = P.synthetic_mnist(N)
dx, dt = upload(X), upload(T)
for prec in ("fp32", "tf32x3"):
    for pair in (1, 0):
        for scale in (1, 16, 256):
            net = make_gpu_net(h, dims, acts, None, precision=prec)
            s = P.CudaSLBFGS(h)
            s.setMaxIterations(E); s.setTolerance(0.0); s.setStepSize(0.02); s.setBatchSize(1000)
            s.setMemory(10); s.setUpdateInterval(10); s.setHessianBatchSize(5000); s.setPairEvaluation(pair); s.setHvpStepScale(scale)
            rec = P.IterationRecorder(); rec.init(E); s.setRecorder(rec)
            s.solve(net.params_size(), net.params_data(), dx, dt, N, net)
            print(json.dumps(dict(prec=prec, pair=pair, scale=scale, loss=[round(float(v), 4) for v in rec.copy_to_host()[0]])), flush=True)
            net.close()
