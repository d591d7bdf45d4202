"""Where does the host-buffers -> solve -> host path spend its time? (GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import lbfgs_ffnn_b200 as P
B = 60000
Xh, Th = P.synthetic_mnist(B)
Xs, Ts = torch.from_numpy(Xh).pin_memory(), torch.from_numpy(Th).pin_memory()
h = P.CublasHandle(0)
stream = torch.cuda.Stream(); h.set_stream(stream.cuda_stream)
net = P.CudaNetwork(h); net.addLayer(784, 128, "relu"); net.addLayer(128, 10, "linear"); net.bindParams(123); net.set_precision("tf32x3")
n = net.params_size(); w0 = torch.from_numpy(net.get_params().copy()).pin_memory()
with torch.cuda.stream(stream):
    dX = torch.empty_like(Xs, device="cuda"); dT = torch.empty_like(Ts, device="cuda")
def t(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(4):
    a = t()
    with torch.cuda.stream(stream):
        dX.copy_(Xs, non_blocking=True); dT.copy_(Ts, non_blocking=True)
        wd = torch.empty(n, dtype=torch.float32, device="cuda"); wd.copy_(w0, non_blocking=True)
    b = t()
    s = P.CudaLBFGS(h); s.setMemory(10); s.setMaxIterations(200); s.setTolerance(0.0)
    rec = P.IterationRecorder(); rec.init(200); s.setRecorder(rec)
    c = t()
    s.solve(n, wd, dX, dT, B, net)
    d = t()
    out = torch.empty(n, dtype=torch.float32).pin_memory()
    with torch.cuda.stream(stream): out.copy_(wd, non_blocking=True)
    e = t()
    print(f"rep {rep}: H2D {1e3*(b-a):.2f} ms | solver create {1e3*(c-b):.2f} | solve(200 it) {1e3*(d-c):.2f} | D2H {1e3*(e-d):.2f} | its {s.iterations()}", flush=True)
