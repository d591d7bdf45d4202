"""S-LBFGS at BASELINE configs[3] size: per-scope CUDA-event times of an epoch, pair network on / off.
usage: python tools/slbfgs_profile.py [epochs]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import lbfgs_ffnn_b200 as P
from helpers import make_gpu_net, upload
E = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dims, acts, N = [784, 128, 64, 10], ["relu", "relu", "linear"], 60000
h = P.CublasHandle(0)
stream = torch.cuda.Stream(); h.set_stream(stream.cuda_stream)
X, T = P.synthetic_mnist(N)
dx, dt = upload(X), upload(T)
for prec in ("tf32x3",):
    for pair, scale, M in ((1, 16, 10), (0, 16, 10), (1, 1, 10), (1, 16, 0)):
        net = make_gpu_net(h, dims, acts, None, precision=prec)
        w0 = net.get_params()
        def run(ep, prof=False):
            net.set_params(w0)
            s = P.CudaSLBFGS(h)
            s.setMaxIterations(ep); s.setTolerance(0.0); s.setStepSize(0.02); s.setBatchSize(1000)
            s.setMemory(M); s.setUpdateInterval(10); s.setHessianBatchSize(5000); s.setPairEvaluation(pair); s.setHvpStepScale(scale)
            rec = P.IterationRecorder(); rec.init(ep); s.setRecorder(rec)
            if prof: h.profile(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            s.solve(net.params_size(), net.params_data(), dx, dt, N, net)
            e1.record(stream); torch.cuda.synchronize()
            rep = h.profile_report() if prof else None
            if prof: h.profile(False)
            return e0.elapsed_time(e1), rec.copy_to_host()[0], s.last_launches_, rep
        run(2)
        ms, loss, launches, _ = run(E)
        _, _, _, rep = run(2, prof=True)
        tot = {k: round(v[1], 2) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:12]}
        avg = {k: round(1e3 * v[1] / v[0], 1) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:12]}
        print(json.dumps(dict(prec=prec, pair=pair, scale=scale, M=M, ms_per_epoch=ms / E, launches_per_epoch=launches / E,
                              loss=[float(v) for v in loss], total_ms_2_epochs=tot, avg_us=avg)), flush=True)
        net.close()
