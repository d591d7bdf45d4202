"""Per-kernel CUDA-event timings of forward_only / loss_grad in each precision mode (GPU box).
usage: python tools/kernel_probe.py [batch] [reps]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lbfgs_ffnn_b200 as P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
h = P.CublasHandle(0)
X, T = P.synthetic_mnist(B)
dx, dt = P.DeviceBuffer(), P.DeviceBuffer(); dx.copy_from_host(X); dt.copy_from_host(T)
for dims, acts in (([784,128,10],["relu","linear"]), ([784,128,64,10],["relu","relu","linear"])):
    for prec in ("tf32", "tf32x3", "fp32"):
        for mask in ((None,) if prec == "fp32" else (None, "15")):
            if mask: os.environ["B200_TC_MASK"] = mask
            else: os.environ.pop("B200_TC_MASK", None)
            P.api.reload_env()
            net = P.CudaNetwork(h)
            for i, a in enumerate(acts): net.addLayer(dims[i], dims[i+1], a)
            net.bindParams(123); net.set_precision(prec)
            for _ in range(3):
                net.forward_only(dx, B); net.compute_loss_and_grad(dx, dt, B)
            h.profile(True)
            for _ in range(reps):
                net.forward_only(dx, B)
                net.loss_grad_async(dx, dt, B)
            rep = h.profile_report(); h.profile(False)
            print(json.dumps({"net": "-".join(map(str, dims)), "prec": prec, "mask": mask or "7",
                              "us": {k: round(1e3 * v[1] / v[0], 1) for k, v in rep.items()}}), flush=True)
