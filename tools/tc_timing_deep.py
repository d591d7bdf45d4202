"""Per-CTA clocks of the fp16 kernels of one evaluation of the deep net (and kernel span in ns), taken in steady state:
500 evaluations first (clocks ramped), then B200_TC_TIMING is switched on for three more."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lbfgs_ffnn_b200 as P
from helpers import make_gpu_net, upload
dims = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "784-128-64-10").split("-")]
B = 60000
h = P.CublasHandle(0)
X, T = P.synthetic_mnist(B)
dx, dt = upload(X), upload(T)
net = make_gpu_net(h, dims, ["relu"] * (len(dims) - 2) + ["linear"], None, precision="tf32x3")
net.quantize_input(dx, B)
for _ in range(500): net.loss_grad_async(dx, dt, B)
os.environ["B200_TC_TIMING"] = "1"
P.api.reload_env()
for _ in range(3): net.compute_loss_and_grad(dx, dt, B)
