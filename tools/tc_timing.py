"""B200_TC_TIMING=1 python tools/tc_timing.py : per-CTA clocks of every tcgen05 launch of one evaluation"""
import os, sys
os.environ["B200_TC_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lbfgs_ffnn_b200 as P
B = 60000
h = P.CublasHandle(0)
X, T = P.synthetic_mnist(B)
dx, dt = P.DeviceBuffer(), P.DeviceBuffer(); dx.copy_from_host(X); dt.copy_from_host(T)
for prec in ("tf32", "tf32x3"):
    for q in (True,):
        net = P.CudaNetwork(h)
        net.addLayer(784, 128, "relu"); net.addLayer(128, 10, "linear")
        net.bindParams(123); net.set_precision(prec)
        if q: net.quantize_input(dx, B)
        net.compute_loss_and_grad(dx, dt, B)
        print(f"--- {prec} u8={q}", file=sys.stderr, flush=True)
        net.compute_loss_and_grad(dx, dt, B)
