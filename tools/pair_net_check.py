"""the stacked pair network (slbfgs.cu) evaluated from Python: gradient blocks vs two separate evaluations and vs the oracle"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lbfgs_ffnn_b200 as P
from oracle import binding as ob
from helpers import make_problem, make_gpu_net, upload
def rel(a, b): return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300))
dims, acts = [784, 128, 64, 10], ["relu", "relu", "linear"]
pd = [dims[0]] + [2 * d for d in dims[1:]]
h = P.CublasHandle(0)
def stack(wa, wb):
    out, oa = [], 0
    for l, (K, N) in enumerate(zip(dims[:-1], dims[1:])):
        Wa, Wb = wa[oa:oa + K * N].reshape(K, N), wb[oa:oa + K * N].reshape(K, N)
        ba, bb = wa[oa + K * N:oa + K * N + N], wb[oa + K * N:oa + K * N + N]
        if l == 0: Ws = np.concatenate([Wa, Wb], axis=1)
        else:
            Ws = np.zeros((2 * K, 2 * N), np.float32); Ws[:K, :N] = Wa; Ws[K:, N:] = Wb
        out += [Ws.ravel(), ba, bb]; oa += K * N + N
    return np.concatenate(out).astype(np.float32)
def unstack(g):
    ga, gb, op = [], [], 0
    for l, (K, N) in enumerate(zip(dims[:-1], dims[1:])):
        Kp = K if l == 0 else 2 * K
        G = g[op:op + Kp * 2 * N].reshape(Kp, 2 * N); b = g[op + Kp * 2 * N:op + Kp * 2 * N + 2 * N]
        ga += [G[:K, :N].ravel(), b[:N]]; gb += [(G[:, N:] if l == 0 else G[K:, N:]).ravel(), b[N:]]
        op += Kp * 2 * N + 2 * N
    return np.concatenate(ga), np.concatenate(gb)
for prec in ("fp32", "tf32x3"):
    for B in (1000, 5000):
        onet, w, X, T = make_problem(ob, dims, acts, B)
        rs = np.random.RandomState(3)
        s = (rs.standard_normal(w.size) * 1e-3).astype(np.float32)
        for eps in (1e-4, 1.6e-3):
            wa, wb = (w + np.float32(eps) * s).astype(np.float32), (w - np.float32(eps) * s).astype(np.float32)
            dx, dt, dt2 = upload(X), upload(T), upload(np.concatenate([T, T], axis=1))
            net = make_gpu_net(h, dims, acts, wa, precision=prec); net.set_l2(1e-4)
            net.compute_loss_and_grad(dx, dt, B); ga = net.get_grads()
            net.set_params(wb); net.compute_loss_and_grad(dx, dt, B); gb = net.get_grads()
            pnet = make_gpu_net(h, pd, acts, stack(wa, wb), precision=prec); pnet.set_l2(1e-4)
            pnet.compute_loss_and_grad(dx, dt2, B); pa, pb = unstack(pnet.get_grads())
            _, oa = onet.slbfgs_batch_grad(wa, X, T, None); _, obb = onet.slbfgs_batch_grad(wb, X, T, None)
            yo = (oa - obb) / (2 * eps)
            print(json.dumps(dict(prec=prec, B=B, eps=eps, ga_pair_vs_single=rel(pa, ga), gb_pair_vs_single=rel(pb, gb), ga_vs_oracle=rel(ga, oa),
                                  pa_vs_oracle=rel(pa, oa), y_single_vs_oracle=rel((ga.astype(np.float64) - gb) / (2 * eps), yo),
                                  y_pair_vs_oracle=rel((pa.astype(np.float64) - pb) / (2 * eps), yo), finite=bool(np.isfinite(pa).all() and np.isfinite(pb).all()))), flush=True)
            net.close(); pnet.close()
