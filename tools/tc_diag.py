"""Per-layer gradient error of each GEMM role on the tensor-core path vs the fp64 oracle (debugging aid).
usage (GPU box): python tools/tc_diag.py [batch]"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import lbfgs_ffnn_b200 as P
from oracle import binding as ob
from helpers import make_gpu_net, make_problem, upload
from conftest import rel_l2
h = P.CublasHandle(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
for dims, acts in [([784,128,10],["relu","linear"]), ([784,128,64,10],["relu","relu","linear"])]:
  onet,w,X,T = make_problem(ob,dims,acts,B)
  lo,go = onet.loss_grad(w,X,T)
  dx,dt = upload(X),upload(T)
  for prec in ["fp32","tf32x3","tf32"]:
    for mask in ([7] if prec=="fp32" else [1,9,2,4,15,7]):
      os.environ["B200_TC_MASK"]=str(mask); P.api.reload_env()
      net = make_gpu_net(h,dims,acts,w,precision=prec)
      l = net.compute_loss_and_grad(dx,dt,B); g = net.get_grads()
      print(dims, prec, mask, 'loss rel %.2e'%(abs(l-lo)/abs(lo)), 'grad rel %.2e'%rel_l2(g,go), flush=True)
