"""torchrun --nproc-per-node N tools/multi_gpu_check.py : N-rank sample-sharded L-BFGS (NCCL allreduce of grad + loss)
vs the same problem on one GPU (rank 0). Prints one JSON line on rank 0."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import lbfgs_ffnn_b200 as P

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
B, iters = 6000, 25
dims, acts = [784, 128, 64, 10], ["relu", "relu", "linear"]
X, T = P.synthetic_mnist(B)


def run(handle, xs, ts, nb, global_b, shard=0):
    net = P.CudaNetwork(handle)
    for i, a in enumerate(acts):
        net.addLayer(dims[i], dims[i + 1], a)
    net.bindParams(123)
    net.set_global_batch(global_b)
    dx, dt = P.DeviceBuffer(), P.DeviceBuffer()
    dx.copy_from_host(xs); dt.copy_from_host(ts)
    s = P.CudaLBFGS(handle)
    s.setMemory(10); s.setMaxIterations(iters); s.setTolerance(0.0); s.setShardHistory(shard)
    rec = P.IterationRecorder(); rec.init(iters); s.setRecorder(rec)
    s.solve(net.params_size(), net.params_data(), dx, dt, nb, net)
    return rec.copy_to_host()[0], net.get_params()


h = P.CublasHandle(lr)
uid = [P.CublasHandle.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
h.init_comm(uid[0], rank, world)
shard = B // world
loss_multi, w_multi = run(h, X[rank * shard:(rank + 1) * shard], T[rank * shard:(rank + 1) * shard], shard, B)
dist.barrier()
# history sharded by parameter index: reduce-scatter of the gradient, all-reduce of the partial dot products only
loss_sh, w_sh = run(h, X[rank * shard:(rank + 1) * shard], T[rank * shard:(rank + 1) * shard], shard, B, shard=1)
dist.barrier()
if rank == 0:
    h1 = P.CublasHandle(lr)
    loss_single, w_single = run(h1, X, T, B, B)
    d = float(np.max(np.abs(loss_multi - loss_single) / np.abs(loss_single)))
    pr = float(np.linalg.norm(w_multi - w_single) / np.linalg.norm(w_single))
    d5 = float(np.max(np.abs(loss_multi[:5] - loss_single[:5]) / np.abs(loss_single[:5])))
    ds = float(np.max(np.abs(loss_sh - loss_single) / np.abs(loss_single)))
    ds5 = float(np.max(np.abs(loss_sh[:5] - loss_single[:5]) / np.abs(loss_single[:5])))
    prs = float(np.linalg.norm(w_sh - w_single) / np.linalg.norm(w_single))
    print(json.dumps({"world": world, "iters": iters, "sharded_max_rel_loss_diff": ds, "sharded_max_rel_loss_diff_first5": ds5,
                      "sharded_params_rel_l2": prs, "loss_sharded_last": float(loss_sh[-1]), "max_rel_loss_diff": d, "max_rel_loss_diff_first5": d5, "params_rel_l2": pr,
                      "loss_multi_last": float(loss_multi[-1]), "loss_single_last": float(loss_single[-1])}))
dist.barrier()
dist.destroy_process_group()
