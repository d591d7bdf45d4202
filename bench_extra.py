"""Secondary bench lines of the same hot path (python bench.py --config slbfgs|gd|sgd|c5). Same conventions as bench.py: one JSON
line from rank 0, device time from CUDA events on the library's stream (max over ranks), synthetic data, W >= 3 warm-up steps.

  slbfgs  BASELINE configs[3]: S-LBFGS (SVRG + finite-difference HVP pairs) on 784-128-64-10, 60 000 samples, mini-batch 1000,
          b_H = 5000, M = 10, L = 10. A step is one EPOCH (full gradient + 60 inner steps + curvature pairs). metric: epochs/s.
  gd      CudaGD (src/cuda/gd.cuh) on the default network, momentum 0.9: iterations/s.
  sgd     CudaSGD (src/cuda/sgd.cuh), mini-batch 256 as in tests/mnist/main-gpu.cpp: epochs/s.
  c5      BASELINE configs[4]: L-BFGS m = 20 on 784-4096-4096-10 (20 037 642 parameters), 1 000 000 samples sharded over the ranks
          (all on one GPU at N = 1), history sharded by parameter index when N > 1. Reports the direction's achieved HBM GB/s
          and every GEMM's TFLOP/s next to the measured peaks.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def run(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import lbfgs_ffnn_b200 as P
    import bench as B

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def tmax(ms):
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cfg = args.config
    if cfg == "c5":
        dims, acts, total, memory = [784, 4096, 4096, 10], ["relu", "relu", "linear"], 1000000, 20
    else:
        dims, acts, total, memory = list(B.DIMS), list(B.ACTS), 60000, 10
    if args.samples:
        total = args.samples
    assert total % world == 0
    shard = total // world
    steps = args.steps if args.steps != 200 else {"slbfgs": 10, "gd": 200, "sgd": 10, "c5": 5}[cfg]
    # c5: the ring of m = 20 pairs is full after 20 iterations; the direction's bytes (4m + 2 vectors) are only streamed from then on
    warmup = max(3, args.warmup if args.warmup != 10 else {"slbfgs": 3, "gd": 10, "sgd": 3, "c5": 21}[cfg])

    h = P.CublasHandle(local_rank)
    stream = torch.cuda.Stream()
    h.set_stream(stream.cuda_stream)
    if world > 1:
        uid = [P.CublasHandle.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.init_comm(uid[0], rank, world)

    # S-LBFGS keeps the whole data set on every rank (each mini-batch's index list is split over the ranks); the others shard
    keep = total if cfg == "slbfgs" else shard
    off = 0 if cfg == "slbfgs" else rank * shard
    t0 = time.time()
    if cfg == "c5":  # only this rank's shard is generated (the stream of the full set would take minutes at 1e6 samples)
        Xh, Th = P.synthetic_mnist(shard, seed=123 + rank)
    else:
        Xa, Ta = P.synthetic_mnist(total)
        Xh, Th = Xa[off:off + keep], Ta[off:off + keep]
    gen_s = time.time() - t0
    with torch.cuda.stream(stream):
        dX = torch.from_numpy(np.ascontiguousarray(Xh)).to("cuda", non_blocking=False)
        dT = torch.from_numpy(np.ascontiguousarray(Th)).to("cuda", non_blocking=False)
    stream.synchronize()

    net = P.CudaNetwork(h)
    for i, a in enumerate(acts):
        net.addLayer(dims[i], dims[i + 1], a)
    net.bindParams(123)
    net.set_precision(args.precision)
    if cfg != "slbfgs":
        net.set_global_batch(total)
    n = net.params_size()
    w0 = net.get_params()
    pk = B.peaks()

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = P.api.launch_count()
        e0.record(stream)
        out = fn()
        e1.record(stream)
        barrier()
        return tmax(e0.elapsed_time(e1)), P.api.launch_count() - l0, out

    line = {"n_gpus": world, "steps": steps, "warmup": warmup, "higher_is_better": True, "vs_baseline": None, "data": "synthetic",
            "dtype": "f32 (fp32-accurate split products on the tensor cores)" if args.precision == "tf32x3" else args.precision}

    if cfg == "slbfgs":
        def solve(epochs, record=True):
            net.set_params(w0)
            s = P.CudaSLBFGS(h)
            s.setMaxIterations(epochs); s.setTolerance(0.0); s.setStepSize(0.02); s.setBatchSize(1000)
            s.setMemory(10); s.setUpdateInterval(10); s.setHessianBatchSize(5000)
            rec = P.IterationRecorder(); rec.init(epochs)
            if record:
                s.setRecorder(rec)
            s.solve(n, net.params_data(), dX, dT, total, net)
            return rec.copy_to_host()[0], s.last_evaluations_
        solve(warmup)
        ms, launches, (loss, evals) = timed(lambda: solve(steps))
        line.update({"metric": "slbfgs_epochs_per_sec", "value": steps / (ms / 1e3), "unit": "epochs/s", "ms_per_step": ms / steps,
                     "scaling": "strong", "gpu_launches": launches,
                     "config": {"workload": "slbfgs_mlp784-128-64-10_N60000_b1000_bH5000_M10_L10_eta0.02", "precision": args.precision,
                                "launches_per_epoch": launches / steps, "evaluations_per_epoch": evals / steps,
                                "loss_first_epoch": float(loss[0]), "loss_last_epoch": float(loss[-1]),
                                "step": "one epoch = full gradient at the anchor + 60 variance-reduced inner steps (two mini-batch "
                                        "gradients each) + curvature pairs every 10 steps (two H-batch gradients each) + the recorder's "
                                        "full loss / gradient"}})
        if rank == 0 and not args.no_cpu_baseline:
            try:
                from oracle import ref_cpu_binding as rc
                cores = os.cpu_count() or 1
                t1 = time.time()
                if rc.available():
                    rc.set_num_threads(cores)
                    devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
                    sys.stdout.flush(); os.dup2(devnull, 1)
                    try:
                        rc.RefCpuNet(dims).train("slbfgs", w0, Xh, Th, max_iters=1, tolerance=0.0, learning_rate=0.02, batch_size=1000,
                                                 m_param=10, L_param=10, b_H_param=5000, log_interval=1)
                    finally:
                        sys.stdout.flush(); os.dup2(saved, 1); os.close(devnull); os.close(saved)
                    kind, threads = "reference", rc.num_threads()
                else:
                    from oracle import binding as ob
                    ob.build(); ob.set_num_threads(cores)
                    ob.OracleNet(dims, acts).slbfgs(w0, Xh, Th, batch_size=1000, M=10, L=10, b_H=5000, step=0.02, max_iters=1, tol=0.0)
                    kind, threads = "port", ob.num_threads()
                dt = time.time() - t1
                line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "epochs/s", "cores": threads, "kind": kind,
                                        "sample": "ONE epoch of UnifiedSLBFGS<CpuBackend> on the same 60 000 samples and options "
                                                  "(wall clock around the train call, includes its history CSV)"}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "error": str(e)[:200]}
    elif cfg in ("gd", "sgd"):
        def solve(iters):
            net.set_params(w0)
            if cfg == "gd":
                s = P.CudaGD(h); s.setLearningRate(0.01); s.setMomentum(0.9)
            else:
                s = P.CudaSGD(h); s.setLearningRate(0.01); s.setMomentum(0.9); s.setBatchSize(256); s.setDimensions(dims[0], dims[-1])
            s.setMaxIterations(iters); s.setTolerance(0.0)
            rec = P.IterationRecorder(); rec.init(iters + 1); s.setRecorder(rec)
            s.solve(n, net.params_data(), dX, dT, shard, net)
            return rec.copy_to_host()[0], s.last_evaluations_
        solve(warmup)
        ms, launches, (loss, evals) = timed(lambda: solve(steps))
        unit = "iterations/s" if cfg == "gd" else "epochs/s"
        line.update({"metric": f"{cfg}_{'iters' if cfg == 'gd' else 'epochs'}_per_sec", "value": steps / (ms / 1e3), "unit": unit,
                     "ms_per_step": ms / steps, "scaling": "strong", "gpu_launches": launches,
                     "config": {"workload": f"{cfg}_mlp{'-'.join(map(str, dims))}_B{total}" + ("_b256" if cfg == "sgd" else ""),
                                "precision": args.precision, "evaluations_per_step": evals / steps, "loss_first": float(loss[0]),
                                "loss_last": float(loss[-1]), "momentum": 0.9, "lr": 0.01}})
        if rank == 0 and world == 1 and not args.no_reference_cuda:
            try:
                from oracle import ref_cuda_binding as rcu
                if rcu.available():
                    torch.cuda.synchronize()
                    r = rcu.RefCudaNet(dims, [{"linear": 0, "tanh": 1, "relu": 2, "sigmoid": 3}[a] for a in acts])
                    kw = dict(lr=0.01, momentum=0.9) if cfg == "gd" else dict(lr=0.01, momentum=0.9, sgd_batch=256)
                    r.bind_params(123); r.solve(cfg, dX.data_ptr(), dT.data_ptr(), shard, 3, **kw, record=True)
                    r.bind_params(123)
                    it = steps if cfg == "gd" else min(steps, 5)
                    o = r.solve(cfg, dX.data_ptr(), dT.data_ptr(), shard, it, **kw, record=True)
                    line["reference_cuda"] = {"value": o["iters"] / (o["total_ms"] / 1e3), "unit": unit,
                                              "sample": f"{it} steps of the reference CUDA backend (cuBLAS) on the same GPU and workload"}
                    r.close()
            except Exception as e:
                line["reference_cuda"] = {"value": None, "error": str(e)[:200]}
    else:  # c5
        def solver(iters):
            s = P.CudaLBFGS(h)
            s.setMemory(memory); s.setMaxIterations(iters); s.setTolerance(0.0)
            return s
        net.set_params(w0)
        s = solver(warmup + steps)
        rec = P.IterationRecorder(); rec.init(warmup + steps); s.setRecorder(rec)
        s.begin(n)
        s.run(net.params_data(), dX, dT, shard, warmup, net)
        ms, launches, _ = timed(lambda: s.run(net.params_data(), dX, dT, shard, steps, net))
        evals_timed = s.last_evaluations_
        loss_timed = rec.copy_to_host()[0]
        h.profile(True)
        s.run(net.params_data(), dX, dT, shard, min(steps, 3), net)
        rep = h.profile_report()
        h.profile(False)
        prof_iters = s.iterations()
        s.end()
        Bs = shard
        flops = {"fwd0": 2.0 * Bs * 784 * 4096, "fwd1": 2.0 * Bs * 4096 * 4096, "dx1": 2.0 * Bs * 4096 * 4096,
                 "dw1": 2.0 * Bs * 4097 * 4096, "dw0": 2.0 * Bs * 785 * 4096}
        tensor_peak = pk["bf16_sustained"]
        wide16 = args.precision == "tf32x3" and os.environ.get("B200_WIDE16", "1") != "0"
        kern, roofs = {}, {}
        tot_prof = sum(v[1] for v in rep.values()) or 1.0
        for k, (calls, tot) in rep.items():
            avg_s = tot / calls / 1e3
            kern[k] = {"launches": calls, "avg_us": avg_s * 1e6, "share": tot / tot_prof}
            if k in flops:
                ach = flops[k] / avg_s / 1e12
                roofs[k] = {"bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak,
                            "avg_launch_us": avg_s * 1e6, "share_of_step": tot / tot_prof, "alg_flops": flops[k]}
                if wide16:
                    # fp32-accurate mode on 16-bit tensor cores: every algorithmic product is issued as four fp16 products
                    # (hi hi, hi lo, lo hi, lo lo: two M = 256, N = 256, K = 16 MMAs per K step), so the tensor pipe does 4x the
                    # algorithmic flops; the scope's time also holds the operand splits (one pass over each fp32 matrix)
                    roofs[k]["issued_flops"] = 4.0 * flops[k]
                    roofs[k]["tensor_pipe_frac"] = 4.0 * ach / tensor_peak
                    roofs[k]["note"] = ("wide16: fp16 pair operands, CTA-pair tcgen05 MMAs; frac = algorithmic flops / 16-bit dense peak "
                                        "(ceiling 0.25 in this mode), tensor_pipe_frac = issued flops / peak")
        # direction: (4k+2) * n_local * 4 bytes over the rank's slice of the history, k = m once the ring is full
        n_local = n if world == 1 else (n + world - 1) // world
        dir_bytes = (4.0 * memory + 2) * n_local * 4
        dir_keys = [k for k in ("lbfgs_direction", "lbfgs_dots", "lbfgs_solve", "lbfgs_apply") if k in rep]
        if dir_keys:
            t_dir = sum(rep[k][1] / rep[k][0] for k in dir_keys) / 1e3
            ach = dir_bytes / t_dir / 1e9
            roofs["lbfgs_direction"] = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                                        "avg_launch_us": t_dir * 1e6, "alg_bytes": dir_bytes, "kernels": dir_keys,
                                        "share_of_step": sum(rep[k][1] for k in dir_keys) / tot_prof,
                                        "ring_full": bool(warmup + steps >= memory),
                                        # the two passes on their own: history pass (2m rows + g, x, x_prev, g_prev read, the new
                                        # pair written), output pass (2m rows + g, x read; p, x_prev, x written)
                                        "passes": {k: {"avg_launch_us": 1e3 * rep[k][1] / rep[k][0], "alg_bytes": b * n_local * 4,
                                                       "GB/s": b * n_local * 4 / (rep[k][1] / rep[k][0] / 1e3) / 1e9}
                                                   for k, b in (("lbfgs_dots", 2.0 * memory + 6), ("lbfgs_apply", 2.0 * memory + 5)) if k in rep},
                                        "note": "history ring of k = m pairs streamed twice (Gram / projection pass, output pass); "
                                                "the bytes assume a full ring (warm-up >= m iterations: ring_full) — with fewer "
                                                "pairs in the ring the passes move fewer bytes than alg_bytes and the rate is overstated"}
            # measured DRAM bytes of the two passes (one `ncu --set full` capture of this workload's direction on one GPU, committed with
            # the hash of the kernel source it was taken from): reported only while the source still hashes to that and the ring is full
            roofs["lbfgs_direction"]["traffic"] = None
            try:
                import hashlib
                cap = json.load(open(os.path.join(ROOT, "profiles", "r02_c5_direction_ncu_full_summary.json")))
                ok = all(hashlib.sha1(open(os.path.join(ROOT, f), "rb").read()).hexdigest()[:12] == h for f, h in cap.get("sources", []))
                if ok and cap.get("sources") and world == 1 and warmup + steps >= memory:
                    per = {}
                    for e in cap["launches"]:
                        per.setdefault("lbfgs_dots" if "dots" in e["kernel"] else "lbfgs_apply", []).append(e["dram_bytes"])
                    roofs["lbfgs_direction"]["traffic"] = int(sum(sum(v) / len(v) for v in per.values()))
                    roofs["lbfgs_direction"]["traffic_source"] = "profiles/r02_c5_direction_ncu_full_summary.json"
            except Exception:
                pass
        dom = max(roofs, key=lambda k: roofs[k]["share_of_step"]) if roofs else None
        line.update({"metric": "lbfgs_iters_per_sec", "value": steps / (ms / 1e3), "unit": "iterations/s", "ms_per_step": ms / steps,
                     "scaling": "strong", "gpu_launches": launches,
                     "config": {"workload": f"lbfgs_m{memory}_mlp784-4096-4096-10_B{total}_fullbatch", "params": n, "precision": args.precision,
                                "samples_per_gpu": shard, "evals_per_iteration": evals_timed / steps,
                                "loss_first_timed": float(loss_timed[0]) if loss_timed.size else None,
                                "loss_last_timed": float(loss_timed[-1]) if loss_timed.size else None,
                                "history": "sharded by parameter index: gradient reduce-scatter, 5(m+1)+1 partial dots all-reduced, "
                                           "parameters all-gathered (NCCL)" if world > 1 else "one GPU: replicated = whole",
                                "flop_per_evaluation": B.flop_per_sample(dims) * total,
                                "data_note": f"each rank generates its own {shard}-sample shard (seed 123 + rank), {gen_s:.1f} s on the host",
                                "profiled_iterations": prof_iters},
                     "roofline": dict(kernel=dom, **roofs[dom]) if dom else None, "rooflines": roofs, "kernels": kern,
                     "cpu_baseline": {"value": None, "unit": "iterations/s", "kind": "reference", "cores": os.cpu_count(),
                                      "sample": "not run at this size: one evaluation is 113.75 TFLOP in fp64 on the host"}})
    if rank == 0:
        print(json.dumps(line))
    net.close()
    h.close()
    if world > 1:
        dist.destroy_process_group()
