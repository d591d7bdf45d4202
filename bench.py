#!/usr/bin/env python
"""bench.py — L-BFGS iterations/s on the BASELINE.json workload, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is ONE L-BFGS iteration of the reference's CUDA algorithm (src/cuda/lbfgs.cuh:90-186: direction,
Armijo line search with all of its loss+gradient evaluations, history update) on the full batch.
Workload (config.workload): BASELINE.json configs[1] — 784-128-10 MLP (ReLU, Linear), 60 000 synthetic
MNIST-shaped samples, m = 10, fp32-accurate arithmetic. For N > 1 the 60 000 samples are sharded over the ranks
and the flat gradient (+ loss) is all-reduced with NCCL once per evaluation ("strong" scaling: total work fixed).

value : iterations/s with X, T and the parameters already resident in HBM (CUDA events on the library's stream,
        max over ranks). The per-iteration working set (190 MB: see config.l2) exceeds L2 (126 MB); no explicit flush.
e2e   : the same K iterations through the public C-ABI solve call starting from PINNED HOST buffers: the timed
        region contains the H2D copy of X, T and the parameters, b200_lbfgs_solve (which returns every
        iteration's loss / gradient norm to the host), and the D2H copy of the final parameters.
roofline: dominant kernel class of the timed step, timed per launch with CUDA events (b200_ctx_profile) in a
        second pass over the same K iterations.
cpu_baseline / --impl reference: the oracle's fp64 restatement of the reference CPU path
        (src/minimizer/lbfgs.hpp weak-Wolfe L-BFGS on src/network.hpp) on the host cores (kind "port": the
        reference's own CPU build needs Eigen, which this image does not have).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = [784, 128, 10]
ACTS = ["relu", "linear"]
TOTAL_SAMPLES = 60000
MEMORY = 10
FLOP_PER_SAMPLE = 409088  # 2*(2*(784*128 + 128*10) + 128*10)
WORKLOAD = "lbfgs_m10_mlp784-128-10_B60000_fullbatch"
CPU_SAMPLE = 6000  # samples per CPU-baseline step (1/10 of the workload; time scaled by 10)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_lbfgs_rate(steps, warmup, params32, X, T):
    """reference CPU algorithm (oracle port) on a CPU_SAMPLE-sample slice; returns (it/s scaled to 60k, info)"""
    from oracle import binding as ob
    ob.build()
    net = ob.OracleNet(DIMS, ACTS)
    r = net.lbfgs(params32, X[:CPU_SAMPLE], T[:CPU_SAMPLE], m=MEMORY, max_iters=warmup + steps, tol=0.0, policy="cpu")
    ms = r["ms"]
    t = (ms[warmup + steps - 1] - (ms[warmup - 1] if warmup > 0 else 0.0)) / 1e3
    scale = TOTAL_SAMPLES / CPU_SAMPLE
    rate = steps / (t * scale)
    evals = (r["n_f"] + r["n_g"]) / max(1, r["iters"])
    info = dict(cores=ob.num_threads(), kind="port",
                sample=f"{steps} L-BFGS iterations (after {warmup} warm-up) of the reference CPU algorithm (weak-Wolfe, fp64, "
                       f"{evals:.1f} objective calls/iteration) on {CPU_SAMPLE} of the {TOTAL_SAMPLES} samples; time scaled x{scale:.0f}",
                seconds=t)
    return rate, info


def run_reference(args, rank):
    if rank != 0:
        return
    import lbfgs_ffnn_b200 as P
    from oracle import binding as ob
    ob.build()
    X, T = P.synthetic_mnist(CPU_SAMPLE)
    w = ob.OracleNet(DIMS, ACTS).init_params_cuda_rule(123)
    rate, info = cpu_lbfgs_rate(args.steps, args.warmup, w, X, T)
    line = {"impl": "reference", "metric": "lbfgs_iters_per_sec", "value": rate, "unit": "iterations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / rate, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "line_search": "weak-wolfe (reference CPU backend)", "memory": MEMORY},
            "cpu_baseline": dict(value=rate, unit="iterations/s", **{k: info[k] for k in ("cores", "kind", "sample")}),
            "e2e": {"value": rate, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--precision", default="tf32x3", choices=["fp32", "tf32x3", "tf32"],
                    help="tf32x3 (default): fp32-accurate hi/lo split on the tensor cores; fp32: FFMA; tf32: single pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--net", default=None,
                    help="layer widths, e.g. 784-128-64-10 (BASELINE.json configs[2], the deep net); default: configs[1], the headline")
    args = ap.parse_args()
    if args.net:
        global DIMS, ACTS, FLOP_PER_SAMPLE, WORKLOAD
        DIMS = [int(v) for v in args.net.split("-")]
        ACTS = ["relu"] * (len(DIMS) - 2) + ["linear"]
        pairs = list(zip(DIMS[:-1], DIMS[1:]))
        FLOP_PER_SAMPLE = 2 * (2 * sum(a * b for a, b in pairs) + sum(a * b for a, b in pairs[1:]))
        WORKLOAD = f"lbfgs_m{MEMORY}_mlp{args.net}_B{TOTAL_SAMPLES}_fullbatch"
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import lbfgs_ffnn_b200 as P

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic data: identical on every rank (same seed); each rank keeps its contiguous shard ------------
    assert TOTAL_SAMPLES % world == 0
    shard = TOTAL_SAMPLES // world
    Xh, Th = P.synthetic_mnist(TOTAL_SAMPLES)
    Xs = torch.from_numpy(Xh[rank * shard:(rank + 1) * shard]).pin_memory()
    Ts = torch.from_numpy(Th[rank * shard:(rank + 1) * shard]).pin_memory()

    h = P.CublasHandle(local_rank)
    stream = torch.cuda.Stream()
    h.set_stream(stream.cuda_stream)  # the library launches on this torch stream, so torch.cuda.Event sees it
    if world > 1:
        uid = [P.CublasHandle.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.init_comm(uid[0], rank, world)
    net = P.CudaNetwork(h)
    for i, a in enumerate(ACTS):
        net.addLayer(DIMS[i], DIMS[i + 1], a)
    net.bindParams(123)  # CudaNetwork::bindParams rule, identical on every rank
    net.set_precision(args.precision)
    net.set_global_batch(TOTAL_SAMPLES)
    n = net.params_size()
    w0 = net.get_params()
    w0_pinned = torch.from_numpy(w0.copy()).pin_memory()

    with torch.cuda.stream(stream):
        dX = torch.empty_like(Xs, device="cuda")
        dT = torch.empty_like(Ts, device="cuda")
        dX.copy_(Xs, non_blocking=True)
        dT.copy_(Ts, non_blocking=True)
    stream.synchronize()

    def new_solver(iters):
        s = P.CudaLBFGS(h)
        s.setMemory(MEMORY); s.setMaxIterations(iters); s.setTolerance(0.0)
        return s

    def timed_iterations(profile=False):
        """W warm-up iterations then exactly K timed iterations of ONE continuing minimisation"""
        net.set_params(w0)
        s = new_solver(args.warmup + args.steps)
        rec = P.IterationRecorder(); rec.init(args.steps + args.warmup)
        s.setRecorder(rec)
        s.begin(n)
        s.run(net.params_data(), dX, dT, shard, args.warmup, net)
        if profile:
            h.profile(True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = P.api.launch_count()
        e0.record(stream)
        s.run(net.params_data(), dX, dT, shard, args.steps, net)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        rep = h.profile_report() if profile else None
        if profile:
            h.profile(False)
        loss, gn, _ = rec.copy_to_host()
        out = dict(ms=ms, launches=P.api.launch_count() - launches0, evals=s.last_evaluations_, iters=s.iterations(),
                   loss_first=float(loss[0]) if loss.size else None, loss_last=float(loss[-1]) if loss.size else None,
                   report=rep)
        s.end()
        return out

    # ---- value: device-resident ---------------------------------------------------------------------------------
    timed_iterations()  # untimed pass: allocations, module load
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res = timed_iterations()
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([res["ms"]], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_total = float(t_ms.item())
    assert res["iters"] == args.steps, f"expected {args.steps} timed iterations, got {res['iters']}"

    # ---- e2e: pinned host buffers -> C-ABI solve -> host ---------------------------------------------------------
    def e2e_once(iters):
        w_out = torch.empty(n, dtype=torch.float32).pin_memory()
        s = new_solver(iters)
        rec = P.IterationRecorder(); rec.init(iters); s.setRecorder(rec)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        with torch.cuda.stream(stream):
            dX.copy_(Xs, non_blocking=True)
            dT.copy_(Ts, non_blocking=True)
            wd = torch.empty(n, dtype=torch.float32, device="cuda")
            wd.copy_(w0_pinned, non_blocking=True)
        s.solve(n, wd, dX, dT, shard, net)  # evaluates at the caller's buffer wd (aliasing contract)
        with torch.cuda.stream(stream):
            w_out.copy_(wd, non_blocking=True)
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), s.iterations(), rec

    e2e_once(args.warmup)
    e2e_ms, e2e_iters, e2e_rec = e2e_once(args.steps)
    t2 = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2.item())
    h2d = (Xs.numel() + Ts.numel() + n) * 4
    d2h = n * 4 + 16 * res["evals"]  # final parameters + (loss, ||g||^2) per evaluation

    # ---- roofline pass: per-launch CUDA events ---------------------------------------------------------------------
    prof = timed_iterations(profile=True)
    rep = prof["report"] or {}
    pk = peaks()
    roofline, kernels = None, {}
    rooflines = {}
    if rep:
        total_prof = sum(v[1] for v in rep.values())
        for k, (calls, tot) in rep.items():
            kernels[k] = {"launches": calls, "avg_us": 1e3 * tot / calls, "share": tot / total_prof}
        # Algorithmic work per launch of the kernels with a roof (shard = samples on this GPU, fp32-accurate mode):
        #   fwd0  layer-0 forward: 2*B*784*128 flop; reads X (exact fp16 copy when the input is 8-bit pixels, else fp32) + writes A1
        #   dw0   layer-0 [dW; db]: 2*B*785*128 flop; reads X + delta_0 (fp16 hi|lo = 4 B/element, or fp32) + writes the split-K partials
        #   tail_fwd / tail_bwd  last layer in two passes: read A1 (+ write delta_0): HBM class
        #   lbfgs_direction  two-loop recursion, (4k+2)*n*4 bytes
        B, K0, N0 = shard, DIMS[0], DIMS[1]
        u8 = args.precision != "fp32"
        x_bytes = B * ((K0 + 1 + 7) // 8 * 8) * 2 if u8 else B * K0 * 4  # 8-bit pixels: the fp16 copy [in | 1 | pad] the GEMMs read
        work = {
            "fwd0": dict(flops=2.0 * B * K0 * N0, bytes=x_bytes + 4.0 * B * N0),
            "dw0": dict(flops=2.0 * B * (K0 + 1) * N0, bytes=x_bytes + 4.0 * B * N0),
            "tail_fwd": dict(flops=2.0 * B * DIMS[-2] * DIMS[-1], bytes=4.0 * B * (DIMS[-2] + 3 * DIMS[-1])),
            "tail_bwd": dict(flops=4.0 * B * DIMS[-2] * DIMS[-1], bytes=4.0 * B * (2 * DIMS[-2] + DIMS[-1])),
            "lbfgs_direction": dict(flops=0.0, bytes=(4.0 * MEMORY + 2) * n * 4),
        }
        tensor_peak = pk["bf16_sustained"]  # kind::f16 MMAs: the measured dense 16-bit figure (TF32 is half of it)
        for k, w in work.items():
            if k not in rep:
                continue
            calls, tot = rep[k]
            avg_s = tot / calls / 1e3
            t_hbm = w["bytes"] / (pk["hbm"] * 1e9)
            t_tc = w["flops"] / (tensor_peak * 1e12)
            if t_tc > t_hbm:
                ach = w["flops"] / avg_s / 1e12
                rooflines[k] = {"bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak}
            else:
                ach = w["bytes"] / avg_s / 1e9
                rooflines[k] = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
            rooflines[k].update({"avg_launch_us": avg_s * 1e6, "share_of_step": tot / total_prof, "alg_flops": w["flops"],
                                 "alg_bytes": w["bytes"], "tensor_TFLOPs": w["flops"] / avg_s / 1e12})
        # DRAM bytes per launch from the committed `ncu --set full` capture of this command (profiles/r01_traffic.json)
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath) and world == 1 and args.precision == "tf32x3" and not args.net:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch", {})
        for k in rooflines:
            rooflines[k]["traffic"] = traffic.get(k)
        if rooflines:
            dom = max(rooflines, key=lambda k: rooflines[k]["share_of_step"])
            roofline = dict(kernel=dom, **rooflines[dom])
            roofline["peak_note"] = (f"{pk['source']} peaks: HBM {pk['hbm']} GB/s, dense 16-bit tensor {tensor_peak} TFLOP/s sustained; the bound is "
                                     f"the larger of bytes/HBM and flops/tensor for the kernel's ALGORITHMIC work; precision mode {args.precision}; "
                                     "lbfgs_direction streams an L2-resident history at this size (latency-bound)")
        if "lbfgs_dots" in rep and "lbfgs_apply" in rep:  # unfused direction (B200_NO_FUSED_DIRECTION)
            t_dir = (rep["lbfgs_dots"][1] / rep["lbfgs_dots"][0] + rep["lbfgs_solve"][1] / rep["lbfgs_solve"][0] +
                     rep["lbfgs_apply"][1] / rep["lbfgs_apply"][0]) / 1e3
            bytes_dir = (4 * MEMORY + 2) * n * 4
            kernels["direction"] = {"avg_us": t_dir * 1e6, "achieved_GBps": bytes_dir / t_dir / 1e9,
                                    "frac_of_hbm": bytes_dir / t_dir / 1e9 / pk["hbm"],
                                    "note": "history (8 MB) is L2-resident at this size; latency-bound"}

    if rank == 0:
        value = args.steps / (ms_total / 1e3)
        line = {"metric": "lbfgs_iters_per_sec", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "f32 (fp32-accurate split products on the tensor cores: fp16 hi+lo x exact uint8 pixels, fp32 accumulate)", "tf32": "tf32 / fp16 single-pass operands"}[args.precision],
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "memory": MEMORY, "line_search": "armijo (reference CUDA backend)",
                           "precision": args.precision, "samples_per_gpu": shard, "params": n,
                           "l2": "no flush: iterations run back to back inside one solver call; per-iteration working set at 60 000 samples = X 95 MB (exact fp16 copy of the 8-bit pixels; 188 MB as fp32) + A1 31 MB + delta 31 MB + split-K partials 16 MB + history 9 MB + T/outputs 8 MB = 190 MB > 126 MB L2",
                           "evals_per_iteration": res["evals"] / args.steps,
                           "parallelism": f"samples sharded x{world}, NCCL allreduce of grad+loss" if world > 1 else "1 GPU"},
                "gpu_launches": res["launches"],
                "loss": {"first_timed": res["loss_first"], "last_timed": res["loss_last"]},
                "clocks": clocks,
                "e2e": {"value": e2e_iters / (e2e_ms / 1e3), "unit": "iterations/s", "h2d_bytes_per_step": h2d / max(1, e2e_iters),
                        "d2h_bytes_per_step": d2h / max(1, e2e_iters), "ms_total": e2e_ms, "iterations": e2e_iters,
                        "note": "H2D of X,T,params from pinned memory + b200_lbfgs_solve + D2H of params inside the timed region"},
                "roofline": roofline, "rooflines": rooflines, "kernels": kernels}
        if world == 1 and not args.no_cpu_baseline:
            k_cpu = max(3, min(args.steps, 20))
            rate, info = cpu_lbfgs_rate(k_cpu, 2, w0, Xh, Th)
            line["cpu_baseline"] = dict(value=rate, unit="iterations/s", **{k: info[k] for k in ("cores", "kind", "sample")})
        print(json.dumps(line))
    net.close()
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
