#!/usr/bin/env python
"""bench.py — L-BFGS iterations/s on the BASELINE.json workload, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--net 784-128-10] [--config lbfgs|slbfgs|gd|sgd|c5]

A "step" is ONE L-BFGS iteration of the reference's CUDA algorithm (src/cuda/lbfgs.cuh:90-186: direction, Armijo line search
with all of its loss+gradient evaluations, history update) on the full batch.
Default workload (config.workload): BASELINE.json configs[2], the north-star target — 784-128-64-10 MLP (ReLU, ReLU, Linear;
tests/fashion-mnist/main_gpu_deep.cpp in spirit), 60 000 synthetic MNIST-shaped samples, m = 10, fp32-accurate arithmetic.
`--net 784-128-10` is configs[1]. For N > 1 the 60 000 samples are sharded over the ranks and the flat gradient (+ loss) is
all-reduced once per evaluation over NVLink peer memory ("strong" scaling: total work fixed).

value : iterations/s with X, T and the parameters already resident in HBM (CUDA events on the library's stream, max over ranks).
e2e   : the same K iterations through the public C-ABI solve call starting from PINNED HOST buffers: the timed region contains
        the H2D copy of X, T and the parameters, b200_lbfgs_solve (which returns every iteration's loss / gradient norm to the
        host), and the D2H copy of the final parameters.
roofline: dominant kernel class of the timed step, timed per launch with CUDA events (b200_ctx_profile) in a second pass over
        the same K iterations; `rooflines` has every class with a roof.
cpu_baseline / --impl reference: the reference's OWN CPU path (src/minimizer/lbfgs.hpp weak-Wolfe L-BFGS on src/network.hpp,
        through run_full_batch_cpu) compiled unmodified against an Eigen-API stand-in (oracle/_ref/libref_cpu.so, kind "reference");
        where that build is missing, the oracle's fp64 restatement of it (kind "port"). FULL 60 000 samples, all host threads.
reference_cuda: the reference's CUDA backend (cuBLAS SGEMM + src/cuda/*.cuh, oracle/_ref/libref_cuda.so) on the same GPU, same
        workload — the comparator BASELINE.md names.
Other --config values are secondary bench lines of the same path (S-LBFGS epochs/s on configs[3]; GD iterations/s and SGD
epochs/s; c5 = the per-GPU share of configs[4], 784-4096-4096-10): see bench_extra.py.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = [784, 128, 64, 10]
ACTS = ["relu", "relu", "linear"]
TOTAL_SAMPLES = 60000
MEMORY = 10
PARITY_ITERS = 5       # N > 1: first iterations replayed on one GPU
PARITY_TOL = 2e-5


def flop_per_sample(dims):  # 2*(2*sum(in*out) + sum over layers >= 2 of in*out)
    pairs = list(zip(dims[:-1], dims[1:]))
    return 2 * (2 * sum(a * b for a, b in pairs) + sum(a * b for a, b in pairs[1:]))


def workload_name(dims, samples=TOTAL_SAMPLES, memory=MEMORY):
    return f"lbfgs_m{memory}_mlp{'-'.join(map(str, dims))}_B{samples}_fullbatch"


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
                                          str(self.index), "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
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


# ---- the CPU arm: the reference's own CPU L-BFGS on the host cores, full-size workload -----------------------------------------
def cpu_lbfgs_rate(steps, warmup, dims, acts, params, X, T, memory=MEMORY):
    """(iterations/s, info). Reference CPU sources when their build is present (kind "reference"), else the oracle port."""
    cores = os.cpu_count() or 1
    n = X.shape[0]
    iters = warmup + steps
    try:
        from oracle import ref_cpu_binding as rc
        use_ref = rc.available() and rc.net_id(dims) is not None
    except Exception:
        use_ref = False
    if use_ref:
        rc.set_num_threads(cores)  # torchrun exports OMP_NUM_THREADS=1 to its workers: pin the arm to every host core
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        sys.stdout.flush()
        os.dup2(devnull, 1)  # the reference prints its banners with std::cout
        try:
            r = rc.RefCpuNet(dims).full_batch("lbfgs", params, X, T, max_iters=iters, tolerance=0.0, m_param=memory)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(devnull); os.close(saved)
        ms, done, threads, kind = r["ms"], r["iters"], rc.num_threads(), "reference"
        what = ("the reference's own CPU sources (src/minimizer/lbfgs.hpp weak-Wolfe L-BFGS on src/network.hpp via run_full_batch_cpu), "
                "compiled unmodified against an Eigen-API stand-in (Eigen is absent from the image), fp64")
    else:
        from oracle import binding as ob
        ob.build()
        ob.set_num_threads(cores)
        r = ob.OracleNet(dims, acts).lbfgs(params, X, T, m=memory, max_iters=iters, tol=0.0, policy="cpu")
        ms, done, threads, kind = r["ms"], r["iters"], ob.num_threads(), "port"
        what = "the oracle's fp64 restatement of the reference CPU algorithm (weak-Wolfe L-BFGS)"
    assert done == iters, (done, iters)
    t = (ms[iters - 1] - (ms[warmup - 1] if warmup > 0 else 0.0)) / 1e3
    rate = steps / t
    info = dict(cores=threads, kind=kind, seconds=t,
                sample=f"{steps} L-BFGS iterations (after {warmup} warm-up) of {what} on ALL {n} samples of the workload, m = {memory}, "
                       f"{threads} host threads; no extrapolation")
    return rate, info


def run_reference(args, rank):
    """--impl reference: rank 0 alone times the reference's CPU implementation of the path; other ranks exit at once."""
    if rank != 0:
        return
    import lbfgs_ffnn_b200 as P
    from oracle import binding as ob
    ob.build()
    X, T = P.synthetic_mnist(TOTAL_SAMPLES)
    w = ob.OracleNet(DIMS, ACTS).init_params_cuda_rule(123)  # CudaNetwork::bindParams rule: the b200 arm's starting point
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    rate, info = cpu_lbfgs_rate(steps, warmup, DIMS, ACTS, w, X, T)
    line = {"impl": "reference", "metric": "lbfgs_iters_per_sec", "value": rate, "unit": "iterations/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 / rate, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(DIMS), "line_search": "weak-wolfe (reference CPU backend)", "memory": MEMORY,
                       "samples": TOTAL_SAMPLES},
            "cpu_baseline": dict(value=rate, unit="iterations/s", **{k: info[k] for k in ("cores", "kind", "sample")}),
            "e2e": {"value": rate, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def reference_cuda_rate(dims, acts, dX, dT, samples, iters=40, warmup=10):
    """the reference's CUDA backend (cuBLAS) on this GPU, same workload, from the same starting point rule"""
    try:
        from oracle import ref_cuda_binding as rcu
        if not rcu.available():
            return None
        act_id = {"linear": 0, "tanh": 1, "relu": 2, "sigmoid": 3}
        net = rcu.RefCudaNet(dims, [act_id[a] for a in acts])
        net.bind_params(123)
        net.solve("lbfgs", dX.data_ptr(), dT.data_ptr(), samples, warmup, memory=MEMORY, record=False)  # cuBLAS init, allocations
        best = None
        for _ in range(2):  # best of two runs: the comparator should not lose to a cold cuBLAS heuristic or a clock ramp
            net.bind_params(123)
            r = net.solve("lbfgs", dX.data_ptr(), dT.data_ptr(), samples, iters, memory=MEMORY, record=False)
            if best is None or r["total_ms"] / max(1, r["iters"]) < best["total_ms"] / max(1, best["iters"]):
                best = r
        r = best
        net.close()
        return dict(value=r["iters"] / (r["total_ms"] / 1e3), unit="iterations/s", iterations=int(r["iters"]),
                    sample=f"best of two runs of {iters} L-BFGS iterations (after a {warmup}-iteration warm-up solve) of the reference CUDA backend "
                           "(cuBLAS SGEMM + src/cuda/*.cuh, compiled unmodified for sm_100a: oracle/_ref/libref_cuda.so) on the same GPU, "
                           f"same {samples} samples, m = {MEMORY}, device-resident inputs")
    except Exception as e:  # comparator only: never fail the bench line over it
        return dict(value=None, unit="iterations/s", error=str(e)[:200])


def traffic_table(workload, precision):
    """DRAM bytes per launch from the committed `ncu --set full` capture of this command — reported only while the capture
    still describes the kernels: same workload, same precision mode and unchanged kernel sources (else null)."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(path):
        return {}, "no capture committed"
    t = json.load(open(path))
    if t.get("workload") != workload or t.get("precision") != precision:
        return {}, "capture is for another workload"
    out = {}
    for k, v in t.get("dram_bytes_per_launch", {}).items():
        srcs = t.get("sources", {}).get(k, [])
        ok = True
        for rel, sha in srcs:
            p = os.path.join(ROOT, rel)
            ok = ok and os.path.exists(p) and hashlib.sha1(open(p, "rb").read()).hexdigest()[:12] == sha
        if ok:
            out[k] = v
    return out, t.get("capture", "profiles/r02_traffic.json")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--precision", default="tf32x3", choices=["fp32", "tf32x3", "tf32"],
                    help="tf32x3 (default): fp32-accurate hi/lo split on the tensor cores; fp32: FFMA; tf32: single pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    ap.add_argument("--net", default=None, help="layer widths; default 784-128-64-10 (BASELINE.json configs[2]); 784-128-10 = configs[1]")
    ap.add_argument("--config", default="lbfgs", choices=["lbfgs", "slbfgs", "gd", "sgd", "c5"],
                    help="lbfgs (default) = the headline; the others are secondary lines (bench_extra.py)")
    ap.add_argument("--samples", type=int, default=None, help="override the sample count (testing)")
    ap.add_argument("--float-input", action="store_true",
                    help="inputs that are NOT 8-bit pixels (x = u/255 + noise): layer 0 cannot use the exact fp16 copy and runs the "
                         "generic 3xTF32 tcgen05 kernels on the fp32 array (secondary line: the headline workload is image data)")
    args = ap.parse_args()
    global DIMS, ACTS, TOTAL_SAMPLES
    if args.net:
        DIMS = [int(v) for v in args.net.split("-")]
        ACTS = ["relu"] * (len(DIMS) - 2) + ["linear"]
    if args.samples:
        TOTAL_SAMPLES = args.samples
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.config != "lbfgs":
        import bench_extra
        bench_extra.run(args, rank, local_rank, world)
        return
    args.warmup = max(args.warmup, 3)
    WORKLOAD = workload_name(DIMS) + ("_floatinput" if args.float_input else "")
    FLOP_PER_SAMPLE = flop_per_sample(DIMS)

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
    if args.float_input:  # same images plus sub-quantum noise: no longer exactly float(u)/255.0f
        Xh = (Xh + np.float32(1e-3) * np.random.RandomState(7).rand(*Xh.shape).astype(np.float32)).astype(np.float32)
    Xs = torch.from_numpy(Xh[rank * shard:(rank + 1) * shard]).pin_memory()
    Ts = torch.from_numpy(Th[rank * shard:(rank + 1) * shard]).pin_memory()

    h = P.CublasHandle(local_rank)
    stream = torch.cuda.Stream()
    h.set_stream(stream.cuda_stream)  # the library launches on this torch stream, so torch.cuda.Event sees it
    if world > 1:
        uid = [P.CublasHandle.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.init_comm(uid[0], rank, world)

    def make_net(handle, global_batch):
        net = P.CudaNetwork(handle)
        for i, a in enumerate(ACTS):
            net.addLayer(DIMS[i], DIMS[i + 1], a)
        net.bindParams(123)  # CudaNetwork::bindParams rule, identical on every rank
        net.set_precision(args.precision)
        net.set_global_batch(global_batch)
        return net

    net = make_net(h, TOTAL_SAMPLES)
    n = net.params_size()
    w0 = net.get_params()
    w0_pinned = torch.from_numpy(w0.copy()).pin_memory()

    with torch.cuda.stream(stream):
        dX = torch.empty_like(Xs, device="cuda")
        dT = torch.empty_like(Ts, device="cuda")
        dX.copy_(Xs, non_blocking=True)
        dT.copy_(Ts, non_blocking=True)
    stream.synchronize()

    def new_solver(handle, iters):
        s = P.CudaLBFGS(handle)
        s.setMemory(MEMORY); s.setMaxIterations(iters); s.setTolerance(0.0)
        return s

    def timed_iterations(profile=False):
        """W warm-up iterations then exactly K timed iterations of ONE continuing minimisation"""
        net.set_params(w0)
        s = new_solver(h, args.warmup + args.steps)
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
        s = new_solver(h, iters)
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

    # ---- N > 1: the sharded run against a one-GPU replay of the same first iterations (rank 0's GPU) ---------------
    parity = None
    if world > 1:
        net.set_params(w0)
        s = new_solver(h, PARITY_ITERS)
        rec = P.IterationRecorder(); rec.init(PARITY_ITERS); s.setRecorder(rec)
        s.solve(n, net.params_data(), dX, dT, shard, net)
        multi_loss = rec.copy_to_host()[0].astype(np.float64)
        barrier()
        if rank == 0:
            h1 = P.CublasHandle(local_rank)  # no communicator: a plain single-GPU context on the same device
            net1 = make_net(h1, TOTAL_SAMPLES)
            net1.set_params(w0)
            fx, ft = P.DeviceBuffer(), P.DeviceBuffer()
            fx.copy_from_host(Xh); ft.copy_from_host(Th)
            s1 = new_solver(h1, PARITY_ITERS)
            rec1 = P.IterationRecorder(); rec1.init(PARITY_ITERS); s1.setRecorder(rec1)
            s1.solve(n, net1.params_data(), fx, ft, TOTAL_SAMPLES, net1)
            single_loss = rec1.copy_to_host()[0].astype(np.float64)
            k = min(len(multi_loss), len(single_loss))
            rel = float(np.max(np.abs(multi_loss[:k] - single_loss[:k]) / np.abs(single_loss[:k]))) if k else float("nan")
            parity = {"max_rel_diff_first5": rel, "iterations_compared": int(k), "tolerance": PARITY_TOL,
                      "multi": [float(v) for v in multi_loss[:k]], "single_gpu": [float(v) for v in single_loss[:k]]}
            net1.close(); h1.close()
        barrier()

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
        # Algorithmic work per launch of the kernels with a roof (B = samples on this GPU, fp32-accurate mode):
        #   fwd0  layer-0 forward: 2*B*in*h1 flop; reads X (exact fp16 copy when the input is 8-bit pixels, else fp32) + writes A1
        #   dw0   layer-0 [dW; db]: 2*B*(in+1)*h1 flop; reads X + delta_0 (fp16 hi|lo = 4 B/element, or fp32)
        #   fwd_l / dx_l / dw_l  hidden layers: read the activations / deltas they consume once, write what they produce once
        #   tail_fwd / tail_bwd  last layer in two passes over the penultimate activations
        #   lbfgs_direction  two-loop recursion, (4k+2)*n*4 bytes
        B, K0, N0 = shard, DIMS[0], DIMS[1]
        u8 = args.precision != "fp32" and not args.float_input
        x_bytes = B * ((K0 + 1 + 63) // 64 * 64) * 2 if u8 else B * K0 * 4  # 8-bit pixels: the block-major fp16 copy [in | 1 | pad]
        work = {
            "fwd0": dict(flops=2.0 * B * K0 * N0, bytes=x_bytes + 4.0 * B * N0),
            "dw0": dict(flops=2.0 * B * (K0 + 1) * N0, bytes=x_bytes + 4.0 * B * N0),
            "tail_fwd": dict(flops=2.0 * B * DIMS[-2] * DIMS[-1], bytes=4.0 * B * (DIMS[-2] + 3 * DIMS[-1])),
            "tail_bwd": dict(flops=4.0 * B * DIMS[-2] * DIMS[-1], bytes=4.0 * B * (2 * DIMS[-2] + DIMS[-1])),
            "lbfgs_direction": dict(flops=0.0, bytes=(4.0 * MEMORY + 2) * n * 4),
        }
        for l in range(1, len(DIMS) - 2):  # hidden layers of a deeper net
            kin, nout = DIMS[l], DIMS[l + 1]
            work[f"fwd{l}"] = dict(flops=2.0 * B * kin * nout, bytes=4.0 * B * (kin + nout))
            work[f"dx{l}"] = dict(flops=2.0 * B * kin * nout, bytes=4.0 * B * (nout + 2 * kin))
            work[f"dw{l}"] = dict(flops=2.0 * B * (kin + 1) * nout, bytes=4.0 * B * (kin + nout))
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
        traffic, traffic_src = traffic_table(WORKLOAD, args.precision) if world == 1 else ({}, "single-GPU capture only")
        for k in rooflines:
            rooflines[k]["traffic"] = traffic.get(k)
        if rooflines:
            dom = max(rooflines, key=lambda k: rooflines[k]["share_of_step"])
            roofline = dict(kernel=dom, **rooflines[dom])
            roofline["traffic_source"] = traffic_src
            roofline["peak_note"] = (f"{pk['source']} peaks: HBM {pk['hbm']} GB/s, dense 16-bit tensor {tensor_peak} TFLOP/s sustained; the bound is "
                                     f"the larger of bytes/HBM and flops/tensor for the kernel's ALGORITHMIC work; precision mode {args.precision}; "
                                     "lbfgs_direction streams an L2-resident history at this size (latency-bound)")

    if rank == 0:
        value = args.steps / (ms_total / 1e3)
        xcopy = "X 95 MB (exact fp16 copy of the 8-bit pixels; 188 MB as fp32)"
        line = {"metric": "lbfgs_iters_per_sec", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "f32 (fp32-accurate split products on the tensor cores: fp16 hi+lo x exact uint8 pixels, fp32 accumulate)", "tf32": "tf32 / fp16 single-pass operands"}[args.precision],
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "memory": MEMORY, "line_search": "armijo (reference CUDA backend)",
                           "precision": args.precision, "samples_per_gpu": shard, "params": n,
                           "flop_per_evaluation": FLOP_PER_SAMPLE * TOTAL_SAMPLES,
                           "l2": "no flush: iterations run back to back inside one solver call; per-iteration working set at 60 000 samples = "
                                 f"{xcopy} + activations / deltas 31 MB per 128-wide layer + split-K partials 16 MB + history 9 MB > 126 MB L2",
                           "evals_per_iteration": res["evals"] / args.steps,
                           "loss_first_timed": res["loss_first"], "loss_last_timed": res["loss_last"],
                           "parallelism": (f"samples sharded x{world}; flat gradient + loss all-reduced once per evaluation by the library's own "
                                           "one-shot kernel over NVLink peer memory (NCCL only for set-up)") if world > 1 else "1 GPU",
                           "multi_gpu_parity": parity},
                "gpu_launches": res["launches"],
                "clocks": clocks,
                "e2e": {"value": e2e_iters / (e2e_ms / 1e3), "unit": "iterations/s", "h2d_bytes_per_step": h2d / max(1, e2e_iters),
                        "d2h_bytes_per_step": d2h / max(1, e2e_iters), "ms_total": e2e_ms, "iterations": e2e_iters,
                        "note": "H2D of X,T,params from pinned memory + b200_lbfgs_solve + D2H of params inside the timed region"},
                "roofline": roofline, "rooflines": rooflines, "kernels": kernels}
        if world == 1 and not args.no_reference_cuda:
            line["reference_cuda"] = reference_cuda_rate(DIMS, ACTS, dX, dT, shard)
        if world == 1 and not args.no_cpu_baseline:
            k_cpu = max(3, min(args.steps, 5))  # full-size iterations cost ~1 s each on the host: a bounded sample of the same workload
            rate, info = cpu_lbfgs_rate(k_cpu, 1, DIMS, ACTS, w0, Xh, Th)
            line["cpu_baseline"] = dict(value=rate, unit="iterations/s", **{k: info[k] for k in ("cores", "kind", "sample")})
        print(json.dumps(line))
    net.close()
    h.close()
    bad_parity = parity is not None and not (parity["max_rel_diff_first5"] <= PARITY_TOL)
    if world > 1:
        flag = torch.tensor([1 if bad_parity else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        bad_parity = bool(flag.item())
        dist.destroy_process_group()
    if bad_parity:
        if rank == 0:
            print(f"multi-GPU parity check failed: {parity}", file=sys.stderr)
        sys.exit(3)


if __name__ == "__main__":
    main()
