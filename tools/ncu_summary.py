"""Turns an `ncu --set full` capture of `python bench.py --steps 5 --warmup 3 ...` into the two committed artefacts:
profiles/<tag>_ncu_full_summary.json (per kernel: duration, DRAM bytes, DRAM %, tensor pipe %, warps active, registers) and
profiles/r02_traffic.json (DRAM bytes per launch per bench scope + the sha of the kernel sources, which bench.py checks before it
reports `roofline.traffic`). usage (here, no GPU): python tools/ncu_summary.py gpurun_out/r2h_full.ncu-rep r02"""
import collections, csv, hashlib, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, units, data = rows[0], rows[1], rows[2:]
want = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread"}
idx = {k: H.index(v) for k, v in want.items()}
ki = H.index("Kernel Name")
num = lambda x: float(x.replace(",", ""))
agg = collections.OrderedDict()
for r in data:
    base = re.sub(r"[<(].*", "", r[ki]).split("::")[-1]
    t = re.search(r"<([^>]*)>", r[ki])
    agg.setdefault(base + (f"<{t.group(1)}>" if t else ""), []).append({k: num(r[i]) for k, i in idx.items()})
bscale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[idx["rd"]]]
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[idx["dur"]]]
summary = {}
for k, v in agg.items():
    mean = lambda f: sum(x[f] for x in v) / len(v)
    summary[k] = dict(launches_captured=len(v), duration_us=round(mean("dur") * tscale, 2), dram_bytes=int((mean("rd") + mean("wr")) * bscale),
                      dram_read_bytes=int(mean("rd") * bscale), dram_write_bytes=int(mean("wr") * bscale), dram_pct_of_peak=round(mean("dram_pct"), 1),
                      tensor_pipe_pct=round(mean("tensor_pct"), 1), warps_active_pct=round(mean("warps_pct"), 1), registers=int(mean("regs")))
json.dump(dict(command="ncu --set full --clock-control none --import-source on -k regex:... -s 60 -c 14 python bench.py --steps 5 --warmup 3 "
                       "--no-cpu-baseline --no-reference-cuda", note="cold-cache, serialised launches: compare shares, not absolutes",
               kernels=summary), open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.json"), "w"), indent=1)
# bench scope -> (prefix of the kernel's name in the capture, sources whose hash guards the figure)
scope = {"fwd0": ("fwd16_kernel<128, 1, 1,", ["gemm_fwd16.cu"]), "fwd1": ("fwd16_kernel<64, 1, 0,", ["gemm_fwd16.cu"]),
         "dx1": ("fwd16_kernel<128, 1, 2,", ["gemm_fwd16.cu"]), "dw1": ("dw16_kernel<128, 1,", ["gemm_dw16.cu"]),
         "dw0": ("dw16_kernel<256, 0,", ["gemm_dw16.cu"]), "tail_fwd": ("tail_fwd2_kernel<2, 10", ["tail_layer.cu"]),
         "tail_bwd": ("tail_bwd_kernel<2, 10, 1", ["tail_layer.cu"]), "lbfgs_direction": ("lbfgs_direction_kernel<2, 1", ["lbfgs_kernels.cu"]),
         "finalize": ("finalize_grad_kernel", ["network.cu"]), "split16": ("prep_w16_kernel", ["gemm_fwd16.cu"])}
sha = lambda rel: hashlib.sha1(open(os.path.join(ROOT, rel), "rb").read()).hexdigest()[:12]
traffic = dict(workload="lbfgs_m10_mlp784-128-64-10_B60000_fullbatch", precision="tf32x3", capture=f"profiles/{tag}_ncu_full_summary.json (ncu --set full, B200)",
               dram_bytes_per_launch={}, sources={})
out_name = sys.argv[3] if len(sys.argv) > 3 else "r02_traffic.json"
for s, (pref, files) in scope.items():
    k = next((name for name in summary if name.startswith(pref)), None)
    if k:
        traffic["dram_bytes_per_launch"][s] = summary[k]["dram_bytes"]
        traffic["sources"][s] = [[f"lbfgs_ffnn_b200/csrc/{f}", sha(f"lbfgs_ffnn_b200/csrc/{f}")] for f in files]
json.dump(traffic, open(os.path.join(ROOT, "profiles", out_name), "w"), indent=1)
print(json.dumps(summary, indent=1))
