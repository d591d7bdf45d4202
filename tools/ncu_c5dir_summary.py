"""Summary of the `ncu --set full` capture of the direction's two passes at BASELINE configs[4] size (tools/gpu_session.sh, step ncuc5:
n = 20 037 642, m = 20, ring full) -> profiles/<tag>_c5_direction_ncu_full_summary.json.
usage (here, no GPU): python tools/ncu_c5dir_summary.py gpurun_out/r4a_c5dir.ncu-rep r02"""
import csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, units, data = rows[0], rows[1], rows[2:]
want = {"duration": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
        "dram_pct_of_peak": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "registers": "launch__registers_per_thread",
        "sm_clock": "sm__cycles_elapsed.avg.per_second", "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "issue_active_pct": "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "l1_shared_pct": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"}
num = lambda x: float(x.replace(",", "")) if x not in ("", "n/a") else None
ki = H.index("Kernel Name")
out = []
for r in data:
    e = {"kernel": re.sub(r"\(.*", "", r[ki]).split("::")[-1]}
    for k, m in want.items():
        if m in H:
            i = H.index(m)
            e[k] = num(r[i]); e[k + "_unit"] = units[i]
    out.append(e)
for e in out:
    sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ts = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    if e.get("dram_read") is not None and e.get("duration"):
        b = e["dram_read"] * sc[e["dram_read_unit"]] + e["dram_write"] * sc[e["dram_write_unit"]]
        e["dram_bytes"] = int(b)
        e["dram_GB_per_s"] = round(b / (e["duration"] * ts[e["duration_unit"]]) / 1e9, 1)
import hashlib
sha = hashlib.sha1(open(os.path.join(ROOT, "lbfgs_ffnn_b200", "csrc", "lbfgs_kernels.cu"), "rb").read()).hexdigest()[:12]
json.dump(dict(sources=[["lbfgs_ffnn_b200/csrc/lbfgs_kernels.cu", sha]],  # bench_extra.py reports `traffic` only while this still matches
               command="ncu --set full --clock-control none --import-source on -k regex:lbfgs_dots_bulk_kernel|lbfgs_apply_kernel -s 42 -c 4 "
                       "python bench.py --config c5 --samples 8192 --steps 2 --warmup 21 --no-cpu-baseline --no-reference-cuda",
               workload="direction of L-BFGS m = 20 on 784-4096-4096-10 (n = 20 037 642), ring full: history pass (dots) and output pass (apply)",
               note="cold-cache, serialised launches under the profiler: rates are the profiler's, the bench line's are CUDA events",
               launches=out), open(os.path.join(ROOT, "profiles", f"{tag}_c5_direction_ncu_full_summary.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
