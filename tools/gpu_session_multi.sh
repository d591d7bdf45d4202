#!/bin/bash
# N-GPU session (gpurun --gpus N): 2-rank parity test, the default bench at N ranks, optionally c5. Outputs gpurun_out/<tag>_*.
TAG="${1:-m}"; N="${2:-2}"; shift 2 || true
STEPS="${*:-test bench}"
OUT=gpurun_out; mkdir -p $OUT
has() { [[ " $STEPS " == *" $1 "* ]]; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if has test; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout=500 -p no:cacheprovider > $OUT/${TAG}_multi_test.log 2>&1; echo "multi test rc=$?"; tail -5 $OUT/${TAG}_multi_test.log
fi
if has bench; then
  timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 10 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"; tail -c 500 $OUT/${TAG}_bench_n$N.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/${TAG}_bench_n$N.json") if l.startswith("{")][-1])
    print("N=$N value", d["value"], "e2e", d["e2e"]["value"], "evals/it", d["config"]["evals_per_iteration"], "parity", d["config"]["multi_gpu_parity"] and d["config"]["multi_gpu_parity"]["max_rel_diff_first5"])
    print("  kernels", {k: round(v["avg_us"],1) for k,v in d["kernels"].items()})
except Exception as e: print("no json", e)
PY
fi
if has bench20; then
  timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench20_n$N.json 2> $OUT/${TAG}_bench20_n$N.err; echo "bench20 N=$N rc=$?"; cut -c1-300 $OUT/${TAG}_bench20_n$N.json
fi
if has c5; then
  timeout 1200 $TR bench.py --gpus $N --config c5 --samples ${C5_SAMPLES:-250000} > $OUT/${TAG}_c5_n$N.json 2> $OUT/${TAG}_c5_n$N.err; echo "c5 N=$N rc=$?"; cut -c1-2500 $OUT/${TAG}_c5_n$N.json; tail -c 500 $OUT/${TAG}_c5_n$N.err
fi
echo "multi session $TAG done"
