#!/bin/bash
# where the bulk-copy dots kernel spends its time at n = 2e7, m = 20: without the consumers' work (B200_DIAG=256), without the row copies (512)
TAG="${1:-dd}"; SAMPLES="${2:-8192}"
OUT=gpurun_out; mkdir -p $OUT
for d in 0 256 512; do
  B200_DIAG=$d timeout 600 python bench.py --config c5 --samples $SAMPLES --steps 3 --warmup 21 --no-cpu-baseline --no-reference-cuda > $OUT/${TAG}_diag$d.json 2> $OUT/${TAG}_diag$d.err
  echo "diag=$d rc=$?"; tail -c 200 $OUT/${TAG}_diag$d.err
  python - <<P
import json
d=json.loads(open("$OUT/${TAG}_diag$d.json").read().strip().splitlines()[-1])
print("diag=$d", {k:round(x["avg_us"],1) for k,x in d["kernels"].items() if k.startswith("lbfgs")})
P
done
