#!/bin/bash
# A/B of the history pass of the direction at BASELINE configs[4] size (n = 2·10^7, m = 20): bulk-copy ring vs register-staged kernel.
# usage (through gpurun): bash tools/dots_bulk_ab.sh <tag> [samples]
TAG="${1:-db}"; SAMPLES="${2:-8192}"
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_dots_bulk.py tests/test_gpu_parity.py -m gpu -q -x --timeout=600 -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -15 $OUT/${TAG}_tests.log
for v in 3 2; do
  B200_DOTS_BULK=$v timeout 600 python bench.py --config c5 --samples $SAMPLES --steps 5 --warmup 21 --no-cpu-baseline --no-reference-cuda > $OUT/${TAG}_c5_bulk$v.json 2> $OUT/${TAG}_c5_bulk$v.err
  echo "bulk=$v rc=$?"; tail -c 300 $OUT/${TAG}_c5_bulk$v.err
  python - <<P
import json
d=json.loads(open("$OUT/${TAG}_c5_bulk$v.json").read().strip().splitlines()[-1])
print("bulk=$v", d["value"], {k:round(x["avg_us"],1) for k,x in d["kernels"].items() if k.startswith("lbfgs")}, d["rooflines"].get("lbfgs_direction"))
P
done
# in situ: inside a configs[4] iteration at the per-GPU share of samples (SM clock power-limited after the GEMMs)
if [ -n "$3" ]; then
  for v in 3 2; do
    B200_DOTS_BULK=$v timeout 600 python bench.py --config c5 --samples $3 --no-cpu-baseline --no-reference-cuda > $OUT/${TAG}_c5insitu_bulk$v.json 2> $OUT/${TAG}_c5insitu_bulk$v.err
    python - <<P
import json
d=json.loads(open("$OUT/${TAG}_c5insitu_bulk$v.json").read().strip().splitlines()[-1])
print("in situ bulk=$v", d["value"], {k:round(x["avg_us"],1) for k,x in d["kernels"].items() if k.startswith("lbfgs")}, round(d["rooflines"]["lbfgs_direction"]["frac"],3))
P
  done
fi
