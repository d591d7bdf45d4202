#!/bin/bash
# A/B of the default bench line under an environment switch. usage (through gpurun): bash tools/bench_ab.sh <tag> "VAR=a" "VAR=b" [bench args]
TAG="$1"; A="$2"; B="$3"; shift 3
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do for cfg in "$A" "$B"; do
  env $cfg timeout 300 python bench.py --no-cpu-baseline --no-reference-cuda "$@" > $OUT/${TAG}_ab.json 2> $OUT/${TAG}_ab.err || tail -5 $OUT/${TAG}_ab.err
  python - "$cfg" <<PY
import json,sys
try:
    d=json.load(open("$OUT/${TAG}_ab.json"))
    print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), {k: round(v["avg_us"],1) for k,v in d.get("kernels",{}).items()})
except Exception as e: print(sys.argv[1], "failed", e)
PY
done; done 2>&1 | tee $OUT/${TAG}_ab.log
