#!/bin/bash
# Per-kernel event timings of one evaluation under several environment settings.
# usage (through gpurun): bash tools/ab_env.sh <tag> <net> "VAR=a VAR2=b" "VAR=c" ...
TAG="$1"; NET="$2"; shift 2
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do for cfg in "$@"; do
  echo "$cfg: $(env $cfg timeout 300 python tools/layer_timing.py $NET 60000 tf32x3 2>&1 | tail -1)"
done; done 2>&1 | tee $OUT/${TAG}_abenv.log
