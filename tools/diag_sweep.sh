#!/bin/bash
# Timing decomposition of the fp16 layer-0 kernels: parts switched off one at a time (B200_DIAG; results are wrong, times are not)
# usage (through gpurun): bash tools/diag_sweep.sh <tag> "<diag values>" [extra env]
TAG="${1:-d}"; DIAGS="${2:-0 1 2 3 4 8 7 15}"; EXTRA="${3:-}"
OUT=gpurun_out; mkdir -p $OUT
for d in $DIAGS; do
  echo "$EXTRA diag=$d: $(env $EXTRA B200_DIAG=$d timeout 300 python tools/layer_timing.py 784-128-64-10 60000 tf32x3 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read())["us"]; print({k:d[k] for k in ("fwd0","dw0","fwd1","dx1","dw1") if k in d})')"
done 2>&1 | tee $OUT/${TAG}_diag.log
