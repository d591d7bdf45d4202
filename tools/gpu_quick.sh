#!/bin/bash
# Quick GPU check of a kernel change: the fp16-path parity tests, then per-kernel event timings with and without a switch.
# usage (through gpurun): bash tools/gpu_quick.sh <tag> [ENVVAR=value to compare against the default]
TAG="${1:-q}"; CMP="${2:-}"
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fp16_path.py tests/test_gpu_tensorcore.py -m gpu -q -x --timeout=300 -p no:cacheprovider > $OUT/${TAG}_quick_tests.log 2>&1
echo "tests rc=$?"; tail -8 $OUT/${TAG}_quick_tests.log
for net in 784-128-64-10 784-128-10; do
  echo "default: $(timeout 300 python tools/layer_timing.py $net 60000 tf32x3 2>&1 | tail -1)"
  if [ -n "$CMP" ]; then echo "$CMP: $(env $CMP timeout 300 python tools/layer_timing.py $net 60000 tf32x3 2>&1 | tail -1)"; fi
done 2>&1 | tee $OUT/${TAG}_quick_timing.log
