#!/bin/bash
# One GPU-box session: GPU test suite, the bench lines, an ncu launch list. Everything lands in gpurun_out/<tag>_*.
# usage (through gpurun): bash tools/gpu_session.sh <tag> [steps...]   steps: tests bench small ref extra c5small c5 ncu runner
TAG="${1:-s}"; shift || true
STEPS="${*:-tests bench small ref extra c5small ncu}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.csv 2>&1
has() { [[ " $STEPS " == *" $1 "* ]]; }
if has tests; then
  timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a $OUT/${TAG}_tests.log
  tail -15 $OUT/${TAG}_tests.log
fi
if has testsk; then  # PYTEST_K="expr": a subset, verbose prints
  timeout 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -s -k "$PYTEST_K" > $OUT/${TAG}_testsk.log 2>&1; echo "testsk rc=$?" | tee -a $OUT/${TAG}_testsk.log
  grep -E "^\[|passed|failed|FAILED|Error|^E " $OUT/${TAG}_testsk.log | tail -40
fi
if has testsall; then
  timeout 1200 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider -s > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a $OUT/${TAG}_tests.log
  grep -E "^\[|passed|failed|FAILED|Error" $OUT/${TAG}_tests.log | tail -60
fi
if has bench; then
  timeout 600 python bench.py > $OUT/${TAG}_bench_deep.json 2> $OUT/${TAG}_bench_deep.err; echo "bench deep rc=$?"; tail -c 600 $OUT/${TAG}_bench_deep.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_deep.json"))
    print("DEEP value", d["value"], "e2e", d["e2e"]["value"], "evals/it", d["config"]["evals_per_iteration"], "launches", d["gpu_launches"])
    print("  kernels", {k: round(v["avg_us"],1) for k,v in d["kernels"].items()})
    print("  roof", {k: round(v["frac"],3) for k,v in d["rooflines"].items()})
    print("  ref_cuda", d.get("reference_cuda"), "\n  cpu", d.get("cpu_baseline"))
except Exception as e: print("no deep json", e)
PY
fi
if has small; then
  timeout 600 python bench.py --net 784-128-10 > $OUT/${TAG}_bench_small.json 2> $OUT/${TAG}_bench_small.err; echo "bench small rc=$?"; tail -c 600 $OUT/${TAG}_bench_small.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_small.json"))
    print("SMALL value", d["value"], "e2e", d["e2e"]["value"], "evals/it", d["config"]["evals_per_iteration"])
    print("  kernels", {k: round(v["avg_us"],1) for k,v in d["kernels"].items()})
    print("  ref_cuda", d.get("reference_cuda"), "\n  cpu", d.get("cpu_baseline"))
except Exception as e: print("no small json", e)
PY
fi
if has ref; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench ref rc=$?"; cut -c1-400 $OUT/${TAG}_bench_ref.json
fi
if has extra; then
  for c in slbfgs gd sgd; do
    timeout 600 python bench.py --config $c > $OUT/${TAG}_bench_$c.json 2> $OUT/${TAG}_bench_$c.err; echo "bench $c rc=$?"; cut -c1-900 $OUT/${TAG}_bench_$c.json; tail -c 400 $OUT/${TAG}_bench_$c.err
  done
fi
if has c5small; then
  timeout 900 python bench.py --config c5 --samples 125000 > $OUT/${TAG}_bench_c5small.json 2> $OUT/${TAG}_bench_c5small.err; echo "bench c5small rc=$?"; cut -c1-3000 $OUT/${TAG}_bench_c5small.json; tail -c 600 $OUT/${TAG}_bench_c5small.err
fi
if has c5; then
  timeout 1500 python bench.py --config c5 > $OUT/${TAG}_bench_c5.json 2> $OUT/${TAG}_bench_c5.err; echo "bench c5 rc=$?"; cut -c1-3000 $OUT/${TAG}_bench_c5.json; tail -c 600 $OUT/${TAG}_bench_c5.err
fi
if has runner; then
  ( cd $OUT && timeout 600 ../examples/_ref/fashion_mnist_main_gpu_deep > ${TAG}_ref_runner_deep.log 2>&1; echo "runner rc=$?" >> ${TAG}_ref_runner_deep.log; tail -30 ${TAG}_ref_runner_deep.log )
fi
if has ncu; then
  CMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-cuda"
  timeout 300 $CMD > $OUT/${TAG}_ncu_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu.log 2>&1
  echo "ncu rc=$?"; tail -3 $OUT/${TAG}_ncu.log
fi
if has ncufull; then  # the top kernels once, full metric set (DRAM bytes, pipe utilisation, stall reasons)
  CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-cuda"
  timeout 300 $CMD > $OUT/${TAG}_ncufull_plain.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"fwd16_kernel|dw16_kernel|lbfgs_direction_kernel|finalize_grad_kernel|tail_|prep_w16" -s 66 -c 16 -o $OUT/${TAG}_full $CMD > $OUT/${TAG}_ncufull.log 2>&1
  echo "ncufull rc=$?"; tail -3 $OUT/${TAG}_ncufull.log
fi
if has racecheck; then
  timeout 300 python tools/sanitize_case.py > $OUT/${TAG}_sanitize_plain.log 2>&1 && \
  timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_case.py > $OUT/${TAG}_racecheck.log 2>&1
  echo "racecheck rc=$?"; tail -12 $OUT/${TAG}_racecheck.log
fi
if has memcheck; then
  timeout 300 python tools/sanitize_case.py > $OUT/${TAG}_sanitize_plain.log 2>&1 && \
  timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_case.py > $OUT/${TAG}_memcheck.log 2>&1
  echo "memcheck rc=$?"; tail -12 $OUT/${TAG}_memcheck.log
fi
if has ncuc5; then  # the direction's two passes at configs[4] size with the ring full (launches 21/22 of each kernel), full metric set
  CMD="python bench.py --config c5 --samples 8192 --steps 2 --warmup 21 --no-cpu-baseline --no-reference-cuda"
  timeout 300 $CMD > $OUT/${TAG}_ncuc5_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lbfgs_dots_bulk_kernel|lbfgs_apply_kernel" -s 42 -c 4 -o $OUT/${TAG}_c5dir $CMD > $OUT/${TAG}_ncuc5.log 2>&1
  echo "ncuc5 rc=$?"; tail -3 $OUT/${TAG}_ncuc5.log
fi
echo "session $TAG done"
