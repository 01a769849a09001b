#!/bin/bash
mkdir -p gpurun_out
for t in tests/test_fused_gpu.py tests/test_step_gpu.py; do
  name=$(basename "$t" .py)
  timeout 1500 python -m pytest -s "$t" -m gpu -q -x --tb=short -p no:cacheprovider > "gpurun_out/c10_${name}.log" 2>&1
  echo "== $t rc=$? =="; tail -n 4 "gpurun_out/c10_${name}.log"
done
TAG=both python tools/step_time.py 2>&1 | tail -1
python tools/kernel_bench.py stem 2>&1 | tail -2
bash tools/calls/r2_launchlist.sh c10
