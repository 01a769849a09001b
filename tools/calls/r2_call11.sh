#!/bin/bash
mkdir -p gpurun_out
for t in tests/test_conv_gpu.py tests/test_fused_gpu.py tests/test_step_gpu.py tests/test_mono_gpu.py; do
  name=$(basename "$t" .py)
  timeout 1500 python -m pytest -s "$t" -m gpu -q -x --tb=short -p no:cacheprovider > "gpurun_out/c11_${name}.log" 2>&1
  echo "== $t rc=$? =="; tail -n 4 "gpurun_out/c11_${name}.log"
done
TAG=both python tools/step_time.py 2>&1 | tail -1
MML_SKIP_ENCODER=audio TAG=image_only python tools/step_time.py 2>&1 | tail -1
bash tools/calls/r2_launchlist.sh c11
