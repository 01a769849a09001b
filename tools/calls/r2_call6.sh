#!/bin/bash
mkdir -p gpurun_out
for t in tests/test_conv_gpu.py tests/test_fused_gpu.py tests/test_step_gpu.py; do
  name=$(basename "$t" .py)
  timeout 1500 python -m pytest -s "$t" -m gpu -q -x --tb=short -p no:cacheprovider > "gpurun_out/c6_${name}.log" 2>&1
  echo "== $t rc=$? =="; tail -n 6 "gpurun_out/c6_${name}.log"
done
TAG=pdl_both python tools/step_time.py 2>&1 | tail -1
MML_PDL=0 TAG=nopdl_both python tools/step_time.py 2>&1 | tail -1
MML_SKIP_ENCODER=audio TAG=pdl_image_only python tools/step_time.py 2>&1 | tail -1
MML_PDL=0 MML_SKIP_ENCODER=audio TAG=nopdl_image_only python tools/step_time.py 2>&1 | tail -1
