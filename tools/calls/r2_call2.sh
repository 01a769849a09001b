#!/bin/bash
# round 2, GPU call 2: BN finalize-once + deterministic wgrad: all GPU tests, step time
mkdir -p gpurun_out
for t in tests/test_conv_gpu.py tests/test_fused_gpu.py tests/test_step_gpu.py tests/test_mono_gpu.py tests/test_gated_gpu.py tests/test_utt_gpu.py; do
  name=$(basename "$t" .py)
  timeout 1500 python -m pytest -s "$t" -m gpu -q -x --tb=short -p no:cacheprovider > "gpurun_out/c2_${name}.log" 2>&1
  echo "== $t rc=$? =="; tail -n 6 "gpurun_out/c2_${name}.log"
done
TAG=both python tools/step_time.py 2>&1 | tail -1
MML_SKIP_ENCODER=audio TAG=image_only python tools/step_time.py 2>&1 | tail -1
MML_SKIP_ENCODER=image TAG=audio_only python tools/step_time.py 2>&1 | tail -2
python tools/kernel_bench.py conv > gpurun_out/c2_kernel_bench.log 2>&1; cat gpurun_out/c2_kernel_bench.log
