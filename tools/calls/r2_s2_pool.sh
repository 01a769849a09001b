#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_conv_gpu.py -x -q > gpurun_out/s7_tests.log 2>&1; echo "fused+conv rc=$?"; grep -E "^E |^FAILED|passed|failed" gpurun_out/s7_tests.log | head
timeout 900 python -m pytest tests/test_step_gpu.py tests/test_mono_gpu.py tests/test_convblock_gpu.py -x -q > gpurun_out/s7_step.log 2>&1; echo "step rc=$?"; grep -E "^E |^FAILED|passed|failed" gpurun_out/s7_step.log | head
for i in 1 2; do TAG=both python tools/step_time.py 2>&1 | tail -1; done
TAG=only_audio MML_SKIP_ENCODER=image python tools/step_time.py 2>&1 | tail -1
