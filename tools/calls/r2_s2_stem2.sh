#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py -x -q -s -k "stem or pool" > gpurun_out/s5_stem_tests.log 2>&1; echo "stem tests rc=$?"; grep -E "stem wgrad\+BN|^E |passed|failed" gpurun_out/s5_stem_tests.log | head -20
timeout 1200 python -m pytest tests/test_step_gpu.py tests/test_mono_gpu.py -x -q > gpurun_out/s5_step_tests.log 2>&1; echo "step tests rc=$?"; grep -E "^E |^FAILED|passed|failed" gpurun_out/s5_step_tests.log | head -20
TAG=both python tools/step_time.py 2>&1 | tail -1
TAG=only_audio MML_SKIP_ENCODER=image python tools/step_time.py 2>&1 | tail -1
TAG=only_image MML_SKIP_ENCODER=audio python tools/step_time.py 2>&1 | tail -1
TAG=both_again python tools/step_time.py 2>&1 | tail -1
