#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py -x -q -s -k "stem" > gpurun_out/s4_stem_tests.log 2>&1; echo "stem tests rc=$?"; grep -E "stem wgrad\+BN|^E |passed|failed" gpurun_out/s4_stem_tests.log | head -20
timeout 1200 python -m pytest tests/test_step_gpu.py tests/test_mono_gpu.py tests/test_convblock_gpu.py -x -q > gpurun_out/s4_step_tests.log 2>&1; echo "step tests rc=$?"; grep -E "^E |^FAILED|passed|failed" gpurun_out/s4_step_tests.log | head -20
TAG=fold1 python tools/step_time.py 2>&1 | tail -1
TAG=fold0 MML_STEM_FOLD=0 python tools/step_time.py 2>&1 | tail -1
TAG=fold1_audio MML_SKIP_ENCODER=image python tools/step_time.py 2>&1 | tail -1
TAG=fold0_audio MML_STEM_FOLD=0 MML_SKIP_ENCODER=image python tools/step_time.py 2>&1 | tail -1
