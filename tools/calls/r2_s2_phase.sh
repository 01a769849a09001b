#!/bin/bash
# one-launch stride-2 dgrad + staging kernels: parity, kernel timings, step time (with split-K variants)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_staging_gpu.py tests/test_conv_gpu.py -x -q > gpurun_out/s2p_tests.log 2>&1; echo "conv+staging rc=$?"; tail -3 gpurun_out/s2p_tests.log
timeout 900 python -m pytest tests/test_step_gpu.py tests/test_mono_gpu.py -x -q > gpurun_out/s2p_step.log 2>&1; echo "step rc=$?"; tail -3 gpurun_out/s2p_step.log
python tools/kernel_bench.py s2 2>&1 | tail -6
TAG=both python tools/step_time.py 2>&1 | tail -1
TAG=only_audio MML_SKIP_ENCODER=image python tools/step_time.py 2>&1 | tail -1
TAG=only_image MML_SKIP_ENCODER=audio python tools/step_time.py 2>&1 | tail -1
TAG=splitk2 MML_SPLITK=2 python tools/step_time.py 2>&1 | tail -1
TAG=splitk4 MML_SPLITK=4 python tools/step_time.py 2>&1 | tail -1
TAG=only_image_splitk4 MML_SPLITK=4 MML_SKIP_ENCODER=audio python tools/step_time.py 2>&1 | tail -1
TAG=only_image_splitk8 MML_SPLITK=8 MML_SKIP_ENCODER=audio python tools/step_time.py 2>&1 | tail -1
