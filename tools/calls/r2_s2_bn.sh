#!/bin/bash
# BatchNorm grids: one resident wave (mml_debug_set key 3) vs the old caps; audio-only step time
mkdir -p gpurun_out
python tools/kernel_bench.py bn 2>&1 | tail -8
timeout 600 python -m pytest tests/test_fused_gpu.py -x -q 2>&1 | tail -2
TAG=both_wave1 python tools/step_time.py 2>&1 | tail -1
TAG=both_wave0 MML_BN_WAVE=0 python tools/step_time.py 2>&1 | tail -1
TAG=only_audio MML_SKIP_ENCODER=image python tools/step_time.py 2>&1 | tail -4
