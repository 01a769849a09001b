#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_gpu.py -x -q -k "fedavg or prefetcher or param_groups or golden or rounding" > gpurun_out/s2_wg2_step.log 2>&1; echo "rc=$?"; grep -E "^E |Error|assert|passed|failed" gpurun_out/s2_wg2_step.log | head -30
for mt in 8 16 32 64; do TAG=both_mintiles$mt MML_WGRAD_MIN_TILES=$mt python tools/step_time.py 2>&1 | tail -1; done
TAG=only_audio_mt16 MML_WGRAD_MIN_TILES=16 MML_SKIP_ENCODER=image python tools/step_time.py 2>&1 | tail -1
TAG=only_image_mt16 MML_WGRAD_MIN_TILES=16 MML_SKIP_ENCODER=audio python tools/step_time.py 2>&1 | tail -1
