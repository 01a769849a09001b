#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_step_gpu.py -q > gpurun_out/s2_steptests.log 2>&1; echo "rc=$?"; grep -E "^E |^FAILED|passed|failed" gpurun_out/s2_steptests.log | head -40
for mt in 24 32 48; do TAG=both_mintiles$mt MML_WGRAD_MIN_TILES=$mt python tools/step_time.py 2>&1 | tail -1; done
