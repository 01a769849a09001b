#!/bin/bash
timeout 600 python -m pytest tests/test_fused_gpu.py -x -q 2>&1 | tail -2
for m in 2 1 2 1; do TAG=mode$m MML_BN_WAVE=$m python tools/step_time.py 2>&1 | tail -1; done
TAG=only_audio MML_SKIP_ENCODER=image python tools/step_time.py 2>&1 | tail -1
