#!/bin/bash
for n in 0 8 32 0 8 32; do TAG=igemm_narrow$n MML_IGEMM_NARROW=$n python tools/step_time.py 2>&1 | tail -1; done
TAG=image_only_narrow0 MML_IGEMM_NARROW=0 MML_SKIP_ENCODER=audio python tools/step_time.py 2>&1 | tail -1
TAG=image_only_narrow32 MML_IGEMM_NARROW=32 MML_SKIP_ENCODER=audio python tools/step_time.py 2>&1 | tail -1
