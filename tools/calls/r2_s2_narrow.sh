#!/bin/bash
for n in 0 8 32 0 8 32; do TAG=narrow$n MML_WGRAD_NARROW=$n python tools/step_time.py 2>&1 | tail -1; done
