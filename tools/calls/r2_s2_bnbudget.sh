#!/bin/bash
for m in 1 2 1 2; do TAG=bnwave$m MML_BN_WAVE=$m python tools/step_time.py 2>&1 | tail -1; done
TAG=bnwave2_reserve48 MML_BN_WAVE=2 MML_RESERVE_SMS=48 python tools/step_time.py 2>&1 | tail -1
TAG=bnwave2_reserve64 MML_BN_WAVE=2 MML_RESERVE_SMS=64 python tools/step_time.py 2>&1 | tail -1
