#!/bin/bash
for m in image all none; do
  for r in 16 0; do
    MML_PDL_MODE=$m MML_RESERVE_SMS=$r TAG=pdl_${m}_reserve${r} python tools/step_time.py 2>&1 | tail -1
  done
done
MML_PDL_MODE=image MML_RESERVE_SMS=32 TAG=pdl_image_reserve32 python tools/step_time.py 2>&1 | tail -1
