#!/bin/bash
# where the step time goes: each encoder alone, PDL modes, split-K on/off
mkdir -p gpurun_out
for v in both image audio; do
  if [ $v = both ]; then SK=""; else SK=$v; fi
  [ $v = image ] && SK=audio; [ $v = audio ] && SK=image
  TAG="only_$v" MML_SKIP_ENCODER=$SK python tools/step_time.py 2>&1 | tail -1
done
TAG="pdl_image" MML_PDL_MODE=image python tools/step_time.py 2>&1 | tail -1
TAG="pdl_all" MML_PDL_MODE=all python tools/step_time.py 2>&1 | tail -1
TAG="image_only_pdl" MML_SKIP_ENCODER=audio MML_PDL_MODE=image python tools/step_time.py 2>&1 | tail -1
TAG="nowgradstreams" MML_WGRAD_STREAMS=0 python tools/step_time.py 2>&1 | tail -1
TAG="reserve0" MML_RESERVE_SMS=0 python tools/step_time.py 2>&1 | tail -1
TAG="reserve32" MML_RESERVE_SMS=32 python tools/step_time.py 2>&1 | tail -1
