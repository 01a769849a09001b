#!/bin/bash
# re-tune the schedule knobs after the wgrad change
TAG=default python tools/step_time.py 2>&1 | tail -1
TAG=pdl_image MML_PDL_MODE=image python tools/step_time.py 2>&1 | tail -1
TAG=pdl_all MML_PDL_MODE=all python tools/step_time.py 2>&1 | tail -1
TAG=reserve0 MML_RESERVE_SMS=0 python tools/step_time.py 2>&1 | tail -1
TAG=reserve8 MML_RESERVE_SMS=8 python tools/step_time.py 2>&1 | tail -1
TAG=reserve24 MML_RESERVE_SMS=24 python tools/step_time.py 2>&1 | tail -1
TAG=reserve32 MML_RESERVE_SMS=32 python tools/step_time.py 2>&1 | tail -1
TAG=prio0 MML_SIDE_PRIO=0 python tools/step_time.py 2>&1 | tail -1
TAG=wgradstreams0 MML_WGRAD_STREAMS=0 python tools/step_time.py 2>&1 | tail -1
TAG=only_audio MML_SKIP_ENCODER=image python tools/step_time.py 2>&1 | tail -1
TAG=only_image MML_SKIP_ENCODER=audio python tools/step_time.py 2>&1 | tail -1
TAG=only_image_pdl MML_SKIP_ENCODER=audio MML_PDL_MODE=image python tools/step_time.py 2>&1 | tail -1
