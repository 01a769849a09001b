#!/bin/bash
MML_PDL_MODE=audio TAG=pdl_audio python tools/step_time.py 2>&1 | tail -1
MML_PDL_MODE=none TAG=pdl_none python tools/step_time.py 2>&1 | tail -1
MML_PDL_MODE=none MML_SKIP_ENCODER=image TAG=audio_only_nopdl python tools/step_time.py 2>&1 | tail -3
MML_PDL_MODE=audio MML_SKIP_ENCODER=image TAG=audio_only_pdl python tools/step_time.py 2>&1 | tail -1
bash tools/calls/r2_launchlist.sh c8
