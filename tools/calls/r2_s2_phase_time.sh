#!/bin/bash
TAG=both python tools/phase_time.py 2>&1 | tail -1
TAG=only_audio MML_SKIP_ENCODER=image python tools/phase_time.py 2>&1 | tail -1
TAG=only_image MML_SKIP_ENCODER=audio python tools/phase_time.py 2>&1 | tail -1
