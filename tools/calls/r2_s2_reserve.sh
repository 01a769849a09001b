#!/bin/bash
for r in 16 32 40 48 16 32; do TAG=reserve$r MML_RESERVE_SMS=$r python tools/step_time.py 2>&1 | tail -1; done
