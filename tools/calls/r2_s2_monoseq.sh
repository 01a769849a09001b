#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mono_gpu.py -x -q -s > gpurun_out/s6_mono_tests.log 2>&1; echo "mono tests rc=$?"; grep -E "grad rel L2|^E |^FAILED|passed|failed|Error" gpurun_out/s6_mono_tests.log | head -30
