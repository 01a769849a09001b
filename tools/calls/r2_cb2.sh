#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_convblock_gpu.py -x -q -s -k "oracle or graph or eval" > gpurun_out/cb_tests.log 2>&1; echo "convblock rc=$?"
grep -E "^B=|worst|passed|failed|Error" gpurun_out/cb_tests.log | tail -40
