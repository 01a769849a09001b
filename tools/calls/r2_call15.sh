#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -s tests/test_conv_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/c15_conv.log 2>&1; echo "conv rc=$?"; tail -3 gpurun_out/c15_conv.log
python tools/kernel_bench.py l1 2>&1 | tail -1
python tools/kernel_bench.py l2 2>&1 | tail -1
TAG=both python tools/step_time.py 2>&1 | tail -1
