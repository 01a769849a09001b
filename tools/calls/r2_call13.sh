#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -s tests/test_conv_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "28x28x64x64 or 14x14x128x128 or 8x24" > gpurun_out/c13_conv.log 2>&1; echo "conv rc=$?"; tail -3 gpurun_out/c13_conv.log
python tools/align_test.py 2>&1 | tail -8
TAG=both python tools/step_time.py 2>&1 | tail -1
