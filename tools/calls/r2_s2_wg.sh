#!/bin/bash
# weight-gradient split depth: kernel sweep, parity, step time
mkdir -p gpurun_out
python tools/kernel_bench.py wg 2>&1 | tail -42
timeout 900 python -m pytest tests/test_conv_gpu.py -x -q -k wgrad 2>&1 | tail -2
timeout 900 python -m pytest tests/test_step_gpu.py -x -q 2>&1 | tail -2
for mt in 1 2 4 8; do TAG=both_mintiles$mt MML_WGRAD_MIN_TILES=$mt python tools/step_time.py 2>&1 | tail -1; done
