#!/bin/bash
# ConvBlock AVMNIST path: kernel + step parity, then the utt/gated suites that share the dense kernels (lddy change)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_convblock_gpu.py -x -q -s > gpurun_out/cb_tests.log 2>&1; echo "convblock rc=$?"
tail -30 gpurun_out/cb_tests.log
timeout 600 python -m pytest tests/test_utt_gpu.py tests/test_gated_gpu.py -x -q > gpurun_out/cb_utt.log 2>&1; echo "utt/gated rc=$?"
tail -5 gpurun_out/cb_utt.log
