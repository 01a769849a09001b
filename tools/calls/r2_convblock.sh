#!/bin/bash
# ConvBlock AVMNIST path: kernel + step parity, the suites sharing the dense kernels, then a bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_convblock_gpu.py -x -q -s > gpurun_out/cb_tests.log 2>&1; echo "convblock rc=$?"
grep -E "worst|passed|failed|Error" gpurun_out/cb_tests.log | tail -20
timeout 600 python -m pytest tests/test_utt_gpu.py tests/test_gated_gpu.py -x -q > gpurun_out/cb_utt.log 2>&1; echo "utt/gated rc=$?"
tail -3 gpurun_out/cb_utt.log
timeout 600 python bench.py --workload convblock --steps 50 --warmup 5 > gpurun_out/cb_bench.json 2> gpurun_out/cb_bench.err; echo "bench rc=$?"
cat gpurun_out/cb_bench.json; tail -3 gpurun_out/cb_bench.err
