#!/bin/bash
# Runs each GPU test module in its own process (a trapped kernel poisons the CUDA context) and keeps the logs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for t in "$@"; do
  name=$(basename "$t" .py)
  timeout 1500 python -m pytest -s "$t" -m gpu -q --tb=short -p no:cacheprovider > "gpurun_out/${name}.log" 2>&1
  echo "== $t rc=$? =="; tail -n 25 "gpurun_out/${name}.log"
done
