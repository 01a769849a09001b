#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of the AVMNIST step; $1 = tag
mkdir -p gpurun_out
TAG=${1:-x}
python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
tail -2 gpurun_out/${TAG}_bench_plain.log | cut -c1-400
