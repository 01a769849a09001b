#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of the AVMNIST step; $1 = tag
mkdir -p gpurun_out
TAG=${1:-x}
python tools/step_once.py > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/step_once.py > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/${TAG}_plain.log
