#!/bin/bash
# ncu --set full of the tensor-core kernels bench.py lists (final build), for roofline.traffic
mkdir -p gpurun_out
python tools/ncu_targets.py 1 halo_l1 halo_l2 a_l3 a_l4 i_l3 > gpurun_out/t_plain.log 2>&1 || { tail -5 gpurun_out/t_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'conv_|wgrad_' -c 40 -f -o gpurun_out/t_full python tools/ncu_targets.py 1 halo_l1 halo_l2 a_l3 a_l4 i_l3 > gpurun_out/t_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/t_full.ncu-rep --page raw --csv > gpurun_out/t_full_raw.csv 2>&1
grep -c . gpurun_out/t_full_raw.csv; grep "^ran" gpurun_out/t_ncu.log | head -20
