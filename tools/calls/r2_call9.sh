#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -s tests/test_fused_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/c9_fused.log 2>&1; echo "fused rc=$?"; tail -3 gpurun_out/c9_fused.log
TAG=both python tools/step_time.py 2>&1 | tail -1
python tools/ncu_targets.py 1 stem halo_l1 bn_l1 > gpurun_out/c9_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"stem_fprop|stem_wgrad_tc|conv_halo_kernel|bn_bwd" -o gpurun_out/c9_prof python tools/ncu_targets.py 1 stem halo_l1 bn_l1 > gpurun_out/c9_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/c9_prof.ncu-rep
ncu -i gpurun_out/c9_prof.ncu-rep --page source --csv > gpurun_out/c9_source.csv 2>/dev/null
ncu -i gpurun_out/c9_prof.ncu-rep --page raw --csv > gpurun_out/c9_raw.csv 2>/dev/null
ls -la gpurun_out/
