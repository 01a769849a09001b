#!/bin/bash
# ncu --set full of the stem kernels on the audio shape (B = 256, 112x112)
mkdir -p gpurun_out
python tools/ncu_targets.py 1 stem > gpurun_out/s4_stem_plain.log 2>&1 || { tail -5 gpurun_out/s4_stem_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'stem_' -c 8 -f -o gpurun_out/s4_stem python tools/ncu_targets.py 1 stem > gpurun_out/s4_stem_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/s4_stem.ncu-rep --page raw --csv > gpurun_out/s4_stem_raw.csv 2>&1
ncu -i gpurun_out/s4_stem.ncu-rep --page details > gpurun_out/s4_stem_details.txt 2>&1
ncu -i gpurun_out/s4_stem.ncu-rep --page source --csv > gpurun_out/s4_stem_source.csv 2>&1
ls -la gpurun_out | grep s4_stem
