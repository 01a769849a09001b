#!/bin/bash
# round 2, GPU call 1: new parity tests at the headline config + baseline diagnostics + ncu captures of the round-1 kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for t in tests/test_conv_gpu.py tests/test_step_gpu.py; do
  name=$(basename "$t" .py)
  timeout 1500 python -m pytest -s "$t" -m gpu -q --tb=short -p no:cacheprovider > "gpurun_out/c1_${name}.log" 2>&1
  echo "== $t rc=$? =="; tail -n 12 "gpurun_out/c1_${name}.log"
done
TAG=both python tools/step_time.py 2>&1 | tail -1
MML_SKIP_ENCODER=image TAG=audio_only python tools/step_time.py 2>&1 | tail -1
MML_SKIP_ENCODER=audio TAG=image_only python tools/step_time.py 2>&1 | tail -1
python tools/kernel_bench.py all > gpurun_out/c1_kernel_bench.log 2>&1; cat gpurun_out/c1_kernel_bench.log
python tools/ncu_targets.py 1 all > gpurun_out/c1_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv_|bn_|stem_|wgrad_" -o gpurun_out/c1_prof python tools/ncu_targets.py 1 all > gpurun_out/c1_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/c1_ncu.log
ls -la gpurun_out/c1_prof.ncu-rep
ncu -i gpurun_out/c1_prof.ncu-rep --page raw --csv > gpurun_out/c1_prof_raw.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/c1_prof.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 45000000 ]; then echo "rep too large ($sz), dropping"; rm -f gpurun_out/c1_prof.ncu-rep; fi
