#!/bin/bash
# session-2 baseline: whole GPU suite, headline bench line, ncu launch list of the step, ncu --set full of the top-share kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/s2_tests.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/s2_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d.get("gpu_eager_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/s2_bench.err
python tools/step_once.py > gpurun_out/s2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s2_launches.csv python tools/step_once.py > gpurun_out/s2_ncu.log 2>&1
echo "ncu list rc=$?"; tail -1 gpurun_out/s2_plain.log
python tools/ncu_targets.py 1 i_l3 a_l3 bn_l1 > gpurun_out/s2_targets_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv_wgrad_kernel|conv_igemm_kernel|bn_' -c 12 -f -o gpurun_out/s2_full python tools/ncu_targets.py 1 i_l3 a_l3 bn_l1 > gpurun_out/s2_full_ncu.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/s2_full.ncu-rep --page details > gpurun_out/s2_full_details.txt 2>&1
ncu -i gpurun_out/s2_full.ncu-rep --page raw --csv > gpurun_out/s2_full_raw.csv 2>&1
ls -la gpurun_out | head -30
