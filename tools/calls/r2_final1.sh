#!/bin/bash
# final build: whole GPU suite, headline bench, cold + warm launch lists of the step
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/f_tests.log | head -20
timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/f_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "launches_per_step")}, "e2e", d["e2e"]["value"], "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4))
    print("step_tensor", d.get("step_tensor_roofline", {}).get("frac"), "hbm", {k: round(v["frac"], 3) for k, v in d.get("hbm_rooflines", {}).items() if isinstance(v, dict)})
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/f_bench.err
python tools/step_once.py > gpurun_out/f_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/f_launches.csv python tools/step_once.py > gpurun_out/f_ncu.log 2>&1
echo "ncu cold rc=$?"; tail -1 gpurun_out/f_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv --log-file gpurun_out/f_launches_warm.csv python tools/step_once.py > gpurun_out/f_ncu_warm.log 2>&1
echo "ncu warm rc=$?"
