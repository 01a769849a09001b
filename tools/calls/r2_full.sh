#!/bin/bash
# whole GPU suite + the headline bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/full_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/full_tests.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/full_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d.get("gpu_eager_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/full_bench.err
