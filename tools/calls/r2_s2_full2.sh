#!/bin/bash
# whole GPU suite + headline bench after: staging kernels, one-launch stride-2 dgrad, one-wave BN grids, deeper wgrad splits, vector mono
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/s3_tests.log 2>&1; echo "tests rc=$?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/s3_tests.log | head -30
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/s3_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/s3_bench.err
