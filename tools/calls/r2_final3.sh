#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/f3_tests.log 2>&1; echo "tests rc=$?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/f3_tests.log | head -10
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/f3_bench.json 2> gpurun_out/f3_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/f3_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "roofline", d["roofline"]["label"], round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"], "alg bytes", d["roofline"]["algorithmic_bytes_per_launch"])
print({k: (round(v["frac"], 3), v.get("traffic")) for k, v in d["kernel_rooflines"].items()})
PY
tail -2 gpurun_out/f3_bench.err
