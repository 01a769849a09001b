#!/bin/bash
# final build: whole GPU suite, smoke, headline bench (driver-style: default flags), reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/f2_tests.log 2>&1; echo "tests rc=$?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/f2_tests.log | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/f2_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "launches_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), "clocks", d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/f2_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/f2_bench_20.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/f2_bench_20.json')); print('20-step run:', d['value'], d['e2e']['value'])"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
