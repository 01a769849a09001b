#!/bin/bash
# bench lines of the other workloads (round-2 build)
mkdir -p gpurun_out
for w in convblock mmimdb mono mosi; do
  timeout 600 python bench.py --workload $w --steps 100 --warmup 5 > gpurun_out/s3_bench_$w.json 2> gpurun_out/s3_bench_$w.err; echo "$w rc=$?"
  python - "$w" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/s3_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], {k: d.get(k) for k in ("value", "ms_per_step", "launches_per_step")}, "e2e", d["e2e"].get("value"), "cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("parse failed", e)
PY
done
