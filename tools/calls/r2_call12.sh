#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c12_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c12_bench.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "launches_per_step"): print(k, d.get(k))
print("e2e", d["e2e"])
print("roofline", {k: d["roofline"].get(k) for k in ("label", "achieved", "frac", "us_per_launch", "share_of_step", "traffic")})
for k, v in d["kernel_rooflines"].items(): print("  ", k, {a: (round(v[a], 3) if isinstance(v.get(a), float) else v.get(a)) for a in ("achieved", "frac", "us_per_launch", "share_of_step", "error") if a in v})
print("hbm", {k: {a: round(b, 3) if isinstance(b, float) else b for a, b in v.items()} for k, v in d["hbm_rooflines"].items() if isinstance(v, dict)})
print("eager", d["gpu_eager_baseline"])
print("cpu", d["cpu_baseline"])
print("clocks", d["clocks"])
PY
