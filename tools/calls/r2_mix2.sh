#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_step_gpu.py tests/test_convblock_gpu.py tests/test_mono_gpu.py -x -q > gpurun_out/mix_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/mix_tests.log
timeout 300 python tools/host_overhead.py 200 > gpurun_out/host_overhead.txt 2>&1; echo "host rc=$?"
grep -E "ms/step" gpurun_out/host_overhead.txt
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/mix_bench.json 2> gpurun_out/mix_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/mix_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"])
PY
timeout 600 python bench.py --workload convblock --steps 100 --warmup 5 > gpurun_out/mix_bench_cb.json 2> gpurun_out/mix_bench_cb.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/mix_bench_cb.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"])
PY
