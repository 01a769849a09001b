#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_convblock_gpu.py tests/test_utt_gpu.py -x -q > gpurun_out/mix_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/mix_tests.log
timeout 300 python tools/host_overhead.py 200 > gpurun_out/host_overhead.txt 2>&1; echo "host rc=$?"
grep -E "ms/step" gpurun_out/host_overhead.txt
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/mix_bench.json 2> gpurun_out/mix_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/mix_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["ms_per_step"])
print("dominant", d["roofline"]["label"], d["roofline"]["frac"], d["roofline"]["us_per_launch"], d["roofline"]["share_of_step"])
for k, v in d["kernel_rooflines"].items():
    print(k, round(v.get("us_per_launch", 0), 1), round(v.get("frac", 0), 3), round(v.get("share_of_step", 0), 4))
print({k: (round(v["us_per_launch"] if "us_per_launch" in v else v["us_per_launch_pair"], 1), round(v["frac"], 3)) for k, v in d["hbm_rooflines"].items() if isinstance(v, dict)})
PY
