#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/d4_bench.json 2> gpurun_out/d4_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/d4_bench.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus")}); print("e2e", {k: d["e2e"].get(k) for k in ("value", "ms_per_step", "unpipelined_ms_per_step")}); print("clocks", d.get("clocks")); print("dp", {k: d["data_parallel"][k] for k in ("dp_consistent", "exposed_comm_ms_per_step", "no_comm_ms_per_step")})
except Exception as e: print("parse failed", e)
PY
tail -3 gpurun_out/d4_bench.err
