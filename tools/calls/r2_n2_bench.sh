#!/bin/bash
mkdir -p gpurun_out
for tag in a b; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/d2_bench_n2_$tag.json 2> gpurun_out/d2_bench_n2_$tag.err; echo "bench rc=$?"
python - $tag <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/d2_bench_n2_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus")}); print("e2e", {k: d["e2e"].get(k) for k in ("value", "ms_per_step", "unpipelined_ms_per_step")}); print("clocks", d.get("clocks")); print("dp", {k: d["data_parallel"][k] for k in ("dp_consistent", "exposed_comm_ms_per_step")})
except Exception as e: print("parse failed", e)
PY
done
