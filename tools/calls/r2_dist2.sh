#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 1200 python -m pytest -s tests/test_dist_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/d2_test_dist.log 2>&1; echo "dist rc=$?"; grep "DIST_\|passed\|failed\|Error" gpurun_out/d2_test_dist.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/d2_bench_n2.json 2> gpurun_out/d2_bench_n2.err; echo "bench rc=$?"; tail -2 gpurun_out/d2_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/d2_bench_n2.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus")}); print("e2e", {k: d["e2e"].get(k) for k in ("value", "ms_per_step", "unpipelined_ms_per_step")}); print("dp", d.get("data_parallel"))
except Exception as e: print("parse failed", e)
PY
