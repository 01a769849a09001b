#!/bin/bash
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/d2_bench_$tag.json 2> gpurun_out/d2_bench_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/d2_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 4), d["data_parallel"]["exposed_comm_ms_per_step"], d["data_parallel"]["no_comm_ms_per_step"], d["data_parallel"]["dp_consistent"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run res32 MML_RESERVE_SMS=32
run res24c24 MML_RESERVE_SMS=24 MML_NCCL_MAX_CTAS=24
