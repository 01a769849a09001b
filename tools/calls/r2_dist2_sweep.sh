#!/bin/bash
# 2 GPUs: numerics of the data-parallel step, then the bench under a few all-reduce schedules
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -x -q -s > gpurun_out/d2_test_dist.log 2>&1; echo "dist test rc=$?"
grep -E "DIST_|passed|failed" gpurun_out/d2_test_dist.log | tail -8
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
run new MML_X=1
run old MML_IMAGE_MID=0 MML_IMAGE_AR_LATE=0
run late_only MML_IMAGE_MID=0
run ctas8 MML_NCCL_MAX_CTAS=8
run ctas32 MML_NCCL_MAX_CTAS=32
