#!/bin/bash
for c in 16 32 8; do
  echo "== MML_NCCL_MAX_CTAS=$c"
  MML_NCCL_MAX_CTAS=$c timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 tools/e2e_diag.py 60 2>&1 | grep "rank 0" | cut -c1-120
done
