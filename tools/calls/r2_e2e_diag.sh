#!/bin/bash
for sync in event stream; do
  MML_PREFETCH_SYNC=$sync timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/e2e_diag.py 100 2>&1 | grep "rank"
done
