#!/bin/bash
for i in 1 2; do
TAG=cur_n32_mt48 python tools/step_time.py 2>&1 | tail -1
TAG=n128_mt64 MML_WGRAD_NARROW=128 MML_WGRAD_MIN_TILES=64 python tools/step_time.py 2>&1 | tail -1
done
python - <<'PY'
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import kernel_bench as kb
from mml_b200 import ops
ops.debug_set(5, 128); ops.debug_set(4, 64)
kb.wgrad_only(256, 7, 7, 256, 256, 3, 1, 1, "a.l3 n128 mt64")
PY
