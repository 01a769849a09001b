#!/bin/bash
for n in 8 128 8 128; do TAG=wgrad_narrow$n MML_WGRAD_NARROW=$n python tools/step_time.py 2>&1 | tail -1; done
python - <<'PY'
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import kernel_bench as kb
from mml_b200 import ops
B = 256
for nar in (8, 128):
    ops.debug_set(5, nar)
    print("--- narrow =", nar)
    for shp, tag in (((B, 7, 7, 256, 256, 3, 1, 1), "a.l3"), ((B, 4, 4, 512, 512, 3, 1, 1), "a.l4"), ((B, 7, 7, 256, 512, 3, 2, 1), "a.l4.0c1")):
        kb.wgrad_only(*shp, tag)
PY
