#!/bin/bash
timeout 900 python -m pytest tests/test_conv_gpu.py -x -q -k wgrad 2>&1 | tail -2
python - <<'PY'
import sys, os
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import kernel_bench as kb
from mml_b200 import ops
B = 256
for nar in (0, 8):
    ops.debug_set(5, nar)
    print("--- narrow =", nar)
    for shp, tag in (((B, 2, 2, 256, 256, 3, 1, 1), "i.l3"), ((B, 1, 1, 512, 512, 3, 1, 1), "i.l4"), ((B, 2, 2, 256, 512, 3, 2, 1), "i.l4.0c1"), ((B, 4, 4, 128, 256, 3, 2, 1), "i.l3.0c1")):
        kb.wgrad_only(*shp, tag)
PY
