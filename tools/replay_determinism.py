"""Eager vs CUDA-graph step losses of the same model / batch (5 steps), and two graph runs against each other."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import late_fusion_oracle as O
from test_step_gpu import build, make_batch, LOSS, DEV
B = 16
d = O.synthetic_batch(B, 9, (112, 112))
runs = []
for graphs in (False, True, True, False):
    model = build(graphs=graphs)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
    runs.append([model.train_step(make_batch(d, B), opt, LOSS, torch.device(DEV), None, dropout_mask=d["dropout_mask"])["loss"] for _ in range(5)])
a = np.array(runs)
print("losses eager :", a[0]); print("losses graph :", a[1])
print("max |eager - graph| =", np.abs(a[0] - a[1]).max(), " |graph - graph| =", np.abs(a[1] - a[2]).max(), " |eager - eager| =", np.abs(a[0] - a[3]).max())
