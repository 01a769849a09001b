"""Runs a few EAGER (no CUDA graph) fused train steps and brackets the last one with cudaProfilerStart/Stop, for
`ncu --profile-from-start off`.  Usage: python tools/profile_step.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import late_fusion_oracle as O  # noqa: E402  (synthetic inputs only)
from mml_b200.avmnist import AVMNIST  # noqa: E402
from mml_b200.resnet import ResNet18, ResNet34  # noqa: E402


class Term:
    loss_fn, weight = torch.nn.CrossEntropyLoss(), 1.0


B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = AVMNIST(ResNet18(1, 64), ResNet34(1, 128), 128, dropout=0.5).to(dev)
eng = model._get_engine(dev)
eng.use_graphs = False
opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
d = O.synthetic_batch(B, 1234)
batch = {"audio_original": d["audio"], "audio_missing_index": d["audio_mask"], "image_original": d["image"],
         "image_missing_index": d["image_mask"], "labels": d["labels"], "pattern_name": ["ai"] * B}
for _ in range(3):
    model.train_step(batch, opt, {"ce": Term()}, dev, None)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = model.train_step(batch, opt, {"ce": Term()}, dev, None)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled step loss", out["loss"])
