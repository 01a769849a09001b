"""A few graph-replayed AVMNIST steps and nothing else (for `ncu --metrics gpu__time_duration.sum` launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import late_fusion_oracle as O
from mml_b200.avmnist import AVMNIST
from mml_b200.resnet import ResNet18, ResNet34
dev = torch.device("cuda:0"); B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
model = AVMNIST(ResNet18(1, 64), ResNet34(1, 128), 128, dropout=0.5).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
class T:  loss_fn, weight = torch.nn.CrossEntropyLoss(), 1.0
d = O.synthetic_batch(B, 1)
hb = {"audio_original": d["audio"], "audio_missing_index": d["audio_mask"], "image_original": d["image"], "image_missing_index": d["image_mask"],
      "labels": d["labels"], "pattern_name": ["ai"] * B}
for i in range(5): out = model.train_step(hb, opt, {"ce": T()}, dev, None)
torch.cuda.synchronize()
print("loss", out["loss"], "launches/step", next(iter(model._engine.plans.values())).launches_per_step)
