"""Device-resident step time of the AVMNIST fused step (100 graph replays); env knobs select schedule variants."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import late_fusion_oracle as O
from mml_b200.avmnist import AVMNIST
from mml_b200.resnet import ResNet18, ResNet34
dev = torch.device("cuda:0"); B = 256
if os.environ.get('MML_BN_WAVE'):
    from mml_b200 import ops as _ops
    _ops.debug_set(3, int(os.environ['MML_BN_WAVE']))
if os.environ.get('MML_WGRAD_MIN_TILES'):
    from mml_b200 import ops as _ops
    _ops.debug_set(4, int(os.environ['MML_WGRAD_MIN_TILES']))
if os.environ.get('MML_WGRAD_NARROW'):
    from mml_b200 import ops as _ops
    _ops.debug_set(5, int(os.environ['MML_WGRAD_NARROW']))
if os.environ.get('MML_IGEMM_NARROW'):
    from mml_b200 import ops as _ops
    _ops.debug_set(6, int(os.environ['MML_IGEMM_NARROW']))
if os.environ.get('MML_SPLITK'):
    from mml_b200 import ops as _ops
    _ops.debug_set(2, int(os.environ['MML_SPLITK']))
torch.manual_seed(0)
model = AVMNIST(ResNet18(1, 64), ResNet34(1, 128), 128, dropout=0.5).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
class T:  loss_fn, weight = torch.nn.CrossEntropyLoss(), 1.0
d = O.synthetic_batch(B, 1)
hb = {"audio_original": d["audio"], "audio_missing_index": d["audio_mask"], "image_original": d["image"], "image_missing_index": d["image_mask"],
      "labels": d["labels"], "pattern_name": ["ai"] * B}
for i in range(4): out = model.train_step(hb, opt, {"ce": T()}, dev, None)
plan = next(iter(model._engine.plans.values()))
for _ in range(20): plan.train_step(False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): plan.train_step(False)
e1.record(); torch.cuda.synchronize()
print(f"{os.environ.get('TAG', '')}: {e0.elapsed_time(e1) / 100:.4f} ms/step, loss {out['loss']:.4f}", flush=True)
