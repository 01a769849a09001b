"""Forward-graph and backward-graph time of the AVMNIST step separately (env MML_SKIP_ENCODER selects one encoder)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import late_fusion_oracle as O
from mml_b200.avmnist import AVMNIST
from mml_b200.resnet import ResNet18, ResNet34
dev = torch.device("cuda:0"); B = 256
torch.manual_seed(0)
model = AVMNIST(ResNet18(1, 64), ResNet34(1, 128), 128, dropout=0.5).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
class T:  loss_fn, weight = torch.nn.CrossEntropyLoss(), 1.0
d = O.synthetic_batch(B, 1)
hb = {"audio_original": d["audio"], "audio_missing_index": d["audio_mask"], "image_original": d["image"], "image_missing_index": d["image_mask"],
      "labels": d["labels"], "pattern_name": ["ai"] * B}
for i in range(4): model.train_step(hb, opt, {"ce": T()}, dev, None)
plan = next(iter(model._engine.plans.values()))
for _ in range(10): plan.train_step(False)
g_fwd, g_bwd = plan.graph_train
def timeit(fn, n=100):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_f, t_b = timeit(g_fwd.replay), timeit(g_bwd.replay)
t_s = timeit(lambda: plan.train_step(False))
print(f"{os.environ.get('TAG', '')}: forward graph {t_f:.4f} ms, backward+update graph {t_b:.4f} ms, step {t_s:.4f} ms", flush=True)
