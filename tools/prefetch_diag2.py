"""Who is stretched when an H2D copy overlaps the step graph?  (diagnostic)"""
import os, sys, time
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
loss = {"ce": T()}
d = O.synthetic_batch(B, 1)
hb = {"audio_original": d["audio"].pin_memory(), "audio_missing_index": d["audio_mask"].pin_memory(), "image_original": d["image"].pin_memory(),
      "image_missing_index": d["image_mask"].pin_memory(), "labels": d["labels"].pin_memory(), "pattern_name": ["ai"] * B}
for i in range(5): model.train_step(hb, opt, loss, dev, None)
plan = next(iter(model._engine.plans.values()))
buf = torch.empty_like(hb["audio_original"], device=dev)
streams = [torch.cuda.Stream() for _ in range(6)] + [torch.cuda.Stream(priority=-1) for _ in range(2)]
def ev(): return torch.cuda.Event(enable_timing=True)
N = 30
for si, s in enumerate(streams):
    cs, gs, tot = [], [], []
    for i in range(N):
        torch.cuda.synchronize()
        c0, c1, g0, g1 = ev(), ev(), ev(), ev()
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            c0.record(); buf.copy_(hb["audio_original"], non_blocking=True); c1.record()
        g0.record(); plan.train_step(False); g1.record()
        torch.cuda.synchronize()
        tot.append((time.perf_counter() - t0) * 1e3); cs.append(c0.elapsed_time(c1)); gs.append(g0.elapsed_time(g1))
    med = lambda v: sorted(v)[len(v) // 2]
    print(f"stream {si} (prio {s.priority}): copy {med(cs):.3f} ms  graph {med(gs):.3f} ms  wall {med(tot):.3f} ms   max graph {max(gs):.3f}", flush=True)
# order variant: launch the graph first, then the copy
s = streams[0]
cs, gs = [], []
for i in range(N):
    torch.cuda.synchronize()
    c0, c1, g0, g1 = ev(), ev(), ev(), ev()
    g0.record(); plan.train_step(False); g1.record()
    with torch.cuda.stream(s):
        c0.record(); buf.copy_(hb["audio_original"], non_blocking=True); c1.record()
    torch.cuda.synchronize()
    cs.append(c0.elapsed_time(c1)); gs.append(g0.elapsed_time(g1))
print(f"graph first, then copy on stream 0: copy {med(cs):.3f} ms  graph {med(gs):.3f} ms")
