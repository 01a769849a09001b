"""Where does the pipelined e2e step lose time?  (diagnostic, run on the GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import late_fusion_oracle as O
from mml_b200.avmnist import AVMNIST
from mml_b200.resnet import ResNet18, ResNet34
from mml_b200.data import DevicePrefetcher

dev = torch.device("cuda:0"); B = 256
torch.manual_seed(0)
model = AVMNIST(ResNet18(1, 64), ResNet34(1, 128), 128, dropout=0.5).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
class T:  loss_fn, weight = torch.nn.CrossEntropyLoss(), 1.0
loss = {"ce": T()}
def pinned(seed):
    d = O.synthetic_batch(B, seed)
    return {"audio_original": d["audio"].pin_memory(), "audio_missing_index": d["audio_mask"].pin_memory(), "image_original": d["image"].pin_memory(),
            "image_missing_index": d["image_mask"].pin_memory(), "labels": d["labels"].pin_memory(), "pattern_name": ["ai"] * B}
hb = [pinned(s) for s in (1, 2, 3)]
for i in range(5): model.train_step(hb[i % 3], opt, loss, dev, None)
N = 100
def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); print(f"{name:50s} {(time.perf_counter()-t0)/N*1e3:.3f} ms/step", flush=True)
def copies_only():
    pf = DevicePrefetcher((hb[i % 3] for i in range(N)), dev); saved = dict(DevicePrefetcher._choice); DevicePrefetcher._choice[pf._key] = 1
    for b in pf: pass
    DevicePrefetcher._choice.clear(); DevicePrefetcher._choice.update(saved)
def direct():
    for i in range(N): model.train_step(hb[i % 3], opt, loss, dev, None)
def pipelined():
    for b in DevicePrefetcher((hb[i % 3] for i in range(N)), dev): model.train_step(b, opt, loss, dev, None)
plan = next(iter(model._engine.plans.values()))
def replay_only():
    for i in range(N): plan.train_step(False)
def replay_with_background_copy():
    s = torch.cuda.Stream(); buf = torch.empty_like(hb[0]["audio_original"], device=dev)
    for i in range(N):
        with torch.cuda.stream(s): buf.copy_(hb[i % 3]["audio_original"], non_blocking=True)
        plan.train_step(False)
        torch.cuda.synchronize()
def replay_sync_each():
    for i in range(N):
        plan.train_step(False); torch.cuda.synchronize()
timed("copies only (prefetcher, no step)", copies_only)
timed("graph replay only, back to back", replay_only)
timed("graph replay + sync each", replay_sync_each)
timed("graph replay + background H2D + sync each", replay_with_background_copy)
timed("train_step direct (blocking H2D)", direct)
timed("train_step via DevicePrefetcher", pipelined)
print("calibration (ms/step by candidate, 0 = inline):", {k: round(v * 1e3, 3) for k, v in DevicePrefetcher.calibration.items()}, "chosen", DevicePrefetcher._choice)
timed("train_step via DevicePrefetcher (calibrated)", pipelined)
timed("train_step direct (blocking H2D) again", direct)
timed("train_step via DevicePrefetcher (calibrated) again", pipelined)
if os.environ.get("MML_PROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable(); pipelined(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
