"""Per-step host timings of the prefetched end-to-end loop (run under torchrun for N > 1).  env: MML_PREFETCH_SYNC=event|stream"""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import torch.distributed as dist
import late_fusion_oracle as O
from mml_b200 import dist as mdist
from mml_b200.avmnist import AVMNIST
from mml_b200.data import DevicePrefetcher
from mml_b200.resnet import ResNet18, ResNet34
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    rank, local, world = mdist.init_from_env("nccl")
else:
    rank, local = 0, 0
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
B = 256
torch.manual_seed(0)
model = AVMNIST(ResNet18(1, 64), ResNet34(1, 128), 128, dropout=0.5).to(dev)
if world > 1:
    model.enable_data_parallel(mdist.DataParallel())
opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
class T:  loss_fn, weight = torch.nn.CrossEntropyLoss(), 1.0
def pinned(seed):
    d = O.synthetic_batch(B, seed)
    return {"audio_original": d["audio"].pin_memory(), "audio_missing_index": d["audio_mask"].pin_memory(), "image_original": d["image"].pin_memory(),
            "image_missing_index": d["image_mask"].pin_memory(), "labels": d["labels"].pin_memory(), "pattern_name": ["ai"] * B}
hb = [pinned(1 + rank), pinned(50 + rank), pinned(99 + rank)]
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for tag, loop in (("prefetched", lambda n: DevicePrefetcher((hb[i % 3] for i in range(n)), dev)), ("blocking", lambda n: (hb[i % 3] for i in range(n))),
                  ("prefetched2", lambda n: DevicePrefetcher((hb[i % 3] for i in range(n)), dev))):
    for b in loop(8):
        model.train_step(b, opt, {"ce": T()}, dev, None)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    ts = [time.perf_counter()]
    for b in loop(steps):
        model.train_step(b, opt, {"ce": T()}, dev, None)
        ts.append(time.perf_counter())
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    d = [(b_ - a_) * 1e3 for a_, b_ in zip(ts[:-1], ts[1:])]
    print(f"[rank {rank}] {tag} sync={os.environ.get('MML_PREFETCH_SYNC', 'event')}: total {(t_end - ts[0]) * 1e3 / steps:.3f} ms/step, median {statistics.median(d):.3f}, "
          f"p90 {sorted(d)[int(0.9 * len(d))]:.3f}, max {max(d):.3f}, >4ms: {sum(1 for x in d if x > 4)}, first5 {[round(x, 2) for x in d[:5]]}", flush=True)
    if world > 1: dist.barrier()
