"""Host-side cost of one AVMNIST.train_step call in the end-to-end loop (pinned batches -> DevicePrefetcher -> train_step -> loss):
cProfile over N graph-replayed steps, top functions by cumulative and by own time.  usage: python tools/host_overhead.py [steps]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import late_fusion_oracle as O  # noqa: E402
from mml_b200.avmnist import AVMNIST  # noqa: E402
from mml_b200.data import DevicePrefetcher  # noqa: E402
from mml_b200.resnet import ResNet18, ResNet34  # noqa: E402


class Term:
    def __init__(self):
        self.loss_fn, self.weight = torch.nn.CrossEntropyLoss(), 1.0


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    dev = torch.device("cuda", 0)
    B = 256
    torch.manual_seed(0)
    model = AVMNIST(ResNet18(1, 64), ResNet34(1, 128), 128, dropout=0.5).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
    loss = {"cross_entropy": Term()}
    host = []
    for i in range(3):
        d = O.synthetic_batch(B, i, (112, 112))
        b = {"audio_original": d["audio"], "audio_missing_index": d["audio_mask"], "image_original": d["image"], "image_missing_index": d["image_mask"],
             "labels": d["labels"]}
        b = {k: v.pin_memory() for k, v in b.items()}
        b["pattern_name"] = ["ai"] * B
        host.append(b)
    for i in range(5):
        model.train_step(host[i % 3], opt, loss, dev, None)
    torch.cuda.synchronize()

    def loop(n):
        for b in DevicePrefetcher((host[i % 3] for i in range(n)), dev):
            model.train_step(b, opt, loss, dev, None)

    loop(20)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(steps)
    torch.cuda.synchronize()
    print(f"e2e loop: {(time.perf_counter() - t0) / steps * 1e3:.3f} ms/step")
    plan = next(iter(model._engine.plans.values()))
    t0 = time.perf_counter()
    for _ in range(steps):
        plan.train_step(False)
    torch.cuda.synchronize()
    print(f"graph replays only: {(time.perf_counter() - t0) / steps * 1e3:.3f} ms/step")
    pr = cProfile.Profile()
    pr.enable()
    loop(steps)
    pr.disable()
    for key in ("cumulative", "tottime"):
        print(f"---- top by {key} (per step = total / {steps})")
        st = pstats.Stats(pr)
        st.sort_stats(key).print_stats(22)


if __name__ == "__main__":
    main()
