"""Host-side throughput of mml_b200.datasets.AVMNIST.batches() / background_batches(); CPU only, no GPU needed.

    python tools/loader_throughput.py [threads] [samples]

Synthetic AVMNIST-shaped arrays (112x112 fp32 audio, 28x28 uint8 images), B = 256: samples/s of the inline row gathers, and the iteration
time seen by a consumer that spends 2.4 ms per batch (the B200 step) while the worker thread gathers the next batches."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mml_b200.datasets import AVMNIST  # noqa: E402


def main():
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else os.cpu_count()
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    ds = AVMNIST.from_arrays(torch.randint(0, 10, (n,), generator=g), torch.rand(n, 112, 112, generator=g),
                             torch.randint(0, 256, (n, 28, 28), dtype=torch.uint8, generator=g), "train",
                             missing_patterns={"ai": {"audio": 0.8, "image": 1.0}}, selected_patterns=["ai"],
                             cmap=np.random.default_rng(0).random((256, 4)), generator=g, pin=False)
    for _ in range(2):  # second pass: thread pools warm
        t0 = time.perf_counter()
        seen = sum(len(b["labels"]) for b in ds.batches(256))
        dt = time.perf_counter() - t0
    print(f"threads {threads}: inline batches() {seen / dt:,.0f} samples/s ({dt / (seen / 256) * 1e3:.2f} ms per 256-sample batch)")
    for _ in range(2):
        t0 = time.perf_counter()
        k = 0
        for _b in ds.background_batches(256):
            k += 1
            time.sleep(0.0024)
        dt = time.perf_counter() - t0
    print(f"threads {threads}: background_batches() under a 2.4 ms consumer: {dt / k * 1e3:.2f} ms per iteration")


if __name__ == "__main__":
    main()
