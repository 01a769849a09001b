"""A/B check of the halo conv kernel (descriptor base-offset hypothesis) against torch; prints max errors per mode."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mml_b200 import ops  # noqa: E402

BF = torch.bfloat16
SHAPES = [(8, 28, 28, 64, 64), (8, 14, 14, 128, 128), (3, 8, 24, 64, 64), (5, 28, 28, 64, 64), (256, 28, 28, 64, 64), (256, 14, 14, 128, 128)]


def run(shape):
    N, H, W, C, K = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).to(BF)
    w = (torch.randn(K, 3, 3, C, device="cuda", generator=g) * (2.0 / (C * 9)) ** 0.5).to(BF)
    dy = torch.randn(N, H, W, K, device="cuda", generator=g).to(BF)
    geom = ops.make_geom(N, H, W, C, K, 3, 3, 1, 1)
    y = torch.full((N, H, W, K), float("nan"), device="cuda", dtype=BF)
    stats = ops.bn_stats_buffer(K, "cuda")
    ops.conv_fprop(geom, x, w, y, stats)
    dx = torch.full((N, H, W, C), float("nan"), device="cuda", dtype=BF)
    ops.conv_dgrad(geom, dy, w, dx)
    torch.cuda.synchronize()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = torch.nn.functional.conv2d(xr, w.float().permute(0, 3, 1, 2), padding=1)
    ref.backward(dy.float().permute(0, 3, 1, 2))
    e1 = (y.float() - ref.detach().permute(0, 2, 3, 1)).abs().max().item() / ref.abs().max().item()
    e2 = (dx.float() - xr.grad.permute(0, 2, 3, 1)).abs().max().item() / xr.grad.abs().max().item()
    yf = y.double().reshape(-1, K)
    st = stats.sum(0)
    e3 = ((st[:, 0] - yf.sum(0)).abs().max() / (yf.abs().sum(0).max() + 1e-9)).item()
    e4 = ((st[:, 1] - (yf * yf).sum(0)).abs().max() / (yf * yf).sum(0).max()).item()
    return e1, e2, e3, e4


for mode in (0, 1):
    ops.debug_set(1, 1)
    ops.debug_set(2, mode)
    for sh in SHAPES:
        try:
            e = run(sh)
            print(f"halo base_offset_mode={mode} {sh}: fprop {e[0]:.3e} dgrad {e[1]:.3e} stats {e[2]:.2e} {e[3]:.2e}  {'OK' if max(e[:2]) < 2 ** -7 and max(e[2:]) < 1e-4 else 'BAD'}", flush=True)
        except Exception as ex:  # noqa: BLE001
            print(f"halo mode={mode} {sh}: EXCEPTION {ex}", flush=True)
            sys.exit(1)
