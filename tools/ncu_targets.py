"""Launches each named target kernel a few times on its real BASELINE configs[1] shape (batch 256), for `ncu -k regex:...`.
usage: python tools/ncu_targets.py [reps] target [target ...]     (targets: see TARGETS below; `all` = every target)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mml_b200 import ops  # noqa: E402

BF = torch.bfloat16
B = 256


def conv_case(N, H, W, C, K, R, st, pad):
    g = ops.make_geom(N, H, W, C, K, R, R, st, pad)
    P, Q = ops.conv_out_hw(H, W, R, R, st, pad)
    x = torch.randn(N, H, W, C, device="cuda").to(BF)
    w = (torch.randn(K, R, R, C, device="cuda") * 0.05).to(BF)
    y = torch.empty(N, P, Q, K, device="cuda", dtype=BF)
    dy = torch.randn(N, P, Q, K, device="cuda").to(BF)
    dx = torch.empty(N, H, W, C, device="cuda", dtype=BF)
    dw = torch.zeros(K, R, R, C, device="cuda")
    stt = ops.bn_stats_buffer(K, "cuda")
    return dict(fprop=lambda: ops.conv_fprop(g, x, w, y, stt), dgrad=lambda: ops.conv_dgrad(g, dy, w, dx), wgrad=lambda ws=ops.WgradScratch("cuda"): ops.conv_wgrad(g, x, dy, dw, ws))


def bn_case(rows, Cn):
    x = torch.randn(rows, Cn, device="cuda").to(BF)
    res = torch.randn(rows, Cn, device="cuda").to(BF)
    y = torch.empty_like(x)
    dy, dy2, dx, gs = (torch.randn(rows, Cn, device="cuda").to(BF) for _ in range(4))
    stats = ops.bn_stats_buffer(Cn, "cuda")
    xf = x.float()
    stats[0, :, 0] = xf.sum(0).double()
    stats[0, :, 1] = (xf * xf).sum(0).double()
    bstat = ops.bn_stats_buffer(Cn, "cuda")
    f = lambda *sh: torch.zeros(*sh, device="cuda")
    bn = ops.BNBuffers(stats, torch.ones(Cn, device="cuda"), f(Cn), f(Cn), torch.ones(Cn, device="cuda"), f(Cn), torch.ones(Cn, device="cuda"))
    dg, db = f(Cn), f(Cn)
    return dict(fwd=lambda: ops.bn_train_fwd(x, bn, res, None, y, rows, Cn, True),
                reduce=lambda: ops.bn_bwd_reduce(dy, dy2, y, x, bn.mean, bn.invstd, bstat, gs, rows, Cn, True),
                apply=lambda: ops.bn_bwd_apply(gs, x, bn.mean, bn.invstd, bn.gamma, bstat, dg, db, dx, rows, Cn))


def stem_case(H, W):
    x = torch.rand(B, H, W, device="cuda")
    m = torch.ones(B, device="cuda")
    w = torch.randn(64, 49, device="cuda") * 0.1
    P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty(B, P, Q, 64, device="cuda", dtype=BF)
    st = ops.bn_stats_buffer(64, "cuda")
    dy = torch.randn(B, P, Q, 64, device="cuda").to(BF)
    ws = torch.empty(ops.stem_wgrad_workspace(x) // 4, device="cuda")
    dw = torch.empty(64, 49, device="cuda")
    # fused BN + ReLU + maxpool tail and its backward on the stem output
    Cn = 64
    P1, Q1 = (P - 1) // 2 + 1, (Q - 1) // 2 + 1
    raw = torch.randn(B, P, Q, Cn, device="cuda").to(BF)
    rd = raw.double().reshape(-1, Cn)
    stb = ops.bn_stats_buffer(Cn, "cuda")
    stb[0] = torch.stack([rd.sum(0), (rd * rd).sum(0)], 1)
    f = lambda *sh: torch.zeros(*sh, device="cuda")
    bn = ops.BNBuffers(stb, torch.ones(Cn, device="cuda"), f(Cn), f(Cn), torch.ones(Cn, device="cuda"), f(Cn), torch.ones(Cn, device="cuda"))
    pool = torch.empty(B, P1, Q1, Cn, device="cuda", dtype=BF)
    am = torch.empty(B, P1, Q1, Cn, device="cuda", dtype=torch.uint8)
    dp1, dp2 = (torch.randn(B, P1, Q1, Cn, device="cuda").to(BF) for _ in range(2))
    bstat = ops.bn_stats_buffer(Cn, "cuda")
    dg, db = f(Cn), f(Cn)
    dx = torch.empty_like(raw)
    ws2 = torch.empty(ops.stem_wgrad_workspace(x) // 4, device="cuda")
    return dict(fprop=lambda: ops.stem_fprop(x, m, w, y, st), wgrad=lambda: ops.stem_wgrad(x, m, dy, dw, ws),
                pool_fwd=lambda: ops.stem_bn_pool_fwd(raw, bn, None, None, pool, am, B, P, Q, Cn, True),
                pool_bwd=lambda: ops.stem_bn_pool_bwd(dp1, dp2, am, raw, bn, bstat, dg, db, dx, B, P, Q, Cn, apply=False),
                wgrad_bn=lambda: ops.stem_wgrad_bn(x, m, dx, w, bn, bstat, dg, db, dw, ws2))


TARGETS = {
    "halo_l1": lambda: conv_case(B, 28, 28, 64, 64, 3, 1, 1),
    "halo_l2": lambda: conv_case(B, 14, 14, 128, 128, 3, 1, 1),
    "a_l3": lambda: conv_case(B, 7, 7, 256, 256, 3, 1, 1),
    "a_l4": lambda: conv_case(B, 4, 4, 512, 512, 3, 1, 1),
    "a_l3s2": lambda: conv_case(B, 14, 14, 128, 256, 3, 2, 1),
    "i_l1": lambda: conv_case(B, 7, 7, 64, 64, 3, 1, 1),
    "i_l2": lambda: conv_case(B, 4, 4, 128, 128, 3, 1, 1),
    "i_l3": lambda: conv_case(B, 2, 2, 256, 256, 3, 1, 1),
    "i_l4": lambda: conv_case(B, 1, 1, 512, 512, 3, 1, 1),
    "bn_l1": lambda: bn_case(B * 28 * 28, 64),
    "bn_l4": lambda: bn_case(B * 4 * 4, 512),
    "bn_i3": lambda: bn_case(B * 2 * 2, 256),
    "stem": lambda: stem_case(112, 112),
}

if __name__ == "__main__":
    args = sys.argv[1:]
    reps = 2
    if args and args[0].isdigit():
        reps = int(args.pop(0))
    names = list(TARGETS) if (not args or args == ["all"]) else args
    for name in names:
        fns = TARGETS[name]()
        for kind, fn in fns.items():
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            print(f"ran {name}.{kind} x{reps}", flush=True)
