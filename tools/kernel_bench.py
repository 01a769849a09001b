"""Times individual kernels through the C ABI with CUDA events (median of N, L2 flushed between launches)."""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mml_b200 import ops  # noqa: E402

BF = torch.bfloat16
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def stem(B, H, W):
    x = torch.rand(B, H, W, device="cuda")
    m = torch.ones(B, device="cuda")
    w = torch.randn(64, 49, device="cuda") * 0.1
    P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty(B, P, Q, 64, device="cuda", dtype=BF)
    st = ops.bn_stats_buffer(64, "cuda")
    dy = torch.randn(B, P, Q, 64, device="cuda").to(BF)
    ws = torch.empty(ops.stem_wgrad_workspace(x) // 4, device="cuda")
    dw = torch.empty(64, 49, device="cuda")
    print(f"stem B={B} {H}x{W}: fprop {timeit(lambda: ops.stem_fprop(x, m, w, y, st)):.1f} us, wgrad {timeit(lambda: ops.stem_wgrad(x, m, dy, dw, ws)):.1f} us")


def conv(N, H, W, C, K, R, st, pad, tag=""):
    g = ops.make_geom(N, H, W, C, K, R, R, st, pad)
    P, Q = ops.conv_out_hw(H, W, R, R, st, pad)
    x = torch.randn(N, H, W, C, device="cuda").to(BF)
    w = (torch.randn(K, R, R, C, device="cuda") * 0.05).to(BF)
    y = torch.empty(N, P, Q, K, device="cuda", dtype=BF)
    dy = torch.randn(N, P, Q, K, device="cuda").to(BF)
    dx = torch.empty(N, H, W, C, device="cuda", dtype=BF)
    dw = torch.zeros(K, R, R, C, device="cuda")
    stt = ops.bn_stats_buffer(K, "cuda")
    fl = 2.0 * N * P * Q * K * C * R * R
    tf = timeit(lambda: ops.conv_fprop(g, x, w, y, stt))
    td = timeit(lambda: ops.conv_dgrad(g, dy, w, dx))
    ws = ops.WgradScratch("cuda")
    tw = timeit(lambda: ops.conv_wgrad(g, x, dy, dw, ws))
    print(f"conv {tag:10s} N={N} {H}x{W} C={C} K={K} R={R} s={st}: fprop {tf:6.1f} us ({fl / tf / 1e6:6.0f} TF)  dgrad {td:6.1f} us ({fl / td / 1e6:6.0f} TF)  wgrad {tw:6.1f} us ({fl / tw / 1e6:6.0f} TF)")


def wgrad_only(N, H, W, C, K, R, st, pad, tag=""):
    g = ops.make_geom(N, H, W, C, K, R, R, st, pad)
    P, Q = ops.conv_out_hw(H, W, R, R, st, pad)
    x = torch.randn(N, H, W, C, device="cuda").to(BF)
    dy = torch.randn(N, P, Q, K, device="cuda").to(BF)
    dw = torch.zeros(K, R, R, C, device="cuda")
    ws = ops.WgradScratch("cuda")
    fl = 2.0 * N * P * Q * K * C * R * R
    tw = timeit(lambda: ops.conv_wgrad(g, x, dy, dw, ws), n=9)
    print(f"wgrad {tag:9s} N={N} {H}x{W} C={C} K={K} s={st}: {tw:6.1f} us ({fl / tw / 1e6:6.0f} TF)")


def bn(rows, Cn, tag=""):
    """Fused BatchNorm kernels on a [rows, C] bf16 tensor: GB/s = algorithmic bytes (tensors read + written once) / time."""
    x, res, dy, dy2 = (torch.randn(rows, Cn, device="cuda").to(BF) for _ in range(4))
    y, gs, dx = (torch.empty_like(x) for _ in range(3))
    stats = ops.bn_stats_buffer(Cn, "cuda")
    xf = x.float()
    stats[0, :, 0] = xf.sum(0).double()
    stats[0, :, 1] = (xf * xf).sum(0).double()
    bstat = ops.bn_stats_buffer(Cn, "cuda")
    f = lambda *sh: torch.zeros(*sh, device="cuda")
    b = ops.BNBuffers(stats, torch.ones(Cn, device="cuda"), f(Cn), f(Cn), torch.ones(Cn, device="cuda"), f(Cn), torch.ones(Cn, device="cuda"))
    dg, db = f(Cn), f(Cn)
    ops.bn_train_fwd(x, b, res, None, y, rows, Cn, True)
    nb = rows * Cn * 2
    out = []
    for mode in (0, 1):
        ops.debug_set(3, mode)
        tf = timeit(lambda: ops.bn_train_fwd(x, b, res, None, y, rows, Cn, True))
        tr = timeit(lambda: (bstat.zero_(), ops.bn_bwd_reduce(dy, dy2, y, x, b.mean, b.invstd, bstat, gs, rows, Cn, True)))
        ta = timeit(lambda: ops.bn_bwd_apply(gs, x, b.mean, b.invstd, b.gamma, bstat, dg, db, dx, rows, Cn))
        out.append(f"wave={mode}: fwd {tf:5.1f} us ({3 * nb / tf / 1e3:5.0f} GB/s)  reduce {tr:5.1f} us ({5 * nb / tr / 1e3:5.0f})  apply {ta:5.1f} us ({3 * nb / ta / 1e3:5.0f})")
    print(f"bn {tag:6s} rows={rows} C={Cn}: " + " | ".join(out))


def head(B):
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *sh: torch.randn(*sh, device="cuda", generator=g) * 0.05
    ws = [r(64, 512), r(64), r(128, 512), r(128), r(128, 192), r(128), r(64, 128), r(64), r(10, 64), r(10)]
    gs = [torch.zeros_like(w) for w in ws]
    hp, hg = ops.head_params(*ws), ops.head_grads(*gs)
    pa, pi = r(B, 512).abs(), r(B, 512).abs()
    labels = torch.randint(0, 10, (B,), device="cuda")
    scratch = torch.zeros(B, ops.head_scratch_per_sample(hp), device="cuda")
    logits, loss, pred = torch.zeros(B, 10, device="cuda"), torch.zeros(1, device="cuda"), torch.zeros(B, device="cuda", dtype=torch.int32)
    da, di = torch.zeros(B, 512, device="cuda"), torch.zeros(B, 512, device="cuda")
    f = timeit(lambda: ops.head_fwd(hp, pa, pi, labels, None, 1.0, scratch, logits, loss, pred))
    b1 = timeit(lambda: ops.head_bwd(hp, hg, pa, pi, labels, None, 1.0, scratch, 1.0, da, di, phases=1))
    b2 = timeit(lambda: ops.head_bwd(hp, hg, pa, pi, labels, None, 1.0, scratch, 1.0, da, di, phases=2))
    print(f"head B={B}: fwd(+loss) {f:.1f} us, bwd data {b1:.1f} us, bwd weights {b2:.1f} us")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    B = 256
    if what in ("all", "stem"):
        stem(B, 112, 112)
        stem(B, 28, 28)
    if what in ("all", "conv"):
        conv(B, 28, 28, 64, 64, 3, 1, 1, "a.l1")
        conv(B, 28, 28, 64, 128, 3, 2, 1, "a.l2.0c1")
        conv(B, 14, 14, 128, 128, 3, 1, 1, "a.l2")
        conv(B, 14, 14, 128, 256, 3, 2, 1, "a.l3.0c1")
        conv(B, 7, 7, 256, 256, 3, 1, 1, "a.l3")
        conv(B, 7, 7, 256, 512, 3, 2, 1, "a.l4.0c1")
        conv(B, 4, 4, 512, 512, 3, 1, 1, "a.l4")
        conv(B, 7, 7, 64, 64, 3, 1, 1, "i.l1")
        conv(B, 4, 4, 128, 128, 3, 1, 1, "i.l2")
        conv(B, 2, 2, 256, 256, 3, 1, 1, "i.l3")
        conv(B, 1, 1, 512, 512, 3, 1, 1, "i.l4")
    if what in ("all", "head"):
        head(B)
    if what in ("all", "bn"):
        bn(B * 28 * 28, 64, "a.l1")
        bn(B * 14 * 14, 128, "a.l2")
        bn(B * 7 * 7, 256, "a.l3")
        bn(B * 4 * 4, 512, "a.l4")
        bn(B * 7 * 7, 64, "i.l1")
        bn(B * 2 * 2, 256, "i.l3")
    if what == "wg":
        for mt in (1, 2, 4, 8, 16):
            ops.debug_set(4, mt)
            print(f"--- min pixel tiles per wgrad split = {mt}")
            for shp, tag in (((B, 7, 7, 64, 64, 3, 1, 1), "i.l1"), ((B, 4, 4, 128, 128, 3, 1, 1), "i.l2"), ((B, 2, 2, 256, 256, 3, 1, 1), "i.l3"),
                             ((B, 1, 1, 512, 512, 3, 1, 1), "i.l4"), ((B, 2, 2, 256, 512, 3, 2, 1), "i.l4.0c1"), ((B, 7, 7, 256, 256, 3, 1, 1), "a.l3"),
                             ((B, 4, 4, 512, 512, 3, 1, 1), "a.l4")):
                wgrad_only(*shp, tag)
    if what == "s2":
        conv(B, 28, 28, 64, 128, 3, 2, 1, "a.l2.0c1")
        conv(B, 14, 14, 128, 256, 3, 2, 1, "a.l3.0c1")
        conv(B, 7, 7, 256, 512, 3, 2, 1, "a.l4.0c1")
        conv(B, 7, 7, 64, 128, 3, 2, 1, "i.l2.0c1")
        conv(B, 4, 4, 128, 256, 3, 2, 1, "i.l3.0c1")
        conv(B, 2, 2, 256, 512, 3, 2, 1, "i.l4.0c1")
    if what == "l1":
        conv(B, 28, 28, 64, 64, 3, 1, 1, "a.l1")
    if what == "l3":
        conv(B, 7, 7, 256, 256, 3, 1, 1, "a.l3")
    if what == "l2":
        conv(B, 14, 14, 128, 128, 3, 1, 1, "a.l2")
