"""GPU parity of the HBM-bound fused kernels (through the C ABI) against plain torch fp32 references of the same op.

mask: bit-exact (data/base_dataset.py:71).  BN / pooling / head / Adam / FedAvg: fp32 tolerances stated per test; bf16
storage where the kernel stores bf16.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


def test_mask_apply_bit_exact():
    from mml_b200 import ops

    x = torch.randn(7, 33, 5, device="cuda", generator=gen(0))
    x[0, 0, 0], x[1, 0, 0], x[2, 0, 0], x[3, 0, 0] = float("inf"), float("nan"), -0.0, 1e-42
    m = torch.tensor([0.0, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0], device="cuda")
    y, yr = ops.mask_apply(x, m, want_reverse=True)
    ref = x.cpu() * m.cpu().view(-1, 1, 1)
    refr = x.cpu() * -1 * (m.cpu().view(-1, 1, 1) - 1)
    assert torch.equal(y.cpu().view(torch.int32), ref.view(torch.int32))
    assert torch.equal(yr.cpu().view(torch.int32), refr.view(torch.int32))
    # committed golden bit patterns produced by torch CPU in the build container
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "mask_bits.npz"))
    xg = torch.from_numpy(gold["x"]).view(torch.float32).cuda().view(1, -1)
    for i, mval in enumerate((0.0, 1.0)):
        yy, rr = ops.mask_apply(xg, torch.tensor([mval], device="cuda"), want_reverse=True)
        assert np.array_equal(yy.cpu().view(torch.int32).numpy()[0], gold["out"][2 * i])
        assert np.array_equal(rr.cpu().view(torch.int32).numpy()[0], gold["out"][2 * i + 1])


@pytest.mark.parametrize("B,H,W", [(4, 112, 112), (5, 28, 28), (3, 32, 94), (2, 9, 13), (256, 28, 28), (37, 112, 112)])
def test_stem_fprop_wgrad(B, H, W):
    from mml_b200 import ops

    x = torch.rand(B, H, W, device="cuda", generator=gen(1))
    m = (torch.rand(B, device="cuda", generator=gen(2)) > 0.3).float()
    w = torch.randn(64, 1, 7, 7, device="cuda", generator=gen(3)) * 0.2
    P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty(B, P, Q, 64, device="cuda", dtype=BF)
    stats = ops.bn_stats_buffer(64, "cuda")
    ops.stem_fprop(x, m, w.view(64, 49).contiguous(), y, stats)
    # operands are rounded to bf16 by the kernel (after the exact fp32 mask multiply); accumulation is fp32
    xm = (x * m.view(-1, 1, 1)).to(BF).float()
    wb = w.to(BF).float()
    ref = F.conv2d(xm.unsqueeze(1), wb, stride=2, padding=3).permute(0, 2, 3, 1)
    assert ref.shape == y.shape
    err = (y.float() - ref).abs().max().item()
    assert err <= 2.0 ** -8 * ref.abs().max().item() + 1e-5, err
    yf = y.double().reshape(-1, 64)
    stats = stats.sum(0)
    assert torch.allclose(stats[:, 0], yf.sum(0), rtol=1e-5, atol=1e-3)
    assert torch.allclose(stats[:, 1], (yf * yf).sum(0), rtol=1e-5, atol=1e-3)
    # wgrad
    dy = torch.randn(B, P, Q, 64, device="cuda", generator=gen(4)).to(BF)
    ws = torch.empty(ops.stem_wgrad_workspace(x) // 4, device="cuda")
    dw = torch.empty(64, 49, device="cuda")
    ops.stem_wgrad(x, m, dy, dw, ws)
    wr = wb.clone().requires_grad_(True)
    F.conv2d(xm.unsqueeze(1), wr, stride=2, padding=3).backward(dy.float().permute(0, 3, 1, 2))
    refw = wr.grad.view(64, 49)
    assert (dw - refw).abs().max().item() <= 1e-3 * refw.abs().max().item() + 1e-4  # fp32 sums over up to 10^6 pixels, other order


@pytest.mark.parametrize("B,H,W", [(4, 112, 112), (5, 28, 28), (3, 32, 94), (2, 9, 13), (256, 28, 28), (64, 112, 112)])
def test_stem_wgrad_with_batchnorm_backward_folded_in(B, H, W):
    """conv1 -> bn1 (train) and their autograd (resnet.py:137-138): the stem's weight / gamma / beta gradients from g, the gradient w.r.t.
    the BatchNorm OUTPUT, without ever forming dx -- against torch autograd through conv + batch_norm on the same bf16 operands, and against
    the two-pass path (BatchNorm backward apply, then the plain weight gradient)."""
    from mml_b200 import ops

    x = torch.rand(B, H, W, device="cuda", generator=gen(11))
    m = (torch.rand(B, device="cuda", generator=gen(12)) > 0.3).float()
    m[0] = 1.0
    w = torch.randn(64, 1, 7, 7, device="cuda", generator=gen(13)) * 0.2
    gamma = torch.rand(64, device="cuda", generator=gen(14)) + 0.5
    beta = torch.randn(64, device="cuda", generator=gen(15)) * 0.2
    P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    rows = B * P * Q
    raw = torch.empty(B, P, Q, 64, device="cuda", dtype=BF)
    stats = ops.bn_stats_buffer(64, "cuda")
    ops.stem_fprop(x, m, w.view(64, 49).contiguous(), raw, stats)
    # batch statistics of the STORED (bf16) stem output, as the forward's BatchNorm kernels derive them
    rd = raw.double().reshape(-1, 64)
    mean = rd.mean(0)
    var = (rd * rd).mean(0) - mean * mean
    bn = ops.BNBuffers(stats, gamma, beta, torch.zeros(64, device="cuda"), torch.ones(64, device="cuda"), mean.float(), (1.0 / torch.sqrt(var + 1e-5)).float())
    g = (torch.randn(B, P, Q, 64, device="cuda", generator=gen(16)) * (torch.rand(B, P, Q, 64, device="cuda", generator=gen(17)) > 0.4)).to(BF)  # ReLU-masked
    gd = g.double().reshape(-1, 64)
    xhat = (rd - mean) * bn.invstd.double()
    bstat = ops.bn_stats_buffer(64, "cuda")
    bstat[0] = torch.stack([gd.sum(0), (gd * xhat).sum(0)], 1)
    # reference: autograd through conv (bf16 operands, fp32 accumulate) + batch_norm, seeded with g at the BatchNorm output
    xm = (x * m.view(-1, 1, 1)).to(BF).float()
    wr = w.to(BF).float().clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    out = F.batch_norm(F.conv2d(xm.unsqueeze(1), wr, stride=2, padding=3), None, None, gr, br, True, 0.1, 1e-5)
    out.backward(g.float().permute(0, 3, 1, 2))
    refw = wr.grad.view(64, 49)
    ws = torch.empty(ops.stem_wgrad_workspace(x) // 4, device="cuda")
    dw, dgamma, dbeta = torch.empty(64, 49, device="cuda"), torch.empty(64, device="cuda"), torch.empty(64, device="cuda")
    ops.stem_wgrad_bn(x, m, g, w.view(64, 49).contiguous(), bn, bstat, dgamma, dbeta, dw, ws)
    scale = refw.abs().max().item()
    err = (dw - refw).abs().max().item()
    # two-pass path on the same inputs: dx rounded to bf16, then the plain weight gradient
    dx = torch.empty_like(raw)
    dg2, db2 = torch.empty(64, device="cuda"), torch.empty(64, device="cuda")
    ops.bn_bwd_apply(g, raw, bn.mean, bn.invstd, gamma, bstat, dg2, db2, dx, rows, 64)
    dw2 = torch.empty(64, 49, device="cuda")
    ops.stem_wgrad(x, m, dx, dw2, ws)
    err2 = (dw2 - refw).abs().max().item()
    print(f"stem wgrad+BN B={B} {H}x{W}: folded max err {err / scale:.2e}, two-pass {err2 / scale:.2e} (of max |dW| = {scale:.3g})")
    # the reference keeps the conv output in fp32 while the kernels' BatchNorm sees the bf16-rounded one: 2^-9 relative noise on xhat
    assert err <= 1e-2 * scale + 1e-4, (err, scale)
    assert err <= 2.0 * err2 + 2e-3 * scale   # at least as close to autograd as the two-pass path it replaces
    assert torch.equal(dgamma, dg2) and torch.equal(dbeta, db2)
    assert (dgamma - gr.grad).abs().max().item() <= 2e-2 * gr.grad.abs().max().item() + 1e-2
    assert (dbeta - br.grad).abs().max().item() <= 1e-3 * br.grad.abs().max().item() + 1e-2


@pytest.mark.parametrize("rows,C", [(256 * 49, 64), (1000, 128), (98, 256), (4096, 512), (7, 512)])
def test_bn_forward_backward(rows, C):
    from mml_b200 import ops

    x = (torch.randn(rows, C, device="cuda", generator=gen(5)) * 2 + 0.5).to(BF)
    res = torch.randn(rows, C, device="cuda", generator=gen(6)).to(BF)
    gamma = torch.rand(C, device="cuda", generator=gen(7)) + 0.5
    beta = torch.randn(C, device="cuda", generator=gen(8)) * 0.1
    xf = x.float()

    def mk(src):  # what a conv epilogue accumulates: fp64 (sum, sum of squares) of the stored values
        sd = src.double()
        st = ops.bn_stats_buffer(C, "cuda")
        st[0] = torch.stack([sd.sum(0), (sd * sd).sum(0)], 1) * 0.25   # spread over slots like concurrent CTAs would
        st[-1] = torch.stack([sd.sum(0), (sd * sd).sum(0)], 1) * 0.75
        return ops.BNBuffers(st, gamma, beta, torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.empty(C, device="cuda"), torch.empty(C, device="cuda"))

    bn = mk(xf)
    rm_ref, rv_ref = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    xr = xf.clone().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = beta.clone().requires_grad_(True)
    resr = res.float().clone().requires_grad_(True)
    bnref = F.batch_norm(xr.t().reshape(1, C, rows), rm_ref, rv_ref, gr, br, True, 0.1, 1e-5).reshape(C, rows).t()
    out_ref = F.relu(bnref + resr)
    y = torch.empty(rows, C, device="cuda", dtype=BF)
    ops.bn_train_fwd(x, bn, res, None, y, rows, C, True)
    assert torch.allclose(bn.mean, xf.mean(0), rtol=1e-4, atol=1e-4)
    assert torch.allclose(bn.rmean, rm_ref, rtol=1e-4, atol=1e-5) and torch.allclose(bn.rvar, rv_ref, rtol=1e-4, atol=1e-5)
    assert (y.float() - out_ref).abs().max().item() <= 2.0 ** -7 * out_ref.abs().max().item() + 1e-3
    # residual through its own BN (downsample path), and the plain variant without ReLU
    rbn = mk(res.float())
    y2 = torch.empty_like(y)
    ops.bn_train_fwd(x, mk(xf), res, rbn, y2, rows, C, True)
    rref = F.batch_norm(res.float().t().reshape(1, C, rows), None, None, gamma, beta, True, 0.1, 1e-5).reshape(C, rows).t()
    ref2 = F.relu(bnref.detach() + rref)
    assert (y2.float() - ref2).abs().max().item() <= 2.0 ** -7 * ref2.abs().max().item() + 2e-3
    y3 = torch.empty_like(y)
    ops.bn_train_fwd(x, mk(xf), None, None, y3, rows, C, False)
    assert (y3.float() - bnref.detach()).abs().max().item() <= 2.0 ** -7 * bnref.abs().max().item() + 1e-3
    # eval-mode (coefficient) form
    scale, shift = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ops.bn_eval_coeffs(C, gamma, beta, bn.rmean, bn.rvar, 1e-5, scale, shift)
    y4 = torch.empty_like(y)
    ops.bn_act_fwd(x, scale, shift, res, None, None, y4, rows, C, True)
    ref4 = F.relu(F.batch_norm(xf.t().reshape(1, C, rows), bn.rmean, bn.rvar, gamma, beta, False, 0.1, 1e-5).reshape(C, rows).t() + res.float())
    assert (y4.float() - ref4).abs().max().item() <= 2.0 ** -7 * ref4.abs().max().item() + 1e-3
    # backward (two incoming gradients, relu mask from the stored output)
    dy1 = torch.randn(rows, C, device="cuda", generator=gen(9)).to(BF)
    dy2 = torch.randn(rows, C, device="cuda", generator=gen(10)).to(BF)
    gin = (dy1.float() + dy2.float()) * (y.float() > 0).float()
    (bnref * gin).sum().backward()  # d/dx of bn with upstream gradient gin (relu mask applied explicitly)
    bstat = ops.bn_stats_buffer(C, "cuda")
    dgamma, dbeta = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    dx, gout = torch.empty(rows, C, device="cuda", dtype=BF), torch.empty(rows, C, device="cuda", dtype=BF)
    ops.bn_bwd_reduce(dy1, dy2, y, x, bn.mean, bn.invstd, bstat, gout, rows, C, True)  # pass 1 stores g, pass 2 reads it
    ops.bn_bwd_apply(gout, x, bn.mean, bn.invstd, gamma, bstat, dgamma, dbeta, dx, rows, C)
    tol = 2e-2
    assert (dgamma - gr.grad).abs().max().item() <= tol * gr.grad.abs().max().item() + 1e-2
    assert (dbeta - br.grad).abs().max().item() <= tol * br.grad.abs().max().item() + 1e-2
    assert (dx.float() - xr.grad).abs().max().item() <= tol * xr.grad.abs().max().item() + 1e-3
    assert (gout.float() - gin).abs().max().item() <= 2.0 ** -7 * gin.abs().max().item() + 1e-3
    # in-place form used for bn1 (one incoming gradient, g overwrites it, dx overwrites g)
    for _ in range(1):
        d1 = dy1.clone()
        bstat.zero_()
        ops.bn_bwd_reduce(d1, None, y, x, bn.mean, bn.invstd, bstat, d1, rows, C, True)
        ops.bn_bwd_apply(d1, x, bn.mean, bn.invstd, gamma, bstat, dgamma, dbeta, d1, rows, C)
        xr.grad = None
        bnref2 = F.batch_norm(xr.t().reshape(1, C, rows), None, None, gamma, beta, True, 0.1, 1e-5).reshape(C, rows).t()
        (bnref2 * (dy1.float() * (y.float() > 0).float())).sum().backward()
        assert (d1.float() - xr.grad).abs().max().item() <= tol * xr.grad.abs().max().item() + 1e-3


@pytest.mark.parametrize("N,H,W,C", [(3, 56, 56, 64), (2, 14, 14, 64), (2, 16, 47, 64), (1, 5, 7, 64)])
def test_maxpool(N, H, W, C):
    from mml_b200 import ops

    x = F.relu(torch.randn(N, H, W, C, device="cuda", generator=gen(11))).to(BF)  # many exact ties at 0
    P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty(N, P, Q, C, device="cuda", dtype=BF)
    am = torch.empty(N, P, Q, C, device="cuda", dtype=torch.uint8)
    ops.maxpool_fwd(x, y, am, N, H, W, C)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.max_pool2d(xr, 3, 2, 1)
    assert torch.equal(y.float(), ref.permute(0, 2, 3, 1))
    dy = torch.randn(N, P, Q, C, device="cuda", generator=gen(12)).to(BF)
    dx = torch.empty(N, H, W, C, device="cuda", dtype=BF)
    dy2 = torch.randn(N, P, Q, C, device="cuda", generator=gen(112)).to(BF)
    ops.maxpool_bwd(dy, dy2, am, dx, N, H, W, C)
    ref.backward((dy.float() + dy2.float()).permute(0, 3, 1, 2))
    refdx = xr.grad.permute(0, 2, 3, 1)
    # where the input is positive the argmax is unique almost surely; ties at zero are masked by the following ReLU
    pos = x.float() > 0
    assert (dx.float() - refdx)[pos].abs().max().item() <= 2.0 ** -6 * refdx.abs().max().item()


@pytest.mark.parametrize("N,H,W", [(3, 56, 56), (2, 14, 14), (2, 16, 47), (1, 5, 7)])
def test_stem_bn_relu_maxpool_fused(N, H, W):
    """BN(train) + ReLU + MaxPool(3,2,1) in one pass, and its backward with the ReLU mask recomputed from the raw tensor."""
    from mml_b200 import ops

    C = 64
    x = (torch.randn(N, H, W, C, device="cuda", generator=gen(60)) * 1.5 + 0.3).to(BF)
    gamma = torch.rand(C, device="cuda", generator=gen(61)) + 0.5
    beta = torch.randn(C, device="cuda", generator=gen(62)) * 0.2
    xd = x.double().reshape(-1, C)
    st = ops.bn_stats_buffer(C, "cuda")
    st[1] = torch.stack([xd.sum(0), (xd * xd).sum(0)], 1)
    bn = ops.BNBuffers(st, gamma, beta, torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.empty(C, device="cuda"), torch.empty(C, device="cuda"))
    P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty(N, P, Q, C, device="cuda", dtype=BF)
    am = torch.empty(N, P, Q, C, device="cuda", dtype=torch.uint8)
    ops.stem_bn_pool_fwd(x, bn, None, None, y, am, N, H, W, C, True)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    act = F.relu(F.batch_norm(xr, rm, rv, gr, br, True, 0.1, 1e-5))
    ref = F.max_pool2d(act, 3, 2, 1)  # the kernel takes the max of the fp32 values (like the fp32 reference) and rounds the winner
    assert (y.float() - ref.detach().permute(0, 2, 3, 1)).abs().max().item() <= 2.0 ** -7 * ref.abs().max().item() + 1e-3
    assert torch.allclose(bn.rmean, rm, rtol=1e-4, atol=1e-5) and torch.allclose(bn.rvar, rv, rtol=1e-4, atol=1e-5)
    dy1 = torch.randn(N, P, Q, C, device="cuda", generator=gen(63)).to(BF)
    dy2 = torch.randn(N, P, Q, C, device="cuda", generator=gen(64)).to(BF)
    ref.backward((dy1.float() + dy2.float()).permute(0, 3, 1, 2))
    bstat = ops.bn_stats_buffer(C, "cuda")
    dgamma, dbeta = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    dx = torch.empty(N, H, W, C, device="cuda", dtype=BF)
    ops.stem_bn_pool_bwd(dy1, dy2, am, x, bn, bstat, dgamma, dbeta, dx, N, H, W, C)
    refdx = xr.grad.permute(0, 2, 3, 1)
    tol = 3e-2
    assert (dgamma - gr.grad).abs().max().item() <= tol * gr.grad.abs().max().item() + 1e-2
    assert (dbeta - br.grad).abs().max().item() <= tol * br.grad.abs().max().item() + 1e-2
    # a handful of elements may differ where two window candidates round to the same bf16 value (tie order); compare in L2
    assert float((dx.float() - refdx).norm() / refdx.norm()) <= tol
    # eval form: coefficients instead of batch statistics
    scale, shift = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ops.bn_eval_coeffs(C, gamma, beta, bn.rmean, bn.rvar, 1e-5, scale, shift)
    y2 = torch.empty_like(y)
    ops.stem_bn_pool_fwd(x, bn, scale, shift, y2, am, N, H, W, C, False)
    ref2 = F.max_pool2d(F.relu(F.batch_norm(x.float().permute(0, 3, 1, 2), bn.rmean, bn.rvar, gamma, beta, False, 0.1, 1e-5)), 3, 2, 1)
    assert (y2.float() - ref2.permute(0, 2, 3, 1)).abs().max().item() <= 2.0 ** -7 * ref2.abs().max().item() + 1e-3


@pytest.mark.parametrize("N,HW,C", [(256, 16, 512), (5, 1, 512), (3, 3, 512)])
def test_avgpool(N, HW, C):
    from mml_b200 import ops

    x = torch.randn(N, HW, C, device="cuda", generator=gen(13)).to(BF)
    y = torch.empty(N, C, device="cuda")
    ops.avgpool_fwd(x, y, N, HW, C)
    assert torch.allclose(y, x.float().mean(1), rtol=1e-5, atol=1e-6)
    dy = torch.randn(N, C, device="cuda", generator=gen(14))
    dx = torch.empty(N, HW, C, device="cuda", dtype=BF)
    ops.avgpool_bwd(dy, dx, N, HW, C)
    ref = (dy / HW).unsqueeze(1).expand(N, HW, C)
    assert (dx.float() - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item() + 1e-7


@pytest.mark.parametrize("B,use_drop", [(256, True), (5, False), (13, True)])
def test_head_fwd_bwd(B, use_drop):
    from mml_b200 import ops

    FA = FI = 512
    EA, EI, H1, H2, NC = 64, 128, 128, 64, 10
    g = gen(15)

    def lin(o, i):
        return (torch.randn(o, i, device="cuda", generator=g) / math.sqrt(i)).requires_grad_(True), (torch.randn(o, device="cuda", generator=g) * 0.1).requires_grad_(True)

    fcA_w, fcA_b = lin(EA, FA)
    fcI_w, fcI_b = lin(EI, FI)
    w0, b0 = lin(H1, EA + EI)
    w3, b3 = lin(H2, H1)
    w5, b5 = lin(NC, H2)
    pa = torch.randn(B, FA, device="cuda", generator=g).requires_grad_(True)
    pi = torch.randn(B, FI, device="cuda", generator=g).requires_grad_(True)
    labels = torch.randint(0, NC, (B,), device="cuda", generator=g)
    drop = (torch.rand(B, H1, device="cuda", generator=g) > 0.5).to(torch.uint8) if use_drop else None
    params = [fcA_w, fcA_b, fcI_w, fcI_b, w0, b0, w3, b3, w5, b5]
    hp = ops.head_params(*[p.detach() for p in params])
    PS = ops.head_scratch_per_sample(hp)
    scratch = torch.zeros(B, PS, device="cuda")
    logits = torch.empty(B, NC, device="cuda")
    loss = torch.zeros(1, device="cuda")
    pred = torch.empty(B, device="cuda", dtype=torch.int32)
    ops.head_fwd(hp, pa.detach(), pi.detach(), labels, drop, 2.0, scratch, logits, loss, pred)
    emb = torch.cat((F.linear(pa, fcA_w, fcA_b), F.linear(pi, fcI_w, fcI_b)), 1)
    h = F.relu(F.linear(emb, w0, b0))
    if use_drop:
        h = h * drop.float() * 2.0
    h = F.relu(F.linear(h, w3, b3))
    ref_logits = F.linear(h, w5, b5)
    ref_loss = F.cross_entropy(ref_logits, labels)
    assert torch.allclose(logits, ref_logits, rtol=1e-4, atol=1e-4)
    assert abs(loss.item() - ref_loss.item()) < 1e-5 * max(1, abs(ref_loss.item())) + 1e-5
    assert torch.equal(pred.long(), torch.softmax(ref_logits, 1).argmax(1))
    grads = [torch.full_like(p, float("nan")) for p in params]
    hg = ops.head_grads(*grads)
    dpa, dpi = torch.empty_like(pa), torch.empty_like(pi)
    ops.head_bwd(hp, hg, pa.detach(), pi.detach(), labels, drop, 2.0, scratch, 1.0, dpa, dpi)
    ref_loss.backward()
    for got, p in zip(grads, params):
        assert torch.allclose(got, p.grad, rtol=1e-3, atol=1e-5), (got - p.grad).abs().max()
    assert torch.allclose(dpa, pa.grad, rtol=1e-3, atol=1e-6) and torch.allclose(dpi, pi.grad, rtol=1e-3, atol=1e-6)


def test_adam_matches_torch():
    from mml_b200 import ops

    n = 100003
    p = torch.randn(n, device="cuda", generator=gen(16))
    pt = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=5e-4, weight_decay=1e-4)
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    pb = torch.empty(n, device="cuda", dtype=BF)
    hyper = torch.tensor([5e-4, 0.9, 0.999, 1e-8, 1e-4, 1.0, 0, 0], device="cuda")
    step = torch.zeros(1, device="cuda", dtype=torch.int64)
    for it in range(5):
        g = torch.randn(n, device="cuda", generator=gen(17 + it)) * 0.01
        pt.grad = g.clone()
        opt.step()
        half = n // 2 // 4 * 4
        ops.adam_step(p[:half], g[:half], m[:half], v[:half], pb[:half], hyper, step, False)  # a step split over two ranges
        ops.adam_step(p[half:], g[half:], m[half:], v[half:], pb[half:], hyper, step, True)
    assert int(step.item()) == 5
    assert torch.allclose(p, pt.detach(), rtol=1e-5, atol=1e-7), (p - pt.detach()).abs().max()
    assert torch.allclose(m, opt.state[pt]["exp_avg"], rtol=1e-5, atol=1e-9)
    assert torch.allclose(v, opt.state[pt]["exp_avg_sq"], rtol=1e-4, atol=1e-12)
    assert torch.equal(pb, p.to(BF))


def test_cast_f32_bf16():
    from mml_b200 import ops

    for total in (1 << 20, 12345, 7):
        src32 = torch.randn(total, device="cuda", generator=gen(30))
        src = torch.empty(total, device="cuda", dtype=BF)
        ops.cast_f32_bf16(src32, src)
        assert torch.equal(src, src32.to(BF))


def test_fedavg():
    from mml_b200 import ops

    K, n = 8, 1_000_003
    clients = [torch.randn(n, device="cuda", generator=gen(40 + k)) * 0.02 for k in range(K)]
    nk = torch.arange(1, K + 1, dtype=torch.float32) * 1000
    w = (nk / nk.sum()).cuda()
    ptrs = torch.tensor([c.data_ptr() for c in clients], device="cuda", dtype=torch.int64)
    out = torch.empty(n, device="cuda")
    ops.fedavg(ptrs, w, K, out)
    ref = sum(c.double() * float(wk) for c, wk in zip(clients, w.cpu()))
    assert torch.allclose(out.double(), ref, rtol=1e-5, atol=1e-8)
    x = clients[0].clone()
    ops.scale_inplace(x, w, 3)
    assert torch.allclose(x, clients[0] * w[3])


def test_dropout_mask_rate_and_determinism():
    from mml_b200 import ops

    n = 256 * 128
    m1 = torch.empty(n, device="cuda", dtype=torch.uint8)
    m2 = torch.empty_like(m1)
    step = torch.zeros(1, device="cuda", dtype=torch.int64)
    ops.dropout_mask(m1, 0.5, 1234, step)
    ops.dropout_mask(m2, 0.5, 1234, step)
    assert torch.equal(m1, m2)
    assert abs(m1.float().mean().item() - 0.5) < 0.02
    step += 1
    ops.dropout_mask(m2, 0.5, 1234, step)
    assert not torch.equal(m1, m2)
