"""GPU parity of the fused MMIMDb gated late-fusion step (config 3) against the CPU oracle and the reference fixtures.

Precision contract (DESIGN.md "Numerics"): GEMM operands / outputs are bf16 with fp32 accumulation, everything between the
GEMMs is fp32.  Un-forced comparison with the fp32 oracle:
  logits: max abs error <= 1e-2 * logit range;  loss: 1e-2;  gradients: relative L2 global <= 0.12.
Why 0.12: rounding the GEMM outputs to bf16 flips MaxOut winners (and moves BatchNorm statistics of a 16..128-sample batch),
which re-routes gradients discretely; the ORACLE ITSELF, rounded to bf16 at the same points (emulate_bf16=True), differs
from its fp32 self by 7.0-7.3 % in the gradients on these inputs (measured).  The kernels are therefore also
held to the same-rounding-points oracle, where only accumulation order differs (1-ulp differences still flip a few
winners; measured 2.9 % at B=128): global <= 4e-2, worst tensor <= 8e-2.
The stand-alone kernels (BatchNorm1d variants, GMU, BCE head) are fp32 and are held to 1e-5 .. 1e-4.
"""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import gated_fusion_oracle as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class Term:
    def __init__(self):
        self.loss_fn, self.weight = torch.nn.BCEWithLogitsLoss(), 1.0


LOSS = {"bce": Term()}


def build(graphs=True, seed=0):
    from mml_b200.mmimdb import GatedBiModalNetwork, MLPGenreClassifier, MMIMDb, MMIMDbModalityEncoder

    torch.manual_seed(seed)
    model = MMIMDb(MMIMDbModalityEncoder(4096, 512), MMIMDbModalityEncoder(300, 512), gated_bimodal_network=GatedBiModalNetwork(512, 512, 512, 512),
                   classifier=MLPGenreClassifier(512, 23, 512)).to(DEV)
    model._get_engine(torch.device(DEV)).use_graphs = graphs
    return model


def make_batch(d, device_mask=True):
    if device_mask:
        return {"image_original": d["image"], "image_missing_index": d["image_mask"], "text_original": d["text"],
                "text_missing_index": d["text_mask"], "label": d["labels"], "pattern_name": d["pattern_name"]}
    return {"image": d["image_masked"], "text": d["text_masked"], "label": d["labels"], "pattern_name": d["pattern_name"]}


def cpu_state(model):
    return OrderedDict((k, v.detach().cpu().clone().contiguous()) for k, v in model.state_dict().items())


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


# ---------------------------------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,C", [(128, 4096), (16, 300), (37, 70), (2, 33)])
@pytest.mark.parametrize("train", [True, False])
def test_bn1d_input_mode(B, C, train):
    from mml_b200 import ops

    g = torch.Generator().manual_seed(B * 1000 + C)
    x = torch.randn(B, C, generator=g) * 2 + 0.5
    mask = (torch.rand(B, generator=g) < 0.7).float()
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    rm, rv = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5
    ld = (C + 1 + 63) // 64 * 64
    dx, dm = x.to(DEV), mask.to(DEV)
    dgam, dbet, drm, drv = gamma.to(DEV), beta.to(DEV), rm.clone().to(DEV), rv.clone().to(DEV)
    xhat, inv = torch.zeros(B, C, device=DEV), torch.zeros(C, device=DEV)
    y16, y32 = torch.zeros(B, ld, device=DEV, dtype=torch.bfloat16), torch.zeros(B, C, device=DEV)
    d = ops.bn1d_fwd_desc(ops.BN1D_INPUT, B, C, dgam, dbet, drm, drv, x=dx, mask=dm, xhat=xhat, invstd=inv, y_bf16=y16, y_f32=y32)
    ops.bn1d_fwd(d, train)
    xm = G.apply_missing_mask(x, mask)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    ref = F.batch_norm(xm, rm_ref, rv_ref, gamma, beta, train, 0.1, 1e-5)
    assert torch.allclose(y32.cpu(), ref, rtol=1e-4, atol=1e-4)
    assert torch.equal(y16[:, :C].cpu(), y32.cpu().to(torch.bfloat16))
    assert float(y16[:, C:].float().abs().max()) == 0.0  # pad / bias columns are never touched
    assert torch.allclose(drm.cpu(), rm_ref, rtol=1e-5, atol=1e-6) and torch.allclose(drv.cpu(), rv_ref, rtol=1e-4, atol=1e-6)
    if train:
        var = xm.var(0, unbiased=False)
        assert torch.allclose(inv.cpu(), 1.0 / torch.sqrt(var + 1e-5), rtol=1e-4)
        # backward (INPUT mode: dgamma / dbeta only)
        dy = torch.randn(B, ld, generator=g).to(torch.bfloat16)
        dgo, dbo = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        bd = ops.bn1d_bwd_desc(ops.BN1D_INPUT, B, C, dy.to(DEV), xhat, dgam, inv, dgo, dbo)
        ops.bn1d_bwd(bd)
        dyf = dy[:, :C].float()
        xh = (xm - xm.mean(0)) / torch.sqrt(var + 1e-5)
        assert torch.allclose(dbo.cpu(), dyf.sum(0), rtol=1e-4, atol=1e-4)
        assert torch.allclose(dgo.cpu(), (dyf * xh).sum(0), rtol=1e-3, atol=2e-3)


def _bf16(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("B,C", [(128, 512), (16, 64), (5, 96)])
@pytest.mark.parametrize("dropout", [True, False])
def test_bn1d_maxout_mode_fwd_bwd(B, C, dropout):
    from mml_b200 import ops

    g = torch.Generator().manual_seed(B + C)
    pre = _bf16(torch.randn(B, 2 * C, generator=g))
    pre[0, 1] = pre[0, C + 1]  # an exact tie: the gradient must be split
    keep = (torch.rand(B, C, generator=g) < 0.5)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    dy = _bf16(torch.randn(B, C, generator=g))
    # torch reference (autograd)
    pr = pre.clone().requires_grad_(True)
    gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    v = torch.max(pr[:, :C], pr[:, C:])
    if dropout:
        v = v * keep.float() / 0.5
    rm, rv = torch.zeros(C), torch.ones(C)
    y = F.batch_norm(v, rm, rv, gm, bt, True, 0.1, 1e-5)
    y.backward(dy)
    # kernels
    dev = lambda t, dt=None: (t if dt is None else t.to(dt)).to(DEV).contiguous()
    dpre16 = dev(pre, torch.bfloat16)
    dkeep = dev(keep, torch.uint8)
    dgam, dbet, drm, drv = dev(gamma), dev(beta), torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    xhat, inv, y32 = torch.zeros(B, C, device=DEV), torch.zeros(C, device=DEV), torch.zeros(B, C, device=DEV)
    fd = ops.bn1d_fwd_desc(ops.BN1D_MAXOUT, B, C, dgam, dbet, drm, drv, pre=dpre16, keep=dkeep, keep_scale=2.0, xhat=xhat, invstd=inv, y_f32=y32)
    ops.bn1d_fwd(fd, True, use_keep=dropout)
    assert torch.allclose(y32.cpu(), y.detach(), rtol=1e-4, atol=1e-4)
    assert torch.allclose(drm.cpu(), rm, rtol=1e-5, atol=1e-6) and torch.allclose(drv.cpu(), rv, rtol=1e-4, atol=1e-6)
    dgo, dbo = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    dpre = torch.zeros(B, 2 * C, device=DEV, dtype=torch.bfloat16)
    bd = ops.bn1d_bwd_desc(ops.BN1D_MAXOUT, B, C, dev(dy, torch.bfloat16), xhat, dgam, inv, dgo, dbo, pre=dpre16, keep=dkeep, keep_scale=2.0, dpre=dpre)
    ops.bn1d_bwd(bd, use_keep=dropout)
    assert torch.allclose(dbo.cpu(), bt.grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(dgo.cpu(), gm.grad, rtol=1e-3, atol=1e-3)
    assert rel(dpre.float().cpu(), pr.grad) < 5e-3  # bf16 output rounding
    assert abs(float(dpre[0, 1]) - float(pr.grad[0, 1])) <= 1e-2 * abs(float(pr.grad[0, 1])) + 1e-6


@pytest.mark.parametrize("B,H", [(128, 512), (7, 64)])
def test_gmu_and_gated_bn(B, H):
    from mml_b200 import ops

    g = torch.Generator().manual_seed(B)
    h1p, h2p = _bf16(torch.randn(B, H, generator=g)), _bf16(torch.randn(B, H, generator=g))
    wz = torch.randn(2 * H, generator=g) * 0.05
    gamma, beta = torch.rand(H, generator=g) + 0.5, torch.randn(H, generator=g)
    dy = _bf16(torch.randn(B, H, generator=g))
    a, b, w = h1p.clone().requires_grad_(True), h2p.clone().requires_grad_(True), wz.clone().requires_grad_(True)
    gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    t1, t2 = torch.tanh(a), torch.tanh(b)
    gate = torch.sigmoid(torch.cat([t1, t2], 1) @ w)
    z = gate.view(-1, 1) * t1 + (1 - gate).view(-1, 1) * t2
    y = F.batch_norm(z, torch.zeros(H), torch.ones(H), gm, bt, True, 0.1, 1e-5)
    y.backward(dy)
    dev = lambda t, dt=None: (t if dt is None else t.to(dt)).to(DEV).contiguous()
    h1, h2, gt = torch.zeros(B, H, device=DEV), torch.zeros(B, H, device=DEV), torch.zeros(B, device=DEV)
    dwz_in = dev(wz)
    ops.gmu_fwd(dev(h1p, torch.bfloat16), dev(h2p, torch.bfloat16), dwz_in, h1, h2, gt)
    assert torch.allclose(gt.cpu(), gate.detach(), rtol=1e-5, atol=1e-6) and torch.allclose(h1.cpu(), t1.detach(), rtol=1e-5, atol=1e-6)
    dgam, dbet = dev(gamma), dev(beta)
    xhat, inv, y32 = torch.zeros(B, H, device=DEV), torch.zeros(H, device=DEV), torch.zeros(B, H, device=DEV)
    fd = ops.bn1d_fwd_desc(ops.BN1D_GATED, B, H, dgam, dbet, torch.zeros(H, device=DEV), torch.ones(H, device=DEV), h1=h1, h2=h2, gate=gt,
                           xhat=xhat, invstd=inv, y_f32=y32)
    ops.bn1d_fwd(fd, True)
    assert torch.allclose(y32.cpu(), y.detach(), rtol=1e-4, atol=1e-4)
    dz, dgo, dbo = torch.zeros(B, H, device=DEV), torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    ops.bn1d_bwd(ops.bn1d_bwd_desc(ops.BN1D_GATED, B, H, dev(dy, torch.bfloat16), xhat, dgam, inv, dgo, dbo, dz=dz))
    assert torch.allclose(dgo.cpu(), gm.grad, rtol=1e-3, atol=1e-3) and torch.allclose(dbo.cpu(), bt.grad, rtol=1e-4, atol=1e-4)
    dwz = torch.zeros(2 * H, device=DEV)
    d1, d2 = torch.zeros(B, H, device=DEV, dtype=torch.bfloat16), torch.zeros(B, H, device=DEV, dtype=torch.bfloat16)
    ops.gmu_bwd(dz, h1, h2, gt, dwz_in, dwz, d1, d2)
    assert rel(dwz.cpu(), w.grad) < 1e-4
    assert rel(d1.float().cpu(), a.grad) < 5e-3 and rel(d2.float().cpu(), b.grad) < 5e-3


@pytest.mark.parametrize("B,H,NC", [(128, 512, 23), (9, 64, 5), (300, 512, 23)])
def test_bce_head(B, H, NC):
    from mml_b200 import ops

    g = torch.Generator().manual_seed(B + NC)
    xn = torch.randn(B, H, generator=g)
    w, bias = torch.randn(NC, H, generator=g) * 0.1, torch.randn(NC, generator=g) * 0.1
    yl = (torch.rand(B, NC, generator=g) < 0.2).float()
    x, ww, bb = xn.clone().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    logits = F.linear(x, ww, bb)
    loss = F.binary_cross_entropy_with_logits(logits, yl)
    loss.backward()
    dxn_, dw_, db_ = xn.to(DEV), w.to(DEV), bias.to(DEV)
    lo, ls, dl = torch.zeros(B, NC, device=DEV), torch.zeros(1, device=DEV), torch.zeros(B, NC, device=DEV)
    pred = torch.zeros(B, NC, device=DEV, dtype=torch.uint8)
    scratch = torch.zeros(ops.bce_head_scratch_floats(B), device=DEV)
    for _ in range(2):  # the scratch counter must reset itself
        ops.bce_head_fwd(dxn_, dw_, db_, yl.to(DEV), lo, ls, dl, pred, scratch, 0.5, 1.0)
    assert torch.allclose(lo.cpu(), logits.detach(), rtol=1e-4, atol=1e-5)
    assert abs(float(ls) - float(loss)) < 1e-5
    assert torch.equal(pred.cpu().bool(), torch.sigmoid(lo.cpu()) > 0.5)
    gw, gb, gx = torch.zeros(NC, H, device=DEV), torch.zeros(NC, device=DEV), torch.zeros(B, H, device=DEV, dtype=torch.bfloat16)
    ops.bce_head_bwd(dl, dxn_, dw_, gw, gb, gx)
    assert rel(gw.cpu(), ww.grad) < 1e-4 and rel(gb.cpu(), bb.grad) < 1e-4 and rel(gx.float().cpu(), x.grad) < 5e-3
    # forward only (no labels): logits and predictions still produced
    lo2 = torch.zeros_like(lo)
    ops.bce_head_fwd(dxn_, dw_, db_, None, lo2, None, None, pred, scratch, 0.5, 1.0)
    assert torch.equal(lo2, lo)


# ---------------------------------------------------------------------------------------------------------------------
# the fused step
# ---------------------------------------------------------------------------------------------------------------------
def test_state_dict_is_the_reference_layout():
    model = build()
    torch.manual_seed(0)
    ref = G.init_mmimdb_state()
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref.keys()) and len(sd) == 38
    for k in ref:
        assert sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype and torch.equal(sd[k].cpu(), ref[k]), k
    # load_state_dict round trip through the strided (weight | bias column) views
    model2 = build(seed=5)
    model2.load_state_dict(sd)
    for k, v in model2.state_dict().items():
        assert torch.equal(v, sd[k]), k


@pytest.mark.parametrize("B,seed", [(16, 99), (128, 7)])
def test_train_step_matches_oracle(B, seed):
    model = build(graphs=False)
    state = cpu_state(model)
    d = G.synthetic_batch(B, seed)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
    out = model.train_step(make_batch(d), opt, LOSS, torch.device(DEV), None, dropout_masks=d["dropout_masks"])
    eng = model._get_engine(torch.device(DEV))
    plan = eng.plan_for(B)
    ost = OrderedDict((k, v.clone()) for k, v in state.items())
    ref = G.train_step(ost, {}, d["image_masked"], d["text_masked"], d["labels"], d["dropout_masks"], lr=1e-3, weight_decay=1e-3)
    logits = plan.logits.cpu()
    span = float(ref["logits"].max() - ref["logits"].min())
    assert float((logits - ref["logits"]).abs().max()) <= 1e-2 * span, (float((logits - ref["logits"]).abs().max()), span)
    assert abs(out["loss"] - ref["loss"]) < 1e-2 * max(1.0, abs(ref["loss"]))
    agree = float((plan.pred.cpu().long() == ref["predictions"]).float().mean())
    assert agree > 0.98, agree
    # gradients, un-forced: against the fp32 oracle and against the oracle rounded to bf16 where the kernels round
    emu = G.train_step(OrderedDict((k, v.clone()) for k, v in state.items()), {}, d["image_masked"], d["text_masked"], d["labels"],
                       d["dropout_masks"], apply_update=False, emulate_bf16=True)

    def grad_error(reference):
        num = den = 0.0
        worst = ("", 0.0)
        for k, p in model.named_parameters():
            gq, gr = p.grad.detach().cpu().double(), reference[k].double()
            e = float((gq - gr).norm())
            num, den = num + e * e, den + float(gr.norm()) ** 2
            r = e / (float(gr.norm()) + 1e-30)
            if r > worst[1]:
                worst = (k, r)
        return float(np.sqrt(num / den)), worst

    g32, w32 = grad_error(ref["grads"])
    g16, w16 = grad_error(emu["grads"])
    print(f"B={B}: grad rel L2 vs fp32 oracle {g32:.4f} (worst {w32}); vs bf16-rounded oracle {g16:.4f} (worst {w16})")
    assert g32 < 0.12 and w32[1] < 0.2, (g32, w32)
    assert g16 < 4e-2 and w16[1] < 8e-2, (g16, w16)
    # Adam: parameters after the step vs the oracle's (lr 1e-3: the update is +-lr for almost every element)
    sd = model.state_dict()
    for k, v in ost.items():
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v) == 1
        elif k.endswith(("running_mean", "running_var")):
            assert torch.allclose(sd[k].cpu(), v, rtol=2e-2, atol=2e-3), k
        else:
            assert float((sd[k].cpu() - v).abs().max()) <= 2.1e-3, k  # sign flips of ~0 gradients move an element by 2 lr at most
            assert float((sd[k].cpu() - v).abs().mean()) <= 2e-4, k
    if B == 16:  # the committed reference fixture (oracle/make_golden.py ran the unmodified reference on the same inputs)
        gold = np.load(os.path.join(GOLD, "mmimdb_b16.npz"))
        assert np.abs(logits.numpy() - gold["logits"]).max() <= 1e-2 * span
        assert abs(out["loss"] - float(gold["loss"])) < 1e-2
        l2 = dict(zip(gold["grad_keys"], gold["grad_l2"]))
        for k, p in model.named_parameters():
            assert abs(float(p.grad.double().norm()) - l2[k]) <= 0.1 * l2[k] + 1e-9, k


def test_pooled_zero_padding_and_bias_column_stay_clean():
    model = build()
    d = G.synthetic_batch(32, 3)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
    for _ in range(4):
        model.train_step(make_batch(d), opt, LOSS, torch.device(DEV), None)
    fs = model._get_engine(torch.device(DEV)).fs
    for name, n_in in (("image_model.net.1.weight", 4096), ("text_model.net.1.weight", 300)):
        m = fs.aug_matrix(fs.P, name)
        assert float(m[:, n_in + 1:].abs().max()) == 0.0
        assert torch.equal(m[:, n_in], dict(model.named_parameters())[name.replace("weight", "bias")].data)
        assert float(m[:, n_in].abs().max()) > 0.0


def test_eval_and_loss_curve_and_graph():
    B = 128
    d = G.synthetic_batch(B, 11)
    dev = torch.device(DEV)
    losses = {}
    for graphs in (False, True):
        model = build(graphs=graphs)
        state = cpu_state(model)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
        losses[graphs] = [model.train_step(make_batch(d), opt, LOSS, dev, None, dropout_masks=d["dropout_masks"])["loss"] for _ in range(40)]
    # oracle curve
    opt_state, ost = {}, OrderedDict((k, v.clone()) for k, v in state.items())
    ref = [G.train_step(ost, opt_state, d["image_masked"], d["text_masked"], d["labels"], d["dropout_masks"], lr=1e-3, weight_decay=1e-3)["loss"]
           for _ in range(40)]
    for a, b, r in zip(losses[False], losses[True], ref):
        assert abs(a - r) < 2e-2 * max(1.0, r) and abs(b - r) < 2e-2 * max(1.0, r), (a, b, r)
    assert ref[-1] < 0.5 * ref[0] and losses[True][-1] < 0.5 * losses[True][0]
    # validation_step after training: running statistics, no dropout; all three missing patterns are in the batch
    ev = model.validation_step(make_batch(d, device_mask=False), LOSS, dev, None, return_test_info=True)
    sd = cpu_state(model)
    rv = G.validation_step(sd, d["image_masked"], d["text_masked"], d["labels"])
    assert abs(ev["loss"] - rv["loss"]) < 1e-2 * max(1.0, rv["loss"])
    assert float((torch.from_numpy(ev["predictions"]) == rv["predictions"]).float().mean()) > 0.98
    model.eval()
    lg = model(d["image_masked"].to(DEV), d["text_masked"].to(DEV)).cpu()
    span = float(rv["logits"].max() - rv["logits"].min())
    assert float((lg - rv["logits"]).abs().max()) <= 1e-2 * span
    assert set(ev["miss_types"]) == {"it", "i", "t"}
    # encoders stand-alone (get_embeddings path)
    emb = model.image_model(d["image"].to(DEV)).cpu()
    st = sd
    ref_e = F.linear(F.batch_norm(d["image"], st["image_model.net.0.running_mean"], st["image_model.net.0.running_var"],
                                  st["image_model.net.0.weight"], st["image_model.net.0.bias"], False), st["image_model.net.1.weight"],
                     st["image_model.net.1.bias"])
    assert rel(emb, ref_e) < 1e-2


def test_premasked_equals_device_mask_and_own_dropout():
    dev = torch.device(DEV)
    d = G.synthetic_batch(64, 21)
    outs = []
    for device_mask in (True, False):
        model = build()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
        model.train_step(make_batch(d, device_mask), opt, LOSS, dev, None, dropout_masks=d["dropout_masks"])
        outs.append(model._get_engine(dev).plan_for(64).logits.clone())
    assert torch.equal(outs[0], outs[1])
    # the engine's own dropout: about half of the units kept, different masks per layer and per step
    model = build()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
    plan = model._get_engine(dev).plan_for(64)
    model.train_step(make_batch(d), opt, LOSS, dev, None)
    k1, k2 = plan.keep1.clone(), plan.keep2.clone()
    model.train_step(make_batch(d), opt, LOSS, dev, None)
    assert 0.45 < float(k1.float().mean()) < 0.55 and not torch.equal(k1, k2) and not torch.equal(k1, plan.keep1)


def test_unsupported_requests_raise():
    from mml_b200.mmimdb import GatedBiModalNetwork, MLPGenreClassifier, MMIMDb, MMIMDbModalityEncoder

    dev = torch.device(DEV)
    with pytest.raises(ValueError):
        MMIMDb(MMIMDbModalityEncoder(8, 64), MMIMDbModalityEncoder(8, 64), multimodal_pooling={"pooling_type": "bilinear"}, classifier=MLPGenreClassifier(64, 3, 64))
    with pytest.raises(NotImplementedError):
        GatedBiModalNetwork(64, 64, 64, 64, use_bias=True)
    model = build()
    d = G.synthetic_batch(8, 1)
    with pytest.raises(NotImplementedError):
        model.train_step(make_batch(d), torch.optim.SGD(model.parameters(), lr=0.1), LOSS, dev, None)
    bad = {"ce": type("T", (), {"loss_fn": torch.nn.CrossEntropyLoss(), "weight": 1.0})()}
    with pytest.raises(NotImplementedError):
        model.train_step(make_batch(d), torch.optim.Adam(model.parameters()), bad, dev, None)
    with pytest.raises(RuntimeError):
        model.train_step(make_batch(d), torch.optim.Adam(model.parameters()), LOSS, torch.device("cpu"), None)


# ---------------------------------------------------------------------------------------------------------------------
# multimodal_pooling variants (mmimdb_pooling.yaml; pooling.py)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H", [(128, 512), (9, 64)])
@pytest.mark.parametrize("kind", ["max", "avg", "sum"])
@pytest.mark.parametrize("dropout", [True, False])
def test_pool_kernels(B, H, kind, dropout):
    from mml_b200 import ops

    g = torch.Generator().manual_seed(B + H)
    pa, pb = _bf16(torch.randn(B, H, generator=g)), _bf16(torch.randn(B, H, generator=g))
    ba, bb = torch.randn(H, generator=g) * 0.1, torch.randn(H, generator=g) * 0.1
    ka, kb = (torch.rand(B, H, generator=g) >= 0.1), (torch.rand(B, H, generator=g) >= 0.1)
    dz = torch.randn(B, H, generator=g)
    xa, xb = pa.clone().requires_grad_(True), pb.clone().requires_grad_(True)
    va, vb = ba.clone().requires_grad_(True), bb.clone().requires_grad_(True)
    a, b = torch.tanh(xa + va), torch.tanh(xb + vb)
    if dropout:
        a, b = a * ka.float() / 0.9, b * kb.float() / 0.9
    z = torch.max(a, b) if kind == "max" else ((a + b) / 2 if kind == "avg" else a + b)
    z.backward(dz)
    dev = lambda t, dt=None: (t if dt is None else t.to(dt)).to(DEV).contiguous()
    ha, hb = torch.zeros(B, H, device=DEV), torch.zeros(B, H, device=DEV)
    dka, dkb = (dev(ka, torch.uint8), dev(kb, torch.uint8)) if dropout else (None, None)
    ops.pool_fwd(dev(pa, torch.bfloat16), dev(pb, torch.bfloat16), dev(ba), dev(bb), dka, dkb, 1 / 0.9, ha, hb)
    assert torch.allclose(ha.cpu(), a.detach(), rtol=1e-5, atol=1e-6) and torch.allclose(hb.cpu(), b.detach(), rtol=1e-5, atol=1e-6)
    mix = {"max": (1.0, 1.0), "avg": (0.5, 0.5), "sum": (1.0, 1.0)}[kind]
    da, db_ = torch.zeros(B, H, device=DEV, dtype=torch.bfloat16), torch.zeros(B, H, device=DEV, dtype=torch.bfloat16)
    gba, gbb = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    ops.pool_bwd(dev(dz), ha, hb, dka, dkb, 1 / 0.9, 0 if kind == "max" else 1, mix[0], mix[1], da, db_, gba, gbb)
    assert rel(da.float().cpu(), xa.grad) < 5e-3 and rel(db_.float().cpu(), xb.grad) < 5e-3
    assert rel(gba.cpu(), va.grad) < 5e-3 and rel(gbb.cpu(), vb.grad) < 5e-3  # sums of the bf16-rounded dpre


def build_pooling(pooling_type, graphs=True):
    from mml_b200.mmimdb import MLPGenreClassifier, MMIMDb, MMIMDbModalityEncoder

    torch.manual_seed(0)
    img, txt = MMIMDbModalityEncoder(4096, 512), MMIMDbModalityEncoder(300, 512)
    clf = MLPGenreClassifier(512, 23, 512)
    model = MMIMDb(img, txt, multimodal_pooling={"pooling_type": pooling_type, "hidden_dim": 512, "dropout": 0.1}, classifier=clf).to(DEV)
    model._get_engine(torch.device(DEV)).use_graphs = graphs
    return model


@pytest.mark.parametrize("B,H,Hd,NS", [(128, 512, 512, 2), (128, 512, 512, 1), (7, 64, 128, 2)])
def test_attention_pool_kernels(B, H, Hd, NS):
    """att_fwd / att_bwd / pool_bwd(gate, dcomb) against torch autograd of pooling.py:113-126 (given the layer-0 GEMM output)."""
    from mml_b200 import ops

    g = torch.Generator().manual_seed(B + NS)
    a0, b0_ = torch.tanh(torch.randn(B, H, generator=g)), torch.tanh(torch.randn(B, H, generator=g))
    hid = _bf16(torch.randn(B, Hd, generator=g))
    bias0, w2, bias2 = torch.randn(Hd, generator=g) * 0.1, torch.randn(NS, Hd, generator=g) * 0.1, torch.randn(NS, generator=g) * 0.1
    dz = torch.randn(B, H, generator=g)
    a, b = a0.clone().requires_grad_(True), b0_.clone().requires_grad_(True)
    hd, v0, vw, v2 = hid.clone().requires_grad_(True), bias0.clone().requires_grad_(True), w2.clone().requires_grad_(True), bias2.clone().requires_grad_(True)
    t = torch.tanh(hd + v0)
    s = t @ vw.t() + v2
    gate = torch.softmax(s, 1)[:, 0:1] if NS == 2 else torch.sigmoid(s)
    z = gate * a + (1 - gate) * b
    z.backward(dz)
    dev = lambda x, dt=None: (x if dt is None else x.to(dt)).to(DEV).contiguous()
    tt, gt = torch.zeros(B, Hd, device=DEV), torch.zeros(B, device=DEV)
    ops.att_fwd(dev(hid, torch.bfloat16), dev(bias0), dev(w2), dev(bias2), tt, gt)
    assert torch.allclose(gt.cpu(), gate.detach().reshape(-1), rtol=1e-5, atol=1e-6) and torch.allclose(tt.cpu(), t.detach(), rtol=1e-5, atol=1e-6)
    dw2, db2, db0 = torch.zeros(NS, Hd, device=DEV), torch.zeros(NS, device=DEV), torch.zeros(Hd, device=DEV)
    dhid = torch.zeros(B, Hd, device=DEV, dtype=torch.bfloat16)
    ha, hb = dev(a0), dev(b0_)
    ops.att_bwd(dev(dz), ha, hb, gt, tt, dev(w2), dw2, db2, db0, dhid)
    assert rel(dw2.cpu(), vw.grad) < 1e-4 and rel(db2.cpu(), v2.grad) < 1e-4 and rel(db0.cpu(), v0.grad) < 1e-4
    assert rel(dhid.float().cpu(), hd.grad) < 5e-3
    # the branch gradients: per-sample mix (da = dz g, db = dz (1 - g)), no dropout; tanh backward is applied by pool_bwd, so compare
    # against a.grad * (1 - a^2)
    da, db_ = torch.zeros(B, H, device=DEV, dtype=torch.bfloat16), torch.zeros(B, H, device=DEV, dtype=torch.bfloat16)
    gba, gbb = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    ops.pool_bwd(dev(dz), ha, hb, None, None, 1.0, 1, 0.0, 0.0, da, db_, gba, gbb, gate=gt, dcomb=None)
    assert rel(da.float().cpu(), a.grad * (1 - a0 * a0)) < 5e-3 and rel(db_.float().cpu(), b.grad * (1 - b0_ * b0_)) < 5e-3


@pytest.mark.parametrize("pooling_type", ["max", "avg", "sum", "attention", "gated"])
def test_pooling_variants_match_oracle(pooling_type):
    B, seed = 16, 5
    model = build_pooling(pooling_type, graphs=False)
    torch.manual_seed(0)
    init = G.init_mmimdb_pooling_state(pooling_type)
    sd = model.state_dict()
    assert list(sd.keys()) == list(init.keys())
    for k in init:
        assert torch.equal(sd[k].cpu(), init[k]), k
    state = cpu_state(model)
    d = G.synthetic_batch(B, seed)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
    out = model.train_step(make_batch(d), opt, LOSS, torch.device(DEV), None, dropout_masks=d["dropout_masks"], pool_masks=d["pool_masks"])
    plan = model._get_engine(torch.device(DEV)).plan_for(B)
    kw = dict(pooling_type=pooling_type, pool_masks=d["pool_masks"], pool_p=0.1, apply_update=False)
    ref = G.train_step(OrderedDict((k, v.clone()) for k, v in state.items()), {}, d["image_masked"], d["text_masked"], d["labels"], d["dropout_masks"], **kw)
    emu = G.train_step(OrderedDict((k, v.clone()) for k, v in state.items()), {}, d["image_masked"], d["text_masked"], d["labels"], d["dropout_masks"],
                       emulate_bf16=True, **kw)
    logits = plan.logits.cpu()
    span = float(ref["logits"].max() - ref["logits"].min())
    assert float((logits - ref["logits"]).abs().max()) <= 1.5e-2 * span
    assert abs(out["loss"] - ref["loss"]) < 1e-2

    def grad_error(reference):
        num = den = 0.0
        for k, p in model.named_parameters():
            gq, gr = p.grad.detach().cpu().double(), reference[k].double()
            num, den = num + float((gq - gr).norm()) ** 2, den + float(gr.norm()) ** 2
        return float(np.sqrt(num / den))

    g32, g16 = grad_error(ref["grads"]), grad_error(emu["grads"])
    print(f"pooling {pooling_type}: grad rel L2 vs fp32 oracle {g32:.4f}, vs bf16-rounded oracle {g16:.4f}")
    assert g32 < 0.15 and g16 < 5e-2, (g32, g16)
    fixture = os.path.join(GOLD, f"mmimdb_pool_{pooling_type}_b16.npz")
    if os.path.exists(fixture):  # written from the unmodified reference (max, sum)
        gold = np.load(fixture)
        assert np.abs(logits.numpy() - gold["logits"]).max() <= 1.5e-2 * span and abs(out["loss"] - float(gold["loss"])) < 1e-2
    # a few more steps with the engine's own dropout through the CUDA graph, then eval against the oracle on the trained weights
    model._get_engine(torch.device(DEV)).use_graphs = True
    for _ in range(4):
        model.train_step(make_batch(d), opt, LOSS, torch.device(DEV), None)
    ev = model.validation_step(make_batch(d, device_mask=False), LOSS, torch.device(DEV), None, return_test_info=True)
    rv = G.validation_step(cpu_state(model), d["image_masked"], d["text_masked"], d["labels"], pooling_type=pooling_type)
    assert abs(ev["loss"] - rv["loss"]) < 1e-2 * max(1.0, rv["loss"])
