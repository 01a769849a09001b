"""GPU parity of the MOSI / UttFusion step (config 4; mml_b200.utt_fusion) against the CPU oracle and the reference fixture.

Precision contract: the LSTMs, the dense layers, the loss and the optimizer are fp32 (tolerances 1e-4 .. 1e-5); the three TextCNN
convolutions run on the bf16 tensor-core path (x and W rounded to bf16, fp32 accumulation), and the max over time can pick a
different position when two candidates are within one bf16 ulp -- so the step is compared at: logits <= 2e-2 of the logit range,
loss <= 1e-2, gradients of the fp32 sub-networks <= 3e-2 relative L2, TextCNN convolution gradients <= 0.15.
"""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import utt_fusion_oracle as U

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class Term:
    def __init__(self):
        self.loss_fn, self.weight = torch.nn.CrossEntropyLoss(), 1.0


LOSS = {"cross_entropy": Term()}


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def build(graphs=True, clip=1.0):
    from mml_b200.utt_fusion import FcClassifier, LSTMEncoder, TextCNN, UttFusionModel

    torch.manual_seed(0)
    model = UttFusionModel(LSTMEncoder(5, 64, "last"), LSTMEncoder(20, 64, "last"),
                           TextCNN(768, embd_size=64, dropout=0.5, in_channels=1, out_channels=128, kernel_heights=[3, 4, 5]),
                           FcClassifier(192, [192, 64, 32], 3, dropout=0.5), clip=clip).to(DEV)
    model._get_engine(torch.device(DEV)).use_graphs = graphs
    return model


def make_batch(d, device_mask=True):
    if device_mask:
        return {"audio_original": d["audio"], "audio_missing_index": d["audio_mask"], "video_original": d["video"], "video_missing_index": d["video_mask"],
                "text_original": d["text"], "text_missing_index": d["text_mask"], "label": d["labels"], "pattern_name": d["pattern_name"]}
    return {"audio": d["audio_masked"], "video": d["video_masked"], "text": d["text_masked"], "label": d["labels"], "pattern_name": d["pattern_name"]}


def cpu_state(model):
    return OrderedDict((k, v.detach().cpu().clone().contiguous()) for k, v in model.state_dict().items())


@pytest.mark.parametrize("B,T,IN", [(32, 50, 5), (32, 50, 20), (3, 7, 32), (2, 70, 5)])
def test_lstm_kernels(B, T, IN):
    from mml_b200 import ops

    torch.manual_seed(B + IN)
    lstm = torch.nn.LSTM(IN, 64, batch_first=True)
    x = torch.randn(B, T, IN)
    dh = torch.randn(B, 64)
    out, (hn, _) = lstm(x)
    hn.squeeze(0).backward(dh)
    w = [p.detach().to(DEV).contiguous() for p in (lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0)]
    dx = x.to(DEV)
    gates, cs, hs = torch.zeros(B, T, 256, device=DEV), torch.zeros(B, T, 64, device=DEV), torch.zeros(B, T, 64, device=DEV)
    hl = torch.zeros(B, 64, device=DEV)
    ops.lstm_fwd(dx, *w, gates, cs, hs, hl)
    assert torch.allclose(hl.cpu(), hn.squeeze(0).detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(hs.cpu(), out.detach(), rtol=1e-4, atol=1e-5)
    dw = [torch.zeros_like(t) for t in w]
    ops.lstm_bwd(dx, w[1], gates, cs, hs, dh.to(DEV), *dw)
    for got, ref in zip(dw, (lstm.weight_ih_l0.grad, lstm.weight_hh_l0.grad, lstm.bias_ih_l0.grad, lstm.bias_hh_l0.grad)):
        assert rel(got.cpu(), ref) < 1e-4


def test_relumax_dense_softmax_clip_kernels():
    from mml_b200 import ops

    g = torch.Generator().manual_seed(3)
    B, P, C = 8, 11, 128
    conv = torch.randn(B, P, C, generator=g).to(torch.bfloat16)
    conv[0, :, 5] = -1.0  # a channel whose ReLU never fires
    bias = torch.randn(C, generator=g) * 0.1
    keep = torch.rand(B, 3 * C, generator=g) >= 0.5
    dy = torch.randn(B, 3 * C, generator=g)
    cv, bv = conv.float().clone().requires_grad_(True), bias.clone().requires_grad_(True)
    y = (F.relu(cv + bv).max(dim=1).values * keep[:, C:2 * C].float() / 0.5)
    y.backward(dy[:, C:2 * C])
    yb, arg = torch.zeros(B, 3 * C, device=DEV), torch.zeros(B, 3 * C, device=DEV, dtype=torch.int32)
    dkeep = keep.to(torch.uint8).to(DEV)
    ops.relumax_fwd(conv.to(DEV), bias.to(DEV), dkeep, 2.0, yb, arg, C)
    assert torch.allclose(yb[:, C:2 * C].cpu(), y.detach(), rtol=1e-5, atol=1e-6) and float(yb[:, :C].abs().max()) == 0.0
    assert int(arg[0, C + 5]) == -1
    dconv, dbias = torch.ones(B, P, C, device=DEV, dtype=torch.bfloat16), torch.zeros(C, device=DEV)
    ops.relumax_bwd(dy.to(DEV), arg, dkeep, 2.0, dconv, dbias, C)
    assert rel(dconv.float().cpu(), cv.grad) < 5e-3 and rel(dbias.cpu(), bv.grad) < 1e-5
    # dense layer with ReLU + dropout, writing into a column slice
    Bn, K, N = 32, 192, 64
    x, w, b = torch.randn(Bn, K, generator=g), torch.randn(N, K, generator=g) * 0.1, torch.randn(N, generator=g) * 0.1
    kp = torch.rand(Bn, N, generator=g) >= 0.5
    dyl = torch.randn(Bn, N, generator=g)
    xv, wv, bv2 = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yl = F.relu(F.linear(xv, wv, bv2)) * kp.float() / 0.5
    yl.backward(dyl)
    wide = torch.zeros(Bn, N + 10, device=DEV)
    dx_, dw_, db_ = torch.zeros(Bn, K, device=DEV), torch.zeros(N, K, device=DEV), torch.zeros(N, device=DEV)
    dk = kp.to(torch.uint8).to(DEV)
    xs, ws, bs = x.to(DEV), w.to(DEV), b.to(DEV)
    ops.dense_fwd(xs, K, ws, bs, dk, 2.0, True, wide[:, 10:], N + 10, Bn)
    assert torch.allclose(wide[:, 10:].cpu(), yl.detach(), rtol=1e-4, atol=1e-5) and float(wide[:, :10].abs().max()) == 0.0
    dyd = dyl.to(DEV).clone()
    ops.dense_bwd(dyd, wide[:, 10:], N + 10, dk, 2.0, True, xs, K, ws, dx_, K, dw_, db_, Bn)
    assert rel(dx_.cpu(), xv.grad) < 1e-4 and rel(dw_.cpu(), wv.grad) < 1e-4 and rel(db_.cpu(), bv2.grad) < 1e-4
    # softmax-CE
    lg, lab = torch.randn(Bn, 3, generator=g), torch.randint(0, 3, (Bn,), generator=g)
    lv = lg.clone().requires_grad_(True)
    ls = F.cross_entropy(lv, lab)
    ls.backward()
    dl, rl, lo, pr = torch.zeros(Bn, 3, device=DEV), torch.zeros(Bn, device=DEV), torch.zeros(1, device=DEV), torch.zeros(Bn, device=DEV, dtype=torch.int32)
    ops.softmax_ce(lg.to(DEV), lab.to(DEV), dl, rl, lo, pr, 1.0)
    assert abs(float(lo) - float(ls.detach())) < 1e-5 and rel(dl.cpu(), lv.grad) < 1e-5 and torch.equal(pr.cpu().long(), lg.argmax(1))
    # gradient-norm clip scale
    grad = torch.randn(100000, generator=g) * 0.01
    hyper = torch.zeros(8, 8, device=DEV)
    partial, nrm = torch.zeros(256, device=DEV, dtype=torch.float64), torch.zeros(1, device=DEV)
    ops.clip_grad_scale(grad.to(DEV), 1.0, 1.0, hyper, 8, partial, nrm)
    n = float(grad.double().norm())
    assert abs(float(nrm) - n) < 1e-4 * n and abs(float(hyper[0, 5]) - min(1.0, 1.0 / (n + 1e-6))) < 1e-5 and float(hyper[7, 5]) == float(hyper[0, 5])


@pytest.mark.parametrize("B,seed", [(8, 21), (32, 4)])
def test_utt_step_matches_oracle(B, seed):
    model = build(graphs=False)
    torch.manual_seed(0)
    init = U.init_utt_state()
    sd = model.state_dict()
    assert list(sd.keys()) == list(init.keys()) and len(sd) == 24
    for k in init:
        assert torch.equal(sd[k].cpu(), init[k]), k
    state = cpu_state(model)
    d = U.synthetic_batch(B, seed)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
    out = model.train_step(make_batch(d), opt, LOSS, torch.device(DEV), None, dropout_masks=d["keeps"])
    plan = model._engine.plan_for(B, 50)
    ref = U.train_step(OrderedDict((k, v.clone()) for k, v in state.items()), {}, d["audio_masked"], d["video_masked"], d["text_masked"], d["labels"],
                       d["keeps"], clip=None, apply_update=False)
    logits = plan.logits.cpu()
    span = float(ref["logits"].max() - ref["logits"].min())
    assert float((logits - ref["logits"]).abs().max()) <= 2e-2 * span, (float((logits - ref["logits"]).abs().max()), span)
    assert abs(out["loss"] - ref["loss"]) < 1e-2
    emu = U.train_step(OrderedDict((k, v.clone()) for k, v in state.items()), {}, d["audio_masked"], d["video_masked"], d["text_masked"], d["labels"],
                       d["keeps"], clip=None, apply_update=False, emulate_bf16=True)

    def group_errors(reference):
        groups = {"netA": [], "netV": [], "netC": [], "netT.embd": [], "netT.conv": []}
        for k, p in model.named_parameters():
            key = next(g for g in ("netT.embd", "netT.conv", "netA", "netV", "netC") if k.startswith(g))
            groups[key].append((p.grad.detach().cpu().double().reshape(-1), reference[k].double().reshape(-1)))
        return {g: round(rel(torch.cat([a for a, _ in v]), torch.cat([b for _, b in v])), 4) for g, v in groups.items()}

    e32, e16 = group_errors(ref["grads"]), group_errors(emu["grads"])
    print(f"B={B}: unclipped gradient rel L2 by sub-network vs fp32 oracle {e32}; vs bf16-rounded oracle {e16}; grad norm {float(plan.grad_norm):.4f} / {ref['grad_norm']:.4f}")
    # against the oracle rounded where the kernels round (text input, conv weights / outputs / output gradients in bf16): accumulation order only
    assert max(e16.values()) < 1e-2, e16  # measured 1e-4
    # against the fp32 oracle: the bf16 convolutions move the max-over-time winners; on these inputs the operand-rounded oracle itself is
    # 1-6 % (B = 32) to 6-12 % (B = 8) away from its fp32 self
    assert max(e32.values()) < (0.4 if B < 16 else 0.15), e32
    assert abs(float(plan.grad_norm) - ref["grad_norm"]) < 5e-2 * ref["grad_norm"]
    # the fused Adam applied torch's rule to the GPU's own gradients scaled by clip / (norm + 1e-6)
    coef = min(1.0, 1.0 / (float(plan.grad_norm) + 1e-6))
    for k, p in model.named_parameters():
        g = p.grad.detach().cpu() * coef + 1e-3 * state[k]
        upd = state[k] - (1e-3 / 0.1) * (0.1 * g) / ((0.001 * g * g).sqrt() / (0.001 ** 0.5) + 1e-8)
        assert torch.allclose(p.detach().cpu(), upd, rtol=1e-4, atol=2e-6), k
    if B == 8:  # the reference's own run (oracle/make_golden.py)
        gold = np.load(os.path.join(GOLD, "mosi_b8.npz"))
        assert np.abs(logits.numpy() - gold["logits"]).max() <= 2e-2 * span and abs(out["loss"] - float(gold["losses"][0])) < 1e-2


def test_utt_training_curve_eval_graph_and_masks():
    B, steps = 32, 40
    d = U.synthetic_batch(B, 9)
    dev = torch.device(DEV)
    curves = {}
    for graphs in (False, True):
        model = build(graphs=graphs)
        state = cpu_state(model)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
        curves[graphs] = [model.train_step(make_batch(d), opt, LOSS, dev, None, dropout_masks=d["keeps"])["loss"] for _ in range(steps)]
    opt_state, ost, ref = {}, OrderedDict((k, v.clone()) for k, v in state.items()), []
    for _ in range(steps):
        ref.append(U.train_step(ost, opt_state, d["audio_masked"], d["video_masked"], d["text_masked"], d["labels"], d["keeps"])["loss"])
    sm = lambda v: np.convolve(v, np.ones(5) / 5, mode="valid")
    assert abs(curves[False][0] - ref[0]) < 1e-2 and abs(curves[True][0] - ref[0]) < 1e-2
    assert np.abs(sm(curves[True]) - sm(ref)).max() < 0.1 * ref[0] and np.abs(sm(curves[False]) - sm(ref)).max() < 0.1 * ref[0], (curves[True][-5:], ref[-5:])
    assert ref[-1] < 0.7 * ref[0] and curves[True][-1] < 0.7 * curves[True][0]
    ev = model.validation_step(make_batch(d, device_mask=False), LOSS, dev, None, return_test_info=True)
    rv = U.validation_step(cpu_state(model), d["audio_masked"], d["video_masked"], d["text_masked"], d["labels"])
    assert abs(ev["loss"] - rv["loss"]) < 2e-2 * max(1.0, rv["loss"])
    assert float((torch.from_numpy(ev["predictions"]) == rv["predictions"]).float().mean()) > 0.9
    assert len(set(ev["miss_types"])) >= 4
    # pre-masked batch == device-side mask; own dropout path runs
    outs = []
    for device_mask in (True, False):
        m2 = build()
        o2 = torch.optim.Adam(m2.parameters(), lr=1e-3, weight_decay=1e-3)
        m2.train_step(make_batch(d, device_mask), o2, LOSS, dev, None, dropout_masks=d["keeps"])
        outs.append(m2._engine.plan_for(B, 50).logits.clone())
    assert torch.equal(outs[0], outs[1])
    m3 = build()
    o3 = torch.optim.Adam(m3.parameters(), lr=1e-3, weight_decay=1e-3)
    l0 = [m3.train_step(make_batch(d), o3, LOSS, dev, None)["loss"] for _ in range(6)]
    assert all(np.isfinite(l0)) and 0.4 < float(m3._engine.plan_for(B, 50).keepT.float().mean()) < 0.6


def test_utt_unsupported_requests_raise():
    from mml_b200.utt_fusion import FcClassifier, LSTMEncoder, TextCNN

    with pytest.raises(NotImplementedError):
        LSTMEncoder(5, 64, "attention")
    with pytest.raises(NotImplementedError):
        LSTMEncoder(5, 128, "last")
    with pytest.raises(NotImplementedError):
        FcClassifier(192, [64], 3, use_bn=True)
    with pytest.raises(NotImplementedError):
        TextCNN(768, kernel_heights=[3, 4])
    model = build()
    d = U.synthetic_batch(4, 1)
    with pytest.raises(NotImplementedError):
        model.train_step(make_batch(d), torch.optim.SGD(model.parameters(), lr=0.1), LOSS, torch.device(DEV), None)
