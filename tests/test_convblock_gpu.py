"""GPU parity of the ConvBlock AVMNIST path (AVMNIST(MNISTAudio, MNISTImage, 128), the model of
configs/avmnist/centralised/train_avmnist.yaml) against the CPU oracle and the reference-generated fixture
tests/golden/avmnist_convblock_b4.npz (tests/golden/make_fixtures.py ran MML_Suite's own classes).

Kernels first (csrc/convblock.cu against torch fp32 on the same bf16-rounded operands: bit-level layout / index checks), then
the whole step.  The network is only four convolutions deep, so -- unlike the ResNet path (tests/test_step_gpu.py) -- the
un-forced comparison is well conditioned; tolerances are ~2x the measured bf16-storage error and written next to each assert.
"""
import copy
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import late_fusion_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Term:
    def __init__(self):
        self.loss_fn, self.weight = torch.nn.CrossEntropyLoss(), 1.0


LOSS = {"cross_entropy": Term()}


def bf(t):
    return t.to(torch.bfloat16).float()


def nchw(t):
    return t.detach().float().permute(0, 3, 1, 2).contiguous()


# ---------------------------------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,K", [(3, 32, 94, 32), (2, 28, 28, 32), (1, 5, 7, 8), (2, 9, 33, 64), (256, 32, 94, 32)])
def test_conv3x3_c1_fprop_and_wgrad(B, H, W, K):
    from mml_b200 import ops

    g = torch.Generator().manual_seed(B * 1000 + H)
    x = torch.randn(B, H, W, generator=g).to(DEV)
    mask = (torch.rand(B, generator=g) > 0.3).float().to(DEV)
    w = (torch.randn(K, 1, 3, 3, generator=g) * 0.3).to(DEV)
    y = torch.full((B, H, W, 64), 7.0, device=DEV, dtype=torch.bfloat16)
    stats = ops.bn_stats_buffer(64, DEV)
    ops.conv3x3_c1_fprop(x, mask, w.reshape(K, 9), y, stats, K)
    xm = bf(x * mask[:, None, None])
    ref = F.conv2d(xm[:, None], bf(w), padding=1)
    got = nchw(y)
    assert torch.equal(got[:, K:], torch.zeros_like(got[:, K:])), "padded channels must be exact zeros"
    # fp32 accumulation of 9 products, one bf16 rounding: half an ulp of bf16 plus accumulation-order noise
    assert (got[:, :K] - ref).abs().max() <= 2 ** -8 * ref.abs().max() + 1e-6
    s = stats.view(-1, 64, 2).sum(0)
    stored = got.double()
    assert torch.allclose(s[:, 0], stored.sum((0, 2, 3)), rtol=1e-6, atol=1e-3)
    assert torch.allclose(s[:, 1], (stored * stored).sum((0, 2, 3)), rtol=1e-6, atol=1e-3)
    # unmasked variant == mask of ones
    y2 = torch.empty_like(y)
    ops.conv3x3_c1_fprop(x, None, w.reshape(K, 9), y2, None, K)
    ops.conv3x3_c1_fprop(x, torch.ones(B, device=DEV), w.reshape(K, 9), y, None, K)
    assert torch.equal(y, y2)
    # weight gradient
    dy = torch.zeros(B, H, W, 64, device=DEV, dtype=torch.bfloat16)
    dy[..., :K] = (torch.randn(B, H, W, K, generator=g) * 0.1).to(DEV).to(torch.bfloat16)
    ws = torch.zeros(ops.conv3x3_c1_wgrad_workspace(x, K) // 4, device=DEV)
    dw = torch.full((K, 9), 3.0, device=DEV)
    ops.conv3x3_c1_wgrad(x, mask, dy, dw, ws, K)
    wr = bf(w).clone().requires_grad_(True)
    (F.conv2d(xm[:, None].double(), wr.double(), padding=1) * nchw(dy)[:, :K].double()).sum().backward()
    want = wr.grad.reshape(K, 9)
    assert (dw - want).abs().max() <= 2e-5 * max(1.0, float(want.abs().max())) * max(1.0, (B * H * W) ** 0.5 / 16)
    dw2 = torch.empty_like(dw)
    ops.conv3x3_c1_wgrad(x, mask, dy, dw2, ws, K)
    assert torch.equal(dw, dw2), "the weight gradient must be bit-reproducible (fixed-order reduction)"


@pytest.mark.parametrize("B,H,W,k", [(3, 32, 94, 2), (2, 16, 47, 3), (2, 28, 28, 2), (1, 7, 9, 3), (64, 16, 47, 3)])
def test_maxpool_k_forward_backward(B, H, W, k):
    from mml_b200 import ops

    g = torch.Generator().manual_seed(H * 100 + W)
    x = torch.randn(B, H, W, 64, generator=g).to(DEV).to(torch.bfloat16)
    x[:, ::3, ::2] = x[:, :1, :1]  # ties: torch keeps the first maximum of a window
    P, Q = H // k, W // k
    y = torch.empty(B, P, Q, 64, device=DEV, dtype=torch.bfloat16)
    flat = torch.empty(B, 64 * P * Q, device=DEV)
    am = torch.empty(B, P, Q, 64, device=DEV, dtype=torch.uint8)
    ops.maxpool_k_fwd(x, y, flat, am, k)
    xr = nchw(x).requires_grad_(True)
    ref = F.max_pool2d(xr, kernel_size=k)
    assert torch.equal(nchw(y), ref.detach())
    assert torch.equal(flat, torch.flatten(ref.detach(), 1))
    dy = torch.randn(B, 64, P, Q, generator=g).to(DEV).to(torch.bfloat16).float()
    ref.backward(dy)
    dx = torch.full((B, H, W, 64), 5.0, device=DEV, dtype=torch.bfloat16)
    ops.maxpool_k_bwd(dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), None, am, dx, k)
    assert torch.equal(nchw(dx), xr.grad)
    dx2 = torch.empty_like(dx)
    ops.maxpool_k_bwd(None, torch.flatten(dy, 1).contiguous(), am, dx2, k)
    assert torch.equal(dx, dx2)


def test_conv_bias_fold():
    from mml_b200 import ops

    b = torch.randn(32, device=DEV)
    rm = torch.randn(64, device=DEV)
    rm0 = rm.clone()
    ops.bn_conv_bias_fold(b, 0.1, running_mean=rm)
    assert torch.allclose(rm[:32], rm0[:32] + 0.1 * b, atol=1e-7) and torch.equal(rm[32:], rm0[32:])
    scale, shift = torch.randn(64, device=DEV), torch.randn(64, device=DEV)
    s0 = shift.clone()
    ops.bn_conv_bias_fold(b, 0.0, scale=scale, shift=shift)
    assert torch.allclose(shift[:32], s0[:32] + b * scale[:32], atol=1e-6) and torch.equal(shift[32:], s0[32:])


# ---------------------------------------------------------------------------------------------------------------------
# the step
# ---------------------------------------------------------------------------------------------------------------------
def build(dropout=0.5, graphs=False):
    from mml_b200.avmnist import AVMNIST
    from mml_b200.convblock import ConvBlockArgs as A
    from mml_b200.convblock import MNISTAudio, MNISTImage

    torch.manual_seed(0)
    au = MNISTAudio(A(1, 32), A(32, 32), A(32, 64), A(64, 64), 64)
    im = MNISTImage(A(1, 32), A(32, 64), A(64, 64), A(64, 64), 128)
    model = AVMNIST(au, im, 128, dropout=dropout).to(DEV)
    eng = model._get_engine(torch.device(DEV))
    eng.use_graphs = graphs
    return model


def make_batch(d, B):
    return {"audio_original": d["audio"], "audio_missing_index": d["audio_mask"], "image_original": d["image"],
            "image_missing_index": d["image_mask"], "labels": d["labels"], "pattern_name": ["ai"] * B}


def grads_by_name(model):
    eng = model._engine
    return {n: eng.fs._view(eng.fs.G, n, p).detach().cpu().clone() for n, p in model.named_parameters()}


def test_reference_golden_fixture_convblock():
    """Two steps on the fixture's inputs: loss, logits and per-parameter gradient norms of the REFERENCE's own classes."""
    g = np.load(os.path.join(GOLD, "avmnist_convblock_b4.npz"))
    batch, seed, steps = (int(v) for v in g["meta"])
    model = build()
    d = O.synthetic_batch(batch, seed, (32, 94))
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
    losses = []
    for step in range(steps):
        out = model.train_step(make_batch(d, batch), opt, LOSS, torch.device(DEV), None, dropout_mask=d["dropout_mask"])
        losses.append(out["loss"])
        if step == 0:
            plan = next(iter(model._engine.plans.values()))
            logits = plan.logits.detach().cpu().numpy()
            span = float(np.abs(g["logits"]).max())
            assert np.abs(logits - g["logits"]).max() < 0.06 * span + 0.02  # bf16 activations, batch of 4 (BatchNorm over few samples)
            mine = grads_by_name(model)
            for k, want in zip(g["grad_keys"], g["grad_l2"]):
                k = str(k)
                if k.endswith("conv_one.bias") or k.endswith("conv_two.bias"):
                    assert float(mine[k].norm()) == 0.0 and want < 1e-6  # cancels through BatchNorm: the reference holds rounding noise
                    continue
                assert abs(float(mine[k].double().norm()) - want) <= 0.15 * want + 1e-6, (k, float(mine[k].norm()), want)
    assert abs(losses[0] - float(g["losses"][0])) < 2e-2
    assert abs(losses[1] - float(g["losses"][1])) < 5e-2


def forced_from_plan(plan):
    """Activations the GPU stored, under the oracle's tap names (the conv bias is not part of the stored conv output: it cancels in
    the BatchNorm that follows, so forcing the un-biased value leaves every later tensor and every gradient unchanged)."""
    forced = {}
    for pre, ep in (("audio_encoder.", plan.audio), ("image_encoder.", plan.image)):
        for name, t in ep.taps.items():
            forced[pre + name] = nchw(t).cpu()
    return forced


@pytest.mark.parametrize("B", [8, 64, 256])
def test_convblock_step_against_oracle(B):
    """B = 256 is the batch bench.py --workload convblock times."""
    model = build()
    torch.manual_seed(0)
    state = O.init_convblock_avmnist_state()
    d = O.synthetic_batch(B, 77, (32, 94))
    A, I = O.apply_missing_mask(d["audio"], d["audio_mask"]), O.apply_missing_mask(d["image"], d["image_mask"])
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
    out = model.train_step(make_batch(d, B), opt, LOSS, torch.device(DEV), None, dropout_mask=d["dropout_mask"])
    plan = next(iter(model._engine.plans.values()))
    mine = grads_by_name(model)
    # 1. teacher forced: the oracle's backward over the activations the GPU stored.  (a) rounding activation-gradients to bf16 where the
    #    kernels store them -> only accumulation-order noise is left; (b) the plain fp32 backward -> the bf16 gradient storage itself
    #    (grows with the number of positions a weight gradient sums over: the first convolution at batch 256 sums 770k pixels)
    forced = forced_from_plan(plan)
    rels = {}
    for tag, eb in (("forced+bf16", True), ("forced", False)):
        ref_f = O.convblock_train_step(copy.deepcopy(state), {}, A, I, d["labels"], d["dropout_mask"], 0.5, apply_update=False,
                                       emulate_bf16=eb, forced=forced)
        for k, gref in ref_f["grads"].items():
            if k.endswith("conv_one.bias") or k.endswith("conv_two.bias"):
                # cancels through BatchNorm: exactly zero here, fp32 rounding noise in the reference (bf16 rounding noise when emulated)
                assert float(mine[k].abs().max()) == 0.0 and float(gref.abs().max()) < (1e-2 if eb else 1e-5)
                continue
            rels[(tag, k)] = float((mine[k] - gref).norm() / (gref.norm() + 1e-12))
    for (tag, k), rel in sorted(rels.items(), key=lambda kv: -kv[1])[:6]:
        print(f"B={B} {tag:12s} {k:45s} {rel:.4f}")
    worst_f = max(v for (tag, _), v in rels.items() if tag == "forced")
    worst_e = max(v for (tag, _), v in rels.items() if tag == "forced+bf16")
    assert worst_e < 2e-2, worst_e
    assert worst_f < 0.1, worst_f
    # 2. un-forced against the fp32 oracle: loss, logits, and a looser gradient bound (ReLU / max-pool decisions flip on a few elements)
    ref = O.convblock_train_step(state, {}, A, I, d["labels"], d["dropout_mask"], 0.5)
    assert abs(out["loss"] - ref["loss"]) < 2e-2
    span = float(ref["logits"].abs().max())
    assert float((plan.logits.cpu() - ref["logits"]).abs().max()) < 0.05 * span + 0.01
    worst = 0.0
    for k, gref in ref["grads"].items():
        if k.endswith("conv_one.bias") or k.endswith("conv_two.bias"):
            continue
        rel = float((mine[k] - gref).norm() / (gref.norm() + 1e-12))
        worst = max(worst, rel)
        assert rel < 0.4, ("unforced", k, rel)
    print(f"B={B}: worst gradient rel L2 forced+bf16 {worst_e:.4f} forced {worst_f:.4f} un-forced {worst:.4f}")
    # parameters after the Adam update (the oracle applied its own), BatchNorm running statistics incl. the folded conv bias
    sd = model.state_dict()
    for k, v in state.items():
        got = sd[k].detach().cpu()
        if k.endswith("num_batches_tracked"):
            assert int(got) == int(v) == 1
        elif "running_" in k:
            assert torch.allclose(got, v, rtol=2e-2, atol=2e-3), k
        else:
            # one Adam step moves every weight by at most lr = 5e-4
            assert float((got - v).abs().max()) <= 1.1e-3, k


def test_convblock_eval_and_standalone_encoder():
    model = build(dropout=0.0)
    torch.manual_seed(0)
    state = O.init_convblock_avmnist_state()
    # make the running statistics and biases non-trivial
    with torch.no_grad():
        for k, v in model.state_dict().items():
            if k.endswith("running_mean"):
                v.add_(0.05 * torch.arange(v.numel(), device=v.device, dtype=v.dtype) / v.numel())
                state[k] = v.detach().cpu().clone()
            if k.endswith("running_var"):
                v.mul_(1.5)
                state[k] = v.detach().cpu().clone()
    B = 16
    d = O.synthetic_batch(B, 5, (32, 94))
    model.eval()
    logits = model(d["audio"].to(DEV), d["image"].to(DEV)).cpu()
    ea = O.convblock_encoder_forward(state, "audio_encoder", d["audio"], False)
    ei = O.convblock_encoder_forward(state, "image_encoder", d["image"], False)
    ref = O.head_forward(state, ea, ei, None, 0.0)
    span = float(ref.abs().max())
    assert float((logits - ref).abs().max()) < 0.03 * span + 5e-3
    got_a = model.audio_encoder(d["audio"].to(DEV)).cpu()
    got_i = model.image_encoder(d["image"].to(DEV)).cpu()  # [B,1,28,28] as the dataset yields it
    assert float((got_a - ea).abs().max()) < 0.03 * float(ea.abs().max()) + 5e-3
    assert float((got_i - ei).abs().max()) < 0.03 * float(ei.abs().max()) + 5e-3
    out = model.validation_step(make_batch(d, B), LOSS, torch.device(DEV), None, return_test_info=True)
    want = float(F.cross_entropy(ref, d["labels"]))
    assert abs(out["loss"] - want) < 1e-2


def test_convblock_graph_replay_matches_eager_and_curve():
    """Eager steps, then CUDA-graph replays, give the same parameters as all-eager; the loss falls on a fixed batch."""
    B = 32
    d = O.synthetic_batch(B, 9, (32, 94))
    runs = []
    for graphs in (False, True):
        model = build(graphs=graphs)
        opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
        losses = [model.train_step(make_batch(d, B), opt, LOSS, torch.device(DEV), None, dropout_mask=d["dropout_mask"])["loss"] for _ in range(6)]
        runs.append((losses, model._engine.fs.P.detach().clone()))
    # same kernels, same order; only the fp64 atomics of the BatchNorm statistics may add in a different order
    assert np.allclose(runs[0][0], runs[1][0], atol=1e-4), (runs[0][0], runs[1][0])
    assert float((runs[0][1] - runs[1][1]).abs().max()) < 1e-4
    assert runs[0][0][-1] < runs[0][0][0]
