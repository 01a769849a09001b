"""GPU parity of the monomodal encoder pre-training step (mml_b200.mono.MonomodalEncoder; train_monomodal.py:64-260).

Same precision contract as the late-fusion step (tests/test_step_gpu.py): gradients are compared teacher-forced (the oracle's
backward over the activations the GPU stored), loss / logits un-forced, plus the committed fixture of the reference class.
"""
import copy
import os

import numpy as np
import pytest
import torch

import late_fusion_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class Term:
    def __init__(self):
        self.loss_fn, self.weight = torch.nn.CrossEntropyLoss(), 1.0


LOSS = {"cross_entropy": Term()}


def build(arch="resnet18", hidden=64, graphs=True):
    from mml_b200.mono import MonomodalEncoder
    from mml_b200.resnet import ResNet18, ResNet34

    torch.manual_seed(0)
    enc = (ResNet18 if arch == "resnet18" else ResNet34)(1, hidden)
    model = MonomodalEncoder(enc, hidden, 10).to(DEV)
    model._get_engine(torch.device(DEV)).use_graphs = graphs
    return model


def nchw(t):
    return t.detach().float().cpu().permute(0, 3, 1, 2).contiguous()


def forced_from_plan(ep, state):
    forced = {"encoder." + name: nchw(t) for name, t in ep.taps.items()}
    forced["encoder.avgpool"] = ep.pooled.detach().cpu().clone()
    act = torch.nn.functional.batch_norm(forced["encoder.conv1"], None, None, state["encoder.bn1.weight"], state["encoder.bn1.bias"], True, 0.1, 1e-5)
    forced["encoder.relu1"] = torch.relu(act)  # fp32: the fused stem tail pools the un-rounded BatchNorm outputs and rounds only the winner
    return forced


@pytest.mark.parametrize("arch,hidden,B,hw", [("resnet18", 64, 4, (32, 94)), ("resnet18", 64, 16, (112, 112)), ("resnet34", 128, 8, (28, 28))])
def test_monomodal_step_matches_oracle(arch, hidden, B, hw):
    model = build(arch, hidden, graphs=False)
    torch.manual_seed(0)
    state = O.init_monomodal_state(arch, 1, hidden, 10)
    sd = model.state_dict()
    assert list(sd.keys()) == list(state.keys())
    for k in state:
        assert torch.equal(sd[k].cpu(), state[k]), k
    seed = 11
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, *hw, generator=g)
    y = torch.randint(0, 10, (B,), generator=g)
    batch = {"AUDIO": x * 0.0, "AUDIO_original": x, "AUDIO_missing_index": torch.zeros(B), "labels": y, "pattern_name": ["a"] * B}
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
    out = model.train_step(batch, opt, LOSS, torch.device(DEV), None, config=None)
    plan = next(iter(model._engine.plans.values()))
    ref = O.monomodal_train_step(copy.deepcopy(state), {}, x, y, apply_update=False, forced=forced_from_plan(plan.enc, state))
    assert abs(out["loss"] - ref["loss"]) < 1e-4 and abs(out["metrics"]["accuracy"] - ref["accuracy"]) < 1e-6
    assert (plan.logits.cpu() - ref["logits"]).abs().max().item() < 1e-4
    gall = torch.cat([p.grad.detach().cpu().float().reshape(-1) for _, p in model.named_parameters()])
    rall = torch.cat([ref["grads"][n].reshape(-1) for n, _ in model.named_parameters()])
    glob = float((gall - rall).norm() / rall.norm())
    print(f"{arch} B={B}: forced-gradient rel L2 {glob:.4f}")
    assert glob < 3e-2, glob
    unf = O.monomodal_train_step(copy.deepcopy(state), {}, x, y, apply_update=False)
    assert abs(out["loss"] - unf["loss"]) < 2e-2 * max(1.0, unf["loss"])  # bf16 activations, batch statistics over as few as 8 x 1 x 1 values
    if (arch, B, hw) == ("resnet18", 4, (32, 94)):  # the reference's own MonomodalEncoder, oracle/make_golden.py
        gold = np.load(os.path.join(GOLD, "mono_resnet18_b4.npz"))
        assert abs(out["loss"] - float(gold["losses"][0])) < 2e-2
        rng = float(gold["logits"].max() - gold["logits"].min())
        assert np.abs(plan.logits.cpu().numpy() - gold["logits"]).max() < 0.1 * rng


def test_monomodal_training_eval_and_encoder_handoff():
    """100 steps through the CUDA graph follow the oracle's loss curve; the trained encoder's state_dict loads into the
    late-fusion model (train_monomodal.py:790-801 -> train_multimodal.py:186-187)."""
    B, steps = 32, 60
    model = build("resnet18", 64)
    torch.manual_seed(0)
    state = O.init_monomodal_state("resnet18", 1, 64, 10)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(B, 32, 94, generator=g)
    y = torch.randint(0, 10, (B,), generator=g)
    batch = {"audio": x, "labels": y}
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
    dev = torch.device(DEV)
    got = [model.train_step(batch, opt, LOSS, dev, None)["loss"] for _ in range(steps)]
    opt_state, ref = {}, []
    for _ in range(steps):
        ref.append(O.monomodal_train_step(state, opt_state, x, y)["loss"])
    assert abs(got[0] - ref[0]) < 2e-2 and got[-1] < 0.05 and ref[-1] < 0.05
    sm = lambda v: np.convolve(v, np.ones(5) / 5, mode="valid")
    assert np.abs(sm(got) - sm(ref)).max() < 0.15 * ref[0], (got[:8], ref[:8])
    ev = model.validation_step(batch, LOSS, dev, None)
    assert ev["metrics"]["accuracy"] == 1.0 and ev["loss"] < 0.2
    model.eval()
    logits = model(x.to(DEV))
    assert torch.equal(logits.argmax(1).cpu(), y)
    # hand the pre-trained encoder over to the fusion model
    from mml_b200.avmnist import AVMNIST
    from mml_b200.resnet import ResNet18, ResNet34

    enc_sd = {k: v.detach().cpu().clone() for k, v in model.get_encoder().state_dict().items()}
    fusion = AVMNIST(ResNet18(1, 64), ResNet34(1, 128), 128, dropout=0.0).to(DEV)
    fusion.audio_encoder.load_state_dict(enc_sd)
    emb_a = fusion.eval().audio_encoder(x.to(DEV))
    emb_b = model.get_encoder()(x.to(DEV))
    assert torch.allclose(emb_a, emb_b, rtol=1e-3, atol=1e-3)


def test_monomodal_unsupported_requests_raise():
    from mml_b200.mono import MonomodalEncoder

    with pytest.raises(NotImplementedError):
        MonomodalEncoder(torch.nn.Linear(4, 4), 4, 10)
    model = build()
    with pytest.raises(NotImplementedError):
        model.train_step({"audio": ["a.pt", "b.pt"], "labels": torch.zeros(2, dtype=torch.long)}, torch.optim.Adam(model.parameters()), LOSS,
                         torch.device(DEV), None)
    with pytest.raises(NotImplementedError):
        model.train_step({"audio": torch.rand(2, 28, 28), "labels": torch.zeros(2, 3)}, torch.optim.Adam(model.parameters()), LOSS, torch.device(DEV), None)
