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


# ---------------------------------------------------------------------------------------------------------------------
# MonomodalEncoder around an MMIMDb encoder (configs/mmimdb/mono/*.yaml): BatchNorm1d -> Linear -> Linear(512, 23), bce_with_logits
# ---------------------------------------------------------------------------------------------------------------------
class BceTerm:
    def __init__(self):
        self.loss_fn, self.weight = torch.nn.BCEWithLogitsLoss(), 1.0


BCE = {"bce": BceTerm()}


def build_vec(in_dim, graphs=True):
    from mml_b200.mmimdb import MMIMDbModalityEncoder
    from mml_b200.mono import MonomodalEncoder

    torch.manual_seed(0)
    model = MonomodalEncoder(MMIMDbModalityEncoder(in_dim, 512), 512, 23).to(DEV)
    model._get_engine(torch.device(DEV)).use_graphs = graphs
    return model


def vec_batch(B, in_dim, seed):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(B, in_dim, generator=gen) * 1.5 + 0.3
    y = (torch.rand(B, 23, generator=gen) < 0.15).float()
    return x, y


@pytest.mark.parametrize("in_dim,B", [(300, 16), (4096, 128), (300, 128), (4096, 37)])
def test_monomodal_mmimdb_step_matches_oracle(in_dim, B):
    """Loss / logits / predictions vs the fp32 oracle; gradients vs the oracle rounded to bf16 where the kernels round (GEMM operands and
    output) at 2 %, vs the fp32 oracle at the bound that rounding itself causes; Adam update exact for the GPU's own gradients."""
    import gated_fusion_oracle as G

    model = build_vec(in_dim, graphs=False)
    torch.manual_seed(0)
    state = G.init_mono_vector_state(in_dim, 512, 23)
    sd = model.state_dict()
    assert list(sd.keys()) == list(state.keys())
    for k in state:
        assert torch.equal(sd[k].cpu(), state[k]), k
    x, y = vec_batch(B, in_dim, 31)
    opt = torch.optim.Adam(model.parameters(), lr=1e-5, weight_decay=1e-3)
    before = {k: v.detach().cpu().clone() for k, v in model.named_parameters()}
    out = model.train_step({"text": x, "label": y}, opt, BCE, torch.device(DEV), None)
    ref = G.mono_vector_train_step(copy.deepcopy(state), {}, x, y, apply_update=False)
    emu = G.mono_vector_train_step(copy.deepcopy(state), {}, x, y, apply_update=False, emulate_bf16=True)
    plan = next(iter(model._engine.plans.values()))
    assert abs(out["loss"] - ref["loss"]) < 2e-3 and "accuracy" not in out["metrics"]
    logits = plan.logits.cpu()
    span = float(ref["logits"].max() - ref["logits"].min())
    assert float((logits - ref["logits"]).abs().max()) < 1e-2 * span
    assert float((logits - emu["logits"]).abs().max()) < 2e-3 * span
    agree = float(((torch.sigmoid(logits) > 0.5) == ref["predictions"].bool()).float().mean())
    assert agree > 0.98
    g = torch.cat([p.grad.detach().cpu().float().reshape(-1) for _, p in model.named_parameters()])
    r_emu = torch.cat([emu["grads"][n].reshape(-1) for n, _ in model.named_parameters()])
    r_fp = torch.cat([ref["grads"][n].reshape(-1) for n, _ in model.named_parameters()])
    rel_emu, rel_fp, self_gap = float((g - r_emu).norm() / r_emu.norm()), float((g - r_fp).norm() / r_fp.norm()), float((r_emu - r_fp).norm() / r_fp.norm())
    print(f"in={in_dim} B={B}: grad rel L2 vs same-rounding oracle {rel_emu:.4f}, vs fp32 oracle {rel_fp:.4f} (oracle bf16-vs-fp32 {self_gap:.4f})")
    assert rel_emu < 2e-2 and rel_fp < max(3e-2, 2.5 * self_gap)
    # Adam (coupled weight decay, lr 1e-5): exact for the gradients the GPU produced; BN running statistics as torch updates them
    for n, p in model.named_parameters():
        gg = p.grad.detach().cpu().float() + 1e-3 * before[n]
        m, v = 0.1 * gg, 0.001 * gg * gg
        want = before[n] - 1e-5 * (m / 0.1) / ((v / 0.001).sqrt() + 1e-8)
        assert torch.allclose(p.detach().cpu(), want, rtol=1e-5, atol=1e-7), n
    sd = model.state_dict()
    assert int(sd["encoder.net.0.num_batches_tracked"]) == 1
    assert torch.allclose(sd["encoder.net.0.running_mean"].cpu(), 0.1 * x.mean(0), rtol=1e-4, atol=1e-5)
    assert torch.allclose(sd["encoder.net.0.running_var"].cpu(), 0.9 + 0.1 * x.var(0, unbiased=True), rtol=1e-4, atol=1e-5)


def test_monomodal_mmimdb_reference_fixture_curve_eval_and_handoff():
    import gated_fusion_oracle as G
    from mml_b200.mmimdb import MMIMDb, MMIMDbModalityEncoder, GatedBiModalNetwork, MLPGenreClassifier

    gld = np.load(os.path.join(GOLD, "mono_mmimdb_text_b16.npz"))
    batch, in_dim, seed, steps = (int(v) for v in gld["meta"])
    model = build_vec(in_dim)
    x, y = vec_batch(batch, in_dim, seed)
    opt = torch.optim.Adam(model.parameters(), lr=1e-5, weight_decay=1e-3)
    for step in range(steps):  # the recorded run of the reference class (eager, eager, graph replay)
        out = model.train_step({"text": x, "label": y}, opt, BCE, torch.device(DEV), None)
        assert abs(out["loss"] - float(gld["losses"][step])) < 2e-3, (step, out["loss"], float(gld["losses"][step]))
    # a longer run at a useful learning rate: the loss curve follows the oracle's, CUDA-graph replays included
    model = build_vec(in_dim)
    torch.manual_seed(0)
    state, opt_state = G.init_mono_vector_state(in_dim, 512, 23), {}
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
    x, y = vec_batch(64, in_dim, 5)
    got, want = [], []
    for step in range(30):
        got.append(model.train_step({"text_original": x, "text": x * 0, "genres": y}, opt, BCE, torch.device(DEV), None)["loss"])
        want.append(G.mono_vector_train_step(state, opt_state, x, y, lr=1e-3)["loss"])
    assert want[-1] < 0.5 * want[0]
    assert max(abs(a - b) for a, b in zip(got, want)) < 0.02 * want[0], (got[-3:], want[-3:])
    # eval mode (running statistics) and the recorder contract for multi-label predictions
    class Rec:
        class config:
            groups = ["classification"]
        calls = []
        def update_group(self, **kw):
            self.calls.append(kw)
    rec = Rec()
    ev = model.validation_step({"text": x, "label": y}, BCE, torch.device(DEV), rec)
    sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    rv = G.mono_vector_validation_step(sd_cpu, x, y)
    assert abs(ev["loss"] - rv["loss"]) < 5e-3
    assert rec.calls and rec.calls[0]["predictions"].dtype == torch.bool and rec.calls[0]["predictions"].shape == (64, 23)
    assert float((rec.calls[0]["predictions"] == rv["predictions"].bool()).float().mean()) > 0.97
    # stand-alone encoder call and the pretrain -> fusion hand-off (train_monomodal.py:790-801 saves get_encoder().state_dict(),
    # train_multimodal.py:186-187 loads it into the fusion model's encoder)
    model.eval()
    emb = model.get_encoder()(x.to(DEV))
    assert emb.shape == (64, 512)
    enc_sd = {k: v.detach().cpu().clone() for k, v in model.get_encoder().state_dict().items()}
    torch.manual_seed(1)
    fusion = MMIMDb(MMIMDbModalityEncoder(4096, 512), MMIMDbModalityEncoder(in_dim, 512), GatedBiModalNetwork(512, 512, 512, 512),
                    classifier=MLPGenreClassifier(512, 23, 512)).to(DEV)
    fusion.text_model.load_state_dict(enc_sd)
    fusion.eval()
    emb2 = fusion.text_model(x.to(DEV))
    assert torch.equal(emb.cpu(), emb2.cpu())


def test_monomodal_unsupported_requests_raise():
    from mml_b200.mono import MonomodalEncoder

    model = build_vec(300)
    x, y = vec_batch(8, 300, 1)
    opt = torch.optim.Adam(model.parameters(), lr=1e-5)
    with pytest.raises(ValueError):
        model.train_step({"text": x, "label": torch.zeros(8, dtype=torch.long)}, opt, BCE, torch.device(DEV), None)  # single-label targets
    with pytest.raises(NotImplementedError):
        model.train_step({"text": x, "label": y}, opt, LOSS, torch.device(DEV), None)  # cross-entropy on a multi-label head
    with pytest.raises(ValueError):
        model.train_step({"text": x[:, :299], "label": y}, opt, BCE, torch.device(DEV), None)
    with pytest.raises(NotImplementedError):
        MonomodalEncoder(torch.nn.LSTM(5, 64), 64, 3)


# ---------------------------------------------------------------------------------------------------------------------
# MonomodalEncoder around a MOSI encoder (configs/mosi/mono/*.yaml): LSTMEncoder / TextCNN -> Linear(64, 3), cross entropy
# ---------------------------------------------------------------------------------------------------------------------
def build_seq(kind, input_size, graphs=True):
    from mml_b200.mono import MonomodalEncoder
    from mml_b200.utt_fusion import LSTMEncoder, TextCNN

    torch.manual_seed(0)
    enc = LSTMEncoder(input_size, 64, "last") if kind == "lstm" else TextCNN(input_size, 64, 1, 128, [3, 4, 5], 0.5)
    model = MonomodalEncoder(enc, 64, 3).to(DEV)
    model._get_engine(torch.device(DEV)).use_graphs = graphs
    return model


@pytest.mark.parametrize("kind,input_size,B", [("lstm", 5, 8), ("lstm", 20, 32), ("textcnn", 768, 8), ("textcnn", 768, 32)])
def test_monomodal_mosi_step_matches_oracle(kind, input_size, B):
    """Loss / logits vs the fp32 oracle, gradients vs the oracle that rounds where the kernels round (TextCNN: text input, conv weights,
    conv output and its gradient in bf16; the LSTM path is fp32 throughout), Adam exact for the GPU's own gradients."""
    import utt_fusion_oracle as U
    from test_oracle_golden import mono_seq_batch

    model = build_seq(kind, input_size, graphs=False)
    torch.manual_seed(0)
    state = U.init_mono_seq_state(kind, input_size)
    sd = model.state_dict()
    assert list(sd.keys()) == list(state.keys()) and all(torch.equal(sd[k].cpu(), state[k]) for k in state)
    x, y, keep = mono_seq_batch(kind, B, input_size, 41)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
    before = {k: v.detach().cpu().clone() for k, v in model.named_parameters()}
    out = model.train_step({"audio": x, "label": y}, opt, LOSS, torch.device(DEV), None, dropout_mask=keep)
    ref = U.mono_seq_train_step(copy.deepcopy(state), {}, x, y, keep, apply_update=False)
    emu = U.mono_seq_train_step(copy.deepcopy(state), {}, x, y, keep, apply_update=False, emulate_bf16=True)
    plan = next(iter(model._engine.plans.values()))
    logits = plan.logits.cpu()
    span = float(ref["logits"].max() - ref["logits"].min()) + 1e-6
    tol = 1e-4 if kind == "lstm" else 2e-2
    assert abs(out["loss"] - ref["loss"]) < (1e-4 if kind == "lstm" else 1e-2)
    assert float((logits - ref["logits"]).abs().max()) < tol * span
    assert abs(out["metrics"]["accuracy"] - float((logits.argmax(1) == y).float().mean())) < 1e-6
    g = torch.cat([p.grad.detach().cpu().float().reshape(-1) for _, p in model.named_parameters()])
    r_emu = torch.cat([emu["grads"][n].reshape(-1) for n, _ in model.named_parameters()])
    r_fp = torch.cat([ref["grads"][n].reshape(-1) for n, _ in model.named_parameters()])
    rel_emu, rel_fp = float((g - r_emu).norm() / r_emu.norm()), float((g - r_fp).norm() / r_fp.norm())
    print(f"{kind} in={input_size} B={B}: grad rel L2 vs same-rounding oracle {rel_emu:.2e}, vs fp32 oracle {rel_fp:.2e}")
    assert rel_emu < (1e-4 if kind == "lstm" else 2e-3)
    assert rel_fp < (1e-4 if kind == "lstm" else 0.15)   # bf16 rounding moves max-over-time winners (same bound as tests/test_utt_gpu.py)
    for n, p in model.named_parameters():
        gg = p.grad.detach().cpu().float() + 1e-3 * before[n]
        m, v = 0.1 * gg, 0.001 * gg * gg
        want = before[n] - 1e-3 * (m / 0.1) / ((v / 0.001).sqrt() + 1e-8)
        assert torch.allclose(p.detach().cpu(), want, rtol=1e-4, atol=1e-6), n


@pytest.mark.parametrize("name,kind", [("mono_mosi_audio_b8", "lstm"), ("mono_mosi_text_b8", "textcnn")])
def test_monomodal_mosi_reference_fixture_curve_and_handoff(name, kind):
    import utt_fusion_oracle as U
    from test_oracle_golden import mono_seq_batch
    from mml_b200.utt_fusion import FcClassifier, LSTMEncoder, TextCNN, UttFusionModel

    gld = np.load(os.path.join(GOLD, name + ".npz"))
    batch, input_size, seed, steps, T = (int(v) for v in gld["meta"])
    model = build_seq(kind, input_size)
    x, y, keep = mono_seq_batch(kind, batch, input_size, seed, T)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
    for step in range(steps):  # the recorded run of the reference class (eager, eager, graph replay)
        out = model.train_step({"text": x, "label": y}, opt, LOSS, torch.device(DEV), None, dropout_mask=keep)
        assert abs(out["loss"] - float(gld["losses"][step])) < (1e-4 if kind == "lstm" else 2e-2), (step, out["loss"], float(gld["losses"][step]))
    # longer run without a given mask (the kernel's own dropout stream for TextCNN): the loss falls like the oracle's
    model = build_seq(kind, input_size)
    torch.manual_seed(0)
    state, opt_state = U.init_mono_seq_state(kind, input_size), {}
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
    x, y, _ = mono_seq_batch(kind, 32, input_size, 7, T)
    got, want = [], []
    for step in range(25):
        got.append(model.train_step({"video": x, "labels": y}, opt, LOSS, torch.device(DEV), None)["loss"])
        want.append(U.mono_seq_train_step(state, opt_state, x, y, None)["loss"])  # oracle without dropout
    assert got[-1] < (0.99 if kind == "lstm" else 0.9) * got[0]  # 25 Adam steps at lr 1e-3: the LSTM moves slowly (oracle: same curve)
    if kind == "lstm":
        assert max(abs(a - b) for a, b in zip(got, want)) < 2e-3, (got[-3:], want[-3:])
    ev = model.validation_step({"video": x, "labels": y}, LOSS, torch.device(DEV), None)
    rv = U.mono_seq_validation_step({k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, x, y)
    assert abs(ev["loss"] - rv["loss"]) < (1e-4 if kind == "lstm" else 2e-2)
    # hand-off: the pre-trained encoder's state_dict loads into the fusion model's encoder slot (train_multimodal.py:186-187)
    enc_sd = {k: v.detach().cpu().clone() for k, v in model.get_encoder().state_dict().items()}
    torch.manual_seed(3)
    fusion = UttFusionModel(LSTMEncoder(5 if kind != "lstm" else input_size, 64), LSTMEncoder(20, 64), TextCNN(768, 64, 1, 128, [3, 4, 5], 0.5),
                            FcClassifier(192, [192, 64, 32], 3, dropout=0.5), clip=1.0)
    (fusion.netA if kind == "lstm" else fusion.netT).load_state_dict(enc_sd)
    got_sd = (fusion.netA if kind == "lstm" else fusion.netT).state_dict()
    assert all(torch.equal(got_sd[k].cpu(), enc_sd[k]) for k in enc_sd)
