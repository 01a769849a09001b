"""CPU oracle for the MML_Suite MOSI / UttFusion training step (BASELINE config 4).  TEST INFRASTRUCTURE ONLY.

Round 1 builds the ORACLE for this row (scope order: oracle first); the CUDA path for it is not built yet.  Same rules as the
other oracles: a functional fp32 restatement over a flat ``state`` dict with the reference's 24 ``state_dict()`` keys, pinned
by ``oracle/make_golden.py`` against the UNMODIFIED reference ``UttFusionModel`` (imported from /root/reference in the build
container) and re-checked against ``tests/golden/mosi_b8.npz`` by ``tests/test_oracle_golden.py``.

Reference files followed (paths relative to /root/reference/MML_Suite):
  models/msa/utt_fusion.py:106-198     UttFusionModel.forward / train_step (cat of three embeddings, CE on squeezed logits,
                                       clip_grad_norm_(clip), Adam)
  models/msa/networks/lstm.py:8-64     LSTMEncoder (nn.LSTM batch_first, "last" = h_T)
  models/msa/networks/textcnn.py:10-69 TextCNN (three Conv2d(1,128,(k,768)) -> ReLU -> max over time, cat, Dropout, Linear+ReLU)
  models/msa/networks/classifier.py:83-117  FcClassifier ([Linear, ReLU, Dropout] x 3, fc_out)
  data/mosi.py:62-70                   seven missing patterns atv/at/av/tv/a/t/v (x * m per modality, base_dataset.py:71)
  configs/mosi/centralised/utt_fusion_base_training.yaml:14-57   sizes 5/20/768 -> 64, classifier 192-192-64-32-3, clip 1.0,
                                       Adam lr 1e-3 wd 1e-3, batch 32
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from late_fusion_oracle import _linear_params, _q, _qg, _qw, adam_step, apply_missing_mask

Tensor = torch.Tensor
PATTERNS = {"atv": (1, 1, 1), "at": (1, 1, 0), "av": (1, 0, 1), "tv": (0, 1, 1), "a": (1, 0, 0), "t": (0, 1, 0), "v": (0, 0, 1)}  # (audio, text, video)
KERNEL_HEIGHTS = (3, 4, 5)


def _lstm_params(st: Dict[str, Tensor], prefix: str, input_size: int, hidden: int) -> None:
    """nn.LSTM(input, hidden): RNNBase.reset_parameters draws every parameter from U(-1/sqrt(hidden), 1/sqrt(hidden)) in
    registration order weight_ih, weight_hh, bias_ih, bias_hh."""
    stdv = 1.0 / math.sqrt(hidden)
    for name, shape in (("weight_ih_l0", (4 * hidden, input_size)), ("weight_hh_l0", (4 * hidden, hidden)), ("bias_ih_l0", (4 * hidden,)),
                        ("bias_hh_l0", (4 * hidden,))):
        st[f"{prefix}.rnn.{name}"] = torch.empty(*shape).uniform_(-stdv, stdv)


def _conv_params(out_c: int, in_c: int, kh: int, kw: int) -> Tuple[Tensor, Tensor]:
    w = torch.empty(out_c, in_c, kh, kw)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    bound = 1.0 / math.sqrt(in_c * kh * kw)
    return w, torch.empty(out_c).uniform_(-bound, bound)


def init_utt_state(audio_dim: int = 5, video_dim: int = 20, text_dim: int = 768, hidden: int = 64, channels: int = 128,
                   layers: Sequence[int] = (192, 64, 32), classes: int = 3) -> "OrderedDict[str, Tensor]":
    """UttFusionModel(LSTMEncoder(5,64), LSTMEncoder(20,64), TextCNN(768,64), FcClassifier(192,[192,64,32],3,dropout=.5)) built in
    YAML order netA, netV, netT, netC (utt_fusion_base_training.yaml:14-44)."""
    st: "OrderedDict[str, Tensor]" = OrderedDict()
    _lstm_params(st, "netA", audio_dim, hidden)
    _lstm_params(st, "netV", video_dim, hidden)
    for i, k in enumerate(KERNEL_HEIGHTS):
        st[f"netT.conv{i + 1}.weight"], st[f"netT.conv{i + 1}.bias"] = _conv_params(channels, 1, k, text_dim)
    st["netT.embd.0.weight"], st["netT.embd.0.bias"] = _linear_params(hidden, len(KERNEL_HEIGHTS) * channels)
    d = 3 * hidden
    for i, width in enumerate(layers):  # Sequential indices: Linear, ReLU, Dropout
        st[f"netC.module.{3 * i}.weight"], st[f"netC.module.{3 * i}.bias"] = _linear_params(width, d)
        d = width
    st["netC.fc_out.weight"], st["netC.fc_out.bias"] = _linear_params(classes, d)
    return st


def lstm_last(st: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    """h_T of a one-layer batch_first nn.LSTM started from zeros (gate order i, f, g, o)."""
    w_ih, w_hh = st[f"{prefix}.rnn.weight_ih_l0"], st[f"{prefix}.rnn.weight_hh_l0"]
    b = st[f"{prefix}.rnn.bias_ih_l0"] + st[f"{prefix}.rnn.bias_hh_l0"]
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h, c = x.new_zeros(B, H), x.new_zeros(B, H)
    xw = F.linear(x, w_ih)  # [B, T, 4H]: the input projection of all steps is one GEMM
    for t in range(T):
        gates = xw[:, t] + F.linear(h, w_hh) + b
        i, f, g, o = gates.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
    return h


def textcnn(st: Dict[str, Tensor], x: Tensor, keep: Optional[Tensor], p: float, emulate_bf16: bool = False, prefix: str = "netT") -> Tensor:
    """``emulate_bf16`` (debugging aid for the CUDA path, not part of the reference): round where the B200 path stores bf16 -- the
    text input, the convolution weights, the convolution output (bias is added afterwards in fp32) and its gradient."""
    B, T, D = x.shape
    q = emulate_bf16
    frame = _q(x, q).view(B, 1, T, D)
    outs = []
    for i in range(len(KERNEL_HEIGHTS)):
        conv = _qg(_q(F.conv2d(frame, _qw(st[f"{prefix}.conv{i + 1}.weight"], q), None), q), q) + st[f"{prefix}.conv{i + 1}.bias"].view(1, -1, 1, 1)
        outs.append(F.relu(conv.squeeze(3)).max(dim=2).values)
    allo = torch.cat(outs, 1)
    if keep is not None:
        allo = allo * keep / (1.0 - p)
    return F.relu(F.linear(allo, st[f"{prefix}.embd.0.weight"], st[f"{prefix}.embd.0.bias"]))


def utt_forward(st: Dict[str, Tensor], A: Tensor, V: Tensor, T: Tensor, keeps: Optional[Sequence[Tensor]] = None, p: float = 0.5,
                emulate_bf16: bool = False) -> Tensor:
    """logits [B, classes]; ``keeps`` = four {0,1} keep-masks (TextCNN dropout [B,384], classifier dropouts [B,192], [B,64], [B,32])
    in train mode, None in eval mode."""
    a, v = lstm_last(st, "netA", A), lstm_last(st, "netV", V)
    t = textcnn(st, T, keeps[0] if keeps is not None else None, p, emulate_bf16)
    x = torch.cat([a, v, t], dim=-1)  # utt_fusion.py:147
    i = 0
    while f"netC.module.{3 * i}.weight" in st:
        x = F.relu(F.linear(x, st[f"netC.module.{3 * i}.weight"], st[f"netC.module.{3 * i}.bias"]))
        if keeps is not None:
            x = x * keeps[1 + i] / (1.0 - p)
        i += 1
    return F.linear(x, st["netC.fc_out.weight"], st["netC.fc_out.bias"])


def train_step(st: "OrderedDict[str, Tensor]", opt_state: Dict, A: Tensor, V: Tensor, T: Tensor, labels: Tensor, keeps: Optional[Sequence[Tensor]],
               lr: float = 1e-3, weight_decay: float = 1e-3, clip: Optional[float] = 1.0, apply_update: bool = True,
               emulate_bf16: bool = False) -> Dict[str, object]:
    """utt_fusion.py:151-198: forward (train), CE on squeezed logits / labels, backward, clip_grad_norm_(clip), Adam.step."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in st.items()}
    logits = utt_forward(leaves, A, V, T, keeps, emulate_bf16=emulate_bf16)
    loss = F.cross_entropy(logits.squeeze(), labels.squeeze()) * 1.0
    gl = torch.autograd.grad(loss, list(leaves.values()))
    grads = dict(zip(leaves.keys(), gl))
    total_norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    if clip is not None:  # torch.nn.utils.clip_grad_norm_: coef = clip / (norm + 1e-6), clamped to 1
        coef = torch.clamp(clip / (total_norm + 1e-6), max=1.0)
        grads = {k: g * coef for k, g in grads.items()}
    if apply_update:
        with torch.no_grad():
            adam_step(st, grads, opt_state, lr=lr, weight_decay=weight_decay)
    preds = F.softmax(logits.detach(), dim=-1).argmax(dim=-1)
    return {"loss": float(loss.item()), "logits": logits.detach(), "predictions": preds, "grads": grads, "grad_norm": float(total_norm)}


@torch.no_grad()
def validation_step(st: Dict[str, Tensor], A: Tensor, V: Tensor, T: Tensor, labels: Tensor) -> Dict[str, object]:
    logits = utt_forward(st, A, V, T, None)
    return {"loss": float(F.cross_entropy(logits.squeeze(), labels.squeeze()).item()), "logits": logits, "predictions": logits.argmax(-1)}


def synthetic_batch(batch: int, seed: int, seq_len: int = 50, audio_dim: int = 5, video_dim: int = 20, text_dim: int = 768,
                    layers: Sequence[int] = (192, 64, 32), channels: int = 128, classes: int = 3) -> Dict[str, object]:
    """Aligned MOSI-like sequences [B, T, F] with per-sample valid lengths (zero padded), one of the seven missing patterns per
    sample, labels in {0,1,2} and the four dropout keep-masks."""
    g = torch.Generator().manual_seed(seed)
    A, V, T = (torch.randn(batch, seq_len, d, generator=g) for d in (audio_dim, video_dim, text_dim))
    lengths = torch.randint(seq_len // 3, seq_len + 1, (batch,), generator=g)
    valid = (torch.arange(seq_len)[None, :] < lengths[:, None]).float()[:, :, None]
    A, V, T = A * valid, V * valid, T * valid
    names = list(PATTERNS)
    pat = [names[int(i)] for i in torch.randint(0, len(names), (batch,), generator=g)]
    ma, mt, mv = (torch.tensor([float(PATTERNS[p][j]) for p in pat]) for j in range(3))
    y = torch.randint(0, classes, (batch,), generator=g)
    keeps = [(torch.rand(batch, n, generator=g) >= 0.5).float() for n in (len(KERNEL_HEIGHTS) * channels, *layers)]
    return {"audio": A, "video": V, "text": T, "lengths": lengths, "pattern_name": pat, "audio_mask": ma, "text_mask": mt, "video_mask": mv,
            "labels": y, "keeps": keeps, "audio_masked": apply_missing_mask(A, ma), "video_masked": apply_missing_mask(V, mv),
            "text_masked": apply_missing_mask(T, mt)}


# ----------------------------------------------------------------------------------------------
# monomodal pre-training of ONE MOSI encoder (train_monomodal.py:64-92,224-260 with configs/mosi/mono/*.yaml):
#   MonomodalEncoder(LSTMEncoder(5 | 20, 64, "last"), 64, 3)  or  MonomodalEncoder(TextCNN(768, 64, dropout .5), 64, 3),
#   cross_entropy, Adam(lr 1e-3, weight_decay 1e-3), no gradient clip.  CUDA path: mml_b200/mono.py (_SeqMonoPlan).
# ----------------------------------------------------------------------------------------------
def init_mono_seq_state(kind: str, input_size: int, hidden: int = 64, classes: int = 3, channels: int = 128) -> "OrderedDict[str, Tensor]":
    """Same RNG draws, in the same order, as ``MonomodalEncoder(encoder, hidden, classes)`` with encoder = LSTMEncoder(input_size, hidden)
    (kind "lstm") or TextCNN(input_size, hidden, out_channels=channels) (kind "textcnn")."""
    st: "OrderedDict[str, Tensor]" = OrderedDict()
    if kind == "lstm":
        _lstm_params(st, "encoder", input_size, hidden)
    elif kind == "textcnn":
        for i, k in enumerate(KERNEL_HEIGHTS):
            st[f"encoder.conv{i + 1}.weight"], st[f"encoder.conv{i + 1}.bias"] = _conv_params(channels, 1, k, input_size)
        st["encoder.embd.0.weight"], st["encoder.embd.0.bias"] = _linear_params(hidden, len(KERNEL_HEIGHTS) * channels)
    else:
        raise ValueError(kind)
    st["classifier.weight"], st["classifier.bias"] = _linear_params(classes, hidden)
    return st


def mono_seq_forward(st: Dict[str, Tensor], x: Tensor, keep: Optional[Tensor] = None, p: float = 0.5, emulate_bf16: bool = False) -> Tensor:
    if "encoder.rnn.weight_ih_l0" in st:
        emb = lstm_last(st, "encoder", x)
    else:
        emb = textcnn(st, x, keep, p, emulate_bf16, prefix="encoder")
    return F.linear(emb.reshape(emb.shape[0], -1), st["classifier.weight"], st["classifier.bias"])


def mono_seq_train_step(st: "OrderedDict[str, Tensor]", opt_state: Dict, x: Tensor, labels: Tensor, keep: Optional[Tensor] = None, p: float = 0.5,
                        lr: float = 1e-3, weight_decay: float = 1e-3, apply_update: bool = True, emulate_bf16: bool = False) -> Dict[str, object]:
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in st.items()}
    logits = mono_seq_forward(leaves, x, keep, p, emulate_bf16)
    loss = F.cross_entropy(logits, labels) * 1.0
    gl = torch.autograd.grad(loss, list(leaves.values()))
    grads = dict(zip(leaves.keys(), gl))
    if apply_update:
        with torch.no_grad():
            adam_step(st, grads, opt_state, lr=lr, weight_decay=weight_decay)
    preds = torch.argmax(logits.detach(), dim=1)  # train_monomodal.py:239
    return {"loss": float(loss.item()), "logits": logits.detach(), "predictions": preds, "grads": grads,
            "accuracy": float((preds == labels).float().mean())}


@torch.no_grad()
def mono_seq_validation_step(st: Dict[str, Tensor], x: Tensor, labels: Tensor) -> Dict[str, object]:
    logits = mono_seq_forward(st, x, None)
    return {"loss": float(F.cross_entropy(logits, labels).item()), "logits": logits, "predictions": logits.argmax(-1)}
