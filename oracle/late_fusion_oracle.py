"""CPU oracle for the MML_Suite late-fusion training step.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (plain PyTorch, fp32, CPU) of the reference's
late-fusion hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only
as the checker or as the timed CPU baseline -- never from the product package
(``task-specific-pretraining-multimodal_b200/``), which must fail loudly when its CUDA
library is missing instead of falling back to this code.

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so this oracle is pinned against outputs of the reference
itself: ``oracle/make_golden.py`` imports the unmodified reference modules from
``/root/reference`` (build container only), runs them on seeded inputs and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this restatement
against those vectors on every CPU test run.

The restatement is *functional*: the model is a flat ``state`` dict carrying
exactly the reference's ``state_dict()`` keys (346 entries, fp32 OIHW conv
weights, int64 ``num_batches_tracked``), so it can be compared entry by entry
with the reference and with the CUDA path.

Reference files followed (paths relative to /root/reference/MML_Suite):
  models/msa/networks/resnet.py:8-54     BasicBlock
  models/msa/networks/resnet.py:113-239  ResNetEncoder, ResNet18, ResNet34
  models/avmnist.py:193-310              AVMNIST ctor / forward / train_step
  experiment_utils/loss.py:37-148        LossFunctionGroup -> CrossEntropyLoss()
  data/base_dataset.py:61-74             mask application  x * m
  config/data_config.py:22-106           pattern -> P(present)
  (torch defaults for BatchNorm2d / Adam / CrossEntropyLoss / Dropout)
"""
from __future__ import annotations

import math
from collections import OrderedDict
from itertools import chain, combinations
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = "OrderedDict[str, Tensor]"

RESNET_LAYERS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3)}
STAGE_PLANES = (64, 128, 256, 512)
NUM_CLASSES = 10
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------------------------
# a1 -- missing-modality mask (data/base_dataset.py:61-74, config/data_config.py:58-106)
# ----------------------------------------------------------------------------------------------
def apply_missing_mask(original: Tensor, mask) -> Tensor:
    """``sample[mod] = original * mask`` (base_dataset.py:71).

    A true IEEE multiply by 0.0 / 1.0: x*1 == x bit for bit, x*0 == +-0 with the
    sign of x, NaN/Inf * 0 == NaN.  ``mask`` is a per-sample scalar (or a [B]
    vector broadcast over the trailing dims for a batch).
    """
    m = torch.as_tensor(mask, dtype=original.dtype)
    if m.dim() == 1 and original.dim() > 1:
        m = m.view(-1, *([1] * (original.dim() - 1)))
    return original * m


def reverse_missing_mask(original: Tensor, mask) -> Tensor:
    """``original * -1 * (mask - 1)`` (base_dataset.py:72)."""
    m = torch.as_tensor(mask, dtype=original.dtype)
    if m.dim() == 1 and original.dim() > 1:
        m = m.view(-1, *([1] * (original.dim() - 1)))
    return original * -1 * (m - 1)


def generate_patterns(
    modalities: "OrderedDict[str, Tuple[float, Optional[Sequence[str]]]]",
    selected_patterns: Optional[Sequence[str]] = None,
) -> Dict[str, Dict[str, float]]:
    """Pattern name -> {modality: P(present)} (config/data_config.py:58-106).

    ``modalities`` maps modality name -> (missing_rate, apply_to).  Every non-empty
    subset gets P=1 for its members (or 1-rate when the pattern is listed in that
    modality's ``apply_to``) and P=0 for non-members; the full pattern is then
    overwritten with round(1-rate, 4) for every modality; finally filtered by
    ``selected_patterns`` (each sorted by character, data_config.py:50-53).
    """
    names = list(modalities.keys())
    combos = list(chain.from_iterable(combinations(names, r) for r in range(1, len(names) + 1)))
    combos = sorted(combos, key=lambda c: (len(c), c))
    full = "".join(m[0] for m in sorted(combos[-1]))
    patterns: Dict[str, Dict[str, float]] = {}
    for combo in combos:
        pname = "".join(m[0] for m in sorted(combo))
        probs = {}
        for m in names:
            rate, apply_to = modalities[m]
            if m in combo:
                probs[m] = round(1.0 - rate, 4) if (apply_to is not None and pname in apply_to) else 1.0
            else:
                probs[m] = 0.0
        patterns[pname] = probs
    patterns[full] = {m: round(1.0 - modalities[m][0], 4) for m in names}
    if selected_patterns:
        sel = ["".join(sorted(p)) for p in selected_patterns]
        patterns = {k: v for k, v in patterns.items() if k in sel}
    return patterns


# ----------------------------------------------------------------------------------------------
# model state construction (resnet.py:113-197, avmnist.py:193-236) -- same RNG consumption
# order as the reference constructors, so the same torch seed gives the same weights.
# ----------------------------------------------------------------------------------------------
def _conv_weight(out_c: int, in_c: int, k: int) -> Tensor:
    w = torch.empty(out_c, in_c, k, k)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))  # nn.Conv2d.reset_parameters
    return w


def _linear_params(out_f: int, in_f: int) -> Tuple[Tensor, Tensor]:
    w = torch.empty(out_f, in_f)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))  # nn.Linear.reset_parameters
    bound = 1.0 / math.sqrt(in_f)
    b = torch.empty(out_f).uniform_(-bound, bound)
    return w, b


def _bn_entries(state: "OrderedDict[str, Tensor]", prefix: str, c: int) -> None:
    state[prefix + ".weight"] = torch.ones(c)
    state[prefix + ".bias"] = torch.zeros(c)
    state[prefix + ".running_mean"] = torch.zeros(c)
    state[prefix + ".running_var"] = torch.ones(c)
    state[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def block_plan(arch: str) -> List[Tuple[str, int, int, int, bool]]:
    """[(block prefix, inplanes, planes, stride, has_downsample)] in module order."""
    plan = []
    inplanes = 64
    for li, (planes, nblocks) in enumerate(zip(STAGE_PLANES, RESNET_LAYERS[arch])):
        for bi in range(nblocks):
            stride = 2 if (bi == 0 and li > 0) else 1
            ds = bi == 0 and (stride != 1 or inplanes != planes)
            plan.append((f"layer{li + 1}.{bi}", inplanes, planes, stride, ds))
            inplanes = planes
    return plan


def init_resnet_state(prefix: str, arch: str, in_channels: int, hidden_dim: int) -> "OrderedDict[str, Tensor]":
    """state entries of ResNetEncoder(BasicBlock, layers, in_channels, hidden_dim) (resnet.py:118-169).

    Construction order == registration order == RNG order of the reference: every
    nn.Conv2d / nn.Linear draws its default init when constructed, then all conv
    weights are re-drawn with kaiming_normal_(fan_out, relu) in ``modules()`` order
    (resnet.py:153-158).
    """
    st: "OrderedDict[str, Tensor]" = OrderedDict()
    conv_keys: List[str] = []

    def conv(key: str, o: int, i: int, k: int) -> None:
        st[key] = _conv_weight(o, i, k)
        conv_keys.append(key)

    conv(f"{prefix}conv1.weight", 64, in_channels, 7)
    _bn_entries(st, f"{prefix}bn1", 64)
    for bp, inpl, planes, stride, ds in block_plan(arch):
        # _make_layer builds the downsample Sequential BEFORE the block (resnet.py:176-183),
        # so its conv draws first; but registration (state_dict / modules()) order puts
        # conv1, bn1, conv2, bn2 before downsample.
        ds_w = _conv_weight(planes, inpl, 1) if ds else None
        conv(f"{prefix}{bp}.conv1.weight", planes, inpl, 3)
        _bn_entries(st, f"{prefix}{bp}.bn1", planes)
        conv(f"{prefix}{bp}.conv2.weight", planes, planes, 3)
        _bn_entries(st, f"{prefix}{bp}.bn2", planes)
        if ds:
            st[f"{prefix}{bp}.downsample.0.weight"] = ds_w
            conv_keys.append(f"{prefix}{bp}.downsample.0.weight")
            _bn_entries(st, f"{prefix}{bp}.downsample.1", planes)
    w, b = _linear_params(hidden_dim, 512)
    st[f"{prefix}fc.weight"], st[f"{prefix}fc.bias"] = w, b
    for key in conv_keys:  # modules() order == registration order
        torch.nn.init.kaiming_normal_(st[key], mode="fan_out", nonlinearity="relu")
    return st


def init_avmnist_state(
    audio_arch: str = "resnet18",
    image_arch: str = "resnet34",
    audio_hidden: int = 64,
    image_hidden: int = 128,
    hidden_dim: int = 128,
) -> "OrderedDict[str, Tensor]":
    """AVMNIST(ResNet18(1,audio_hidden), ResNet34(1,image_hidden), hidden_dim) state (avmnist.py:193-236).

    Encoders are constructed by the caller before the AVMNIST ctor (YAML tag order:
    audio then image), then the three head Linears.
    """
    st: "OrderedDict[str, Tensor]" = OrderedDict()
    st.update(init_resnet_state("audio_encoder.", audio_arch, 1, audio_hidden))
    st.update(init_resnet_state("image_encoder.", image_arch, 1, image_hidden))
    for idx, (o, i) in zip((0, 3, 5), ((hidden_dim, audio_hidden + image_hidden), (hidden_dim // 2, hidden_dim), (NUM_CLASSES, hidden_dim // 2))):
        w, b = _linear_params(o, i)
        st[f"net.{idx}.weight"], st[f"net.{idx}.bias"] = w, b
    return st


def is_parameter(key: str) -> bool:
    return not key.endswith(("running_mean", "running_var", "num_batches_tracked"))


def arch_of(state: Dict[str, Tensor], prefix: str) -> str:
    return "resnet34" if f"{prefix}layer1.2.conv1.weight" in state else "resnet18"


# ----------------------------------------------------------------------------------------------
# optional bf16 storage emulation (a DEBUGGING AID for the CUDA path, not part of the reference):
# the B200 path stores activations and activation-gradients as bf16 and keeps fp32 accumulation.
# With emulate_bf16=True the oracle rounds at the same points, so that schedule / wiring bugs can
# be told apart from precision noise.
# ----------------------------------------------------------------------------------------------
class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


class _RoundGradBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


def _q(x: Tensor, on: bool) -> Tensor:
    return _RoundBF16.apply(x) if on else x


def _qg(x: Tensor, on: bool) -> Tensor:
    return _RoundGradBF16.apply(x) if on else x


def _qw(w: Tensor, on: bool) -> Tensor:
    """bf16 operand copy of an fp32 master weight (gradient stays fp32)."""
    return w + (w.detach().to(torch.bfloat16).float() - w.detach()) if on else w


# ----------------------------------------------------------------------------------------------
# a2-a5 -- encoders (resnet.py:37-54, 199-219)
# ----------------------------------------------------------------------------------------------
def _batch_norm(state: Dict[str, Tensor], prefix: str, x: Tensor, training: bool, update_running: bool) -> Tensor:
    rm, rv = state[prefix + ".running_mean"], state[prefix + ".running_var"]
    if training and not update_running:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(x, rm, rv, state[prefix + ".weight"], state[prefix + ".bias"], training, BN_MOMENTUM, BN_EPS)
    if training and update_running:
        state[prefix + ".num_batches_tracked"] += 1
    return y


def resnet_forward(
    state: Dict[str, Tensor],
    prefix: str,
    x: Tensor,
    training: bool,
    update_running: bool = True,
    taps: Optional[Dict[str, Tensor]] = None,
    emulate_bf16: bool = False,
    forced: Optional[Dict[str, Tensor]] = None,
) -> Tensor:
    """ResNetEncoder.forward (resnet.py:199-219).

    ``taps`` collects every stored intermediate (NCHW fp32) under ``prefix + name``.
    ``forced`` (testing aid) substitutes the VALUE of those intermediates by the given tensors while
    keeping the autograd graph (x + (forced - x).detach()): the backward pass then runs over exactly the
    activations another implementation stored, which makes a gradient comparison well conditioned.
    """
    eb = emulate_bf16
    if x.dim() == 3:
        x = x.unsqueeze(1)  # resnet.py:201-203

    def bn(p: str, t: Tensor) -> Tensor:
        return _batch_norm(state, prefix + p, t, training, update_running)

    def pt(name: str, t: Tensor) -> Tensor:
        if forced is not None and (prefix + name) in forced:
            t = t + (forced[prefix + name].to(t.dtype) - t).detach()
        if taps is not None:
            taps[prefix + name] = t
        return t

    x = pt("conv1", _q(F.conv2d(x, state[prefix + "conv1.weight"], None, stride=2, padding=3), eb))
    x = pt("relu1", _q(F.relu(bn("bn1", x)), eb))
    x = pt("maxpool", F.max_pool2d(x, kernel_size=3, stride=2, padding=1))
    for bp, _inpl, _planes, stride, ds in block_plan(arch_of(state, prefix)):
        identity = _qg(x, eb)
        out = pt(f"{bp}.conv1", _q(F.conv2d(_qg(x, eb), _qw(state[f"{prefix}{bp}.conv1.weight"], eb), None, stride=stride, padding=1), eb))
        out = pt(f"{bp}.relu1", _q(F.relu(bn(f"{bp}.bn1", out)), eb))
        out = pt(f"{bp}.conv2", _q(F.conv2d(out, _qw(state[f"{prefix}{bp}.conv2.weight"], eb), None, stride=1, padding=1), eb))
        out = bn(f"{bp}.bn2", out)
        if ds:
            identity = pt(f"{bp}.downsample", _q(F.conv2d(_qg(x, eb), _qw(state[f"{prefix}{bp}.downsample.0.weight"], eb), None, stride=stride), eb))
            identity = bn(f"{bp}.downsample.1", identity)
        x = pt(bp, _q(F.relu(out + identity), eb))
    x = pt("avgpool", torch.flatten(F.adaptive_avg_pool2d(x, (1, 1)), 1))
    return F.linear(x, state[prefix + "fc.weight"], state[prefix + "fc.bias"])


# ----------------------------------------------------------------------------------------------
# a6-a7 -- concat fusion head + loss (avmnist.py:219-267, loss.py:98-148)
# ----------------------------------------------------------------------------------------------
def head_forward(state: Dict[str, Tensor], audio: Tensor, image: Tensor, dropout_mask: Optional[Tensor], dropout_p: float) -> Tensor:
    """``net(cat(audio, image))``; dropout applied as ``h * mask / (1-p)`` with a GIVEN 0/1 mask
    (torch's Philox stream is not reproduced; parity runs pass the mask in, or p=0 / eval)."""
    fused = torch.cat((audio, image), dim=1)
    h = F.relu(F.linear(fused, state["net.0.weight"], state["net.0.bias"]))
    if dropout_mask is not None and dropout_p > 0:
        h = h * dropout_mask / (1.0 - dropout_p)
    h = F.relu(F.linear(h, state["net.3.weight"], state["net.3.bias"]))
    return F.linear(h, state["net.5.weight"], state["net.5.bias"])


def late_fusion_forward(
    state: Dict[str, Tensor],
    A: Tensor,
    I: Tensor,
    training: bool,
    dropout_mask: Optional[Tensor] = None,
    dropout_p: float = 0.5,
    update_running: bool = True,
    taps: Optional[Dict[str, Tensor]] = None,
    emulate_bf16: bool = False,
    forced: Optional[Dict[str, Tensor]] = None,
) -> Tensor:
    """AVMNIST.forward(A=A, I=I) (avmnist.py:238-267); inputs are already masked (x * m)."""
    audio = resnet_forward(state, "audio_encoder.", A.float(), training, update_running, taps, emulate_bf16, forced)
    image = resnet_forward(state, "image_encoder.", I.float(), training, update_running, taps, emulate_bf16, forced)
    if taps is not None:
        taps["audio_emb"], taps["image_emb"] = audio, image
    return head_forward(state, audio, image, dropout_mask if training else None, dropout_p)


def total_loss(logits: Tensor, labels: Tensor) -> Tensor:
    """LossFunctionGroup({"cross_entropy": CrossEntropyLoss() x 1.0})(...)["total_loss"].

    YAML ``loss_args`` is never read (loss.py:91 looks up ``loss_kwargs``) so CE runs with defaults:
    mean reduction, no smoothing, no class weights."""
    return F.cross_entropy(logits, labels) * 1.0


# ----------------------------------------------------------------------------------------------
# a9 -- torch.optim.Adam(lr, weight_decay) restated (coupled L2, bias-corrected; torch defaults)
# ----------------------------------------------------------------------------------------------
def adam_step(
    params: Dict[str, Tensor],
    grads: Dict[str, Tensor],
    opt_state: Dict[str, Dict[str, Tensor]],
    lr: float = 5e-4,
    weight_decay: float = 1e-4,
    betas: Tuple[float, float] = (0.9, 0.999),
    eps: float = 1e-8,
) -> None:
    b1, b2 = betas
    for k, p in params.items():
        g = grads[k]
        s = opt_state.setdefault(k, {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)})
        s["step"] += 1
        t = s["step"]
        if weight_decay != 0:
            g = g + weight_decay * p
        s["exp_avg"].mul_(b1).add_(g, alpha=1 - b1)
        s["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** t
        bc2 = 1 - b2 ** t
        denom = (s["exp_avg_sq"].sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(s["exp_avg"], denom, value=-(lr / bc1))


# ----------------------------------------------------------------------------------------------
# a8 -- one training step (avmnist.py:269-310 body: zero_grad, forward, loss, backward, step)
# ----------------------------------------------------------------------------------------------
def train_step(
    state: "OrderedDict[str, Tensor]",
    opt_state: Dict[str, Dict[str, Tensor]],
    A: Tensor,
    I: Tensor,
    labels: Tensor,
    dropout_mask: Optional[Tensor] = None,
    dropout_p: float = 0.5,
    lr: float = 5e-4,
    weight_decay: float = 1e-4,
    apply_update: bool = True,
    grad_scale: float = 1.0,
    emulate_bf16: bool = False,
    forced: Optional[Dict[str, Tensor]] = None,
) -> Dict[str, object]:
    """Returns {"loss", "logits", "predictions", "grads"}; mutates ``state`` / ``opt_state`` in place."""
    params = {k: v for k, v in state.items() if is_parameter(k)}
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    work = dict(state)
    work.update(leaves)
    logits = late_fusion_forward(work, A, I, True, dropout_mask, dropout_p, emulate_bf16=emulate_bf16, forced=forced)
    loss = total_loss(logits, labels)
    gl = torch.autograd.grad(loss, list(leaves.values()))
    grads = {k: g * grad_scale for k, g in zip(leaves.keys(), gl)}
    for k in state:  # running stats were updated in ``work`` (same tensor objects except num_batches_tracked)
        if k.endswith("num_batches_tracked"):
            state[k] = work[k]
    if apply_update:
        with torch.no_grad():
            adam_step(params, grads, opt_state, lr=lr, weight_decay=weight_decay)
    preds = torch.softmax(logits.detach(), dim=1).argmax(dim=1)  # avmnist.py:305
    return {"loss": float(loss.item()), "logits": logits.detach(), "predictions": preds, "grads": grads}


@torch.no_grad()
def validation_step(state: Dict[str, Tensor], A: Tensor, I: Tensor, labels: Tensor) -> Dict[str, object]:
    """avmnist.py:312-360: eval-mode forward (running stats, no dropout), CE, argmax."""
    logits = late_fusion_forward(state, A, I, False)
    loss = total_loss(logits, labels)
    return {"loss": float(loss.item()), "logits": logits, "predictions": torch.softmax(logits, 1).argmax(1)}


# ----------------------------------------------------------------------------------------------
# 8f rank 3 -- monomodal encoder pre-training (train_monomodal.py:64-92, 224-232): encoder -> Linear -> CE -> Adam
# ----------------------------------------------------------------------------------------------
def init_monomodal_state(arch: str = "resnet18", in_channels: int = 1, hidden_dim: int = 64, num_classes: int = NUM_CLASSES) -> "OrderedDict[str, Tensor]":
    """MonomodalEncoder(ResNetXX(in_channels, hidden_dim), output_dim=hidden_dim, num_classes): the encoder is built by the YAML
    loader first, the classifier Linear in the wrapper's constructor (train_monomodal.py:68-71)."""
    st = init_resnet_state("encoder.", arch, in_channels, hidden_dim)
    st["classifier.weight"], st["classifier.bias"] = _linear_params(num_classes, hidden_dim)
    return st


def monomodal_forward(state: Dict[str, Tensor], x: Tensor, training: bool, emulate_bf16: bool = False,
                      forced: Optional[Dict[str, Tensor]] = None, taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    emb = resnet_forward(state, "encoder.", x, training, taps=taps, emulate_bf16=emulate_bf16, forced=forced)
    return F.linear(emb.reshape(emb.shape[0], -1), state["classifier.weight"], state["classifier.bias"])


def monomodal_train_step(state: "OrderedDict[str, Tensor]", opt_state: Dict, x: Tensor, labels: Tensor, lr: float = 5e-4,
                         weight_decay: float = 1e-4, apply_update: bool = True, emulate_bf16: bool = False,
                         forced: Optional[Dict[str, Tensor]] = None) -> Dict[str, object]:
    params = {k: v for k, v in state.items() if is_parameter(k)}
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    work = dict(state)
    work.update(leaves)
    logits = monomodal_forward(work, x, True, emulate_bf16, forced)
    loss = total_loss(logits, labels)
    gl = torch.autograd.grad(loss, list(leaves.values()))
    grads = dict(zip(leaves.keys(), gl))
    for k in state:
        if k.endswith("num_batches_tracked"):
            state[k] = work[k]
    if apply_update:
        with torch.no_grad():
            adam_step(params, grads, opt_state, lr=lr, weight_decay=weight_decay)
    preds = torch.argmax(logits.detach(), dim=1)  # train_monomodal.py:239
    return {"loss": float(loss.item()), "logits": logits.detach(), "predictions": preds, "grads": grads,
            "accuracy": float((preds == labels).float().mean())}


# ----------------------------------------------------------------------------------------------
# 8f rank 4 -- AVMNIST with the ConvBlock encoders (models/avmnist.py:34-185, models/conv.py:16-59; configs/avmnist/centralised/
# train_avmnist.yaml).  CUDA path: mml_b200/convblock.py, parity in tests/test_convblock_gpu.py.
# ----------------------------------------------------------------------------------------------
CONVBLOCK_CHANNELS = {"audio_encoder": ((1, 32), (32, 32), (32, 64), (64, 64)), "image_encoder": ((1, 32), (32, 64), (64, 64), (64, 64))}
CONVBLOCK_POOLS = {"audio_encoder": (2, 3), "image_encoder": (2, 2)}   # MaxPool2d kernel sizes (stride = kernel)
CONVBLOCK_FLAT = {"audio_encoder": 4800, "image_encoder": 3136}


def _conv_params_bias(out_c: int, in_c: int, k: int) -> Tuple[Tensor, Tensor]:
    w = _conv_weight(out_c, in_c, k)  # nn.Conv2d.reset_parameters: weight, then bias ~ U(+-1/sqrt(fan_in))
    bound = 1.0 / math.sqrt(in_c * k * k)
    return w, torch.empty(out_c).uniform_(-bound, bound)


def init_convblock_avmnist_state(audio_hidden: int = 64, image_hidden: int = 128, hidden_dim: int = 128) -> "OrderedDict[str, Tensor]":
    """AVMNIST(MNISTAudio(...), MNISTImage(...), hidden_dim) built in YAML order; inside a ConvBlock the registration (and RNG)
    order is conv_one, conv_two, batch_norm_one, batch_norm_two (conv.py:24-45)."""
    st: "OrderedDict[str, Tensor]" = OrderedDict()
    for enc, hid in (("audio_encoder", audio_hidden), ("image_encoder", image_hidden)):
        ch = CONVBLOCK_CHANNELS[enc]
        for bi, slot in enumerate((0, 2)):  # Sequential: ConvBlock, MaxPool2d, ConvBlock, MaxPool2d, Flatten, Linear
            (i1, o1), (i2, o2) = ch[2 * bi], ch[2 * bi + 1]
            p = f"{enc}.net.{slot}"
            st[p + ".conv_one.weight"], st[p + ".conv_one.bias"] = _conv_params_bias(o1, i1, 3)
            st[p + ".conv_two.weight"], st[p + ".conv_two.bias"] = _conv_params_bias(o2, i2, 3)
            _bn_entries(st, p + ".batch_norm_one", o1)
            _bn_entries(st, p + ".batch_norm_two", o2)
        st[f"{enc}.net.5.weight"], st[f"{enc}.net.5.bias"] = _linear_params(hid, CONVBLOCK_FLAT[enc])
    for idx, (o, i) in zip((0, 3, 5), ((hidden_dim, audio_hidden + image_hidden), (hidden_dim // 2, hidden_dim), (NUM_CLASSES, hidden_dim // 2))):
        st[f"net.{idx}.weight"], st[f"net.{idx}.bias"] = _linear_params(o, i)
    return st


def convblock_encoder_forward(state: Dict[str, Tensor], enc: str, x: Tensor, training: bool, taps: Optional[Dict[str, Tensor]] = None,
                              emulate_bf16: bool = False, forced: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """MNISTAudio.forward / MNISTImage.forward (avmnist.py:108-118, 177-184).  ``taps`` / ``emulate_bf16`` / ``forced``: the testing
    aids of resnet_forward (tap names: net.<slot>.conv_one, .relu_one, .conv_two, net.<slot> = block output, net.<slot+1> = pooled)."""
    eb = emulate_bf16
    if x.dim() == 3:
        x = x.unsqueeze(1)

    def pt(name: str, t: Tensor) -> Tensor:
        key = f"{enc}.{name}"
        if forced is not None and key in forced:
            t = t + (forced[key][:, :t.shape[1]].to(t.dtype) - t).detach()  # the CUDA path stores 64-channel padded tensors
        if taps is not None:
            taps[key] = t
        return t

    x = _q(x, eb)
    for slot, pool in zip((0, 2), CONVBLOCK_POOLS[enc]):
        p = f"{enc}.net.{slot}"
        x = pt(f"net.{slot}.conv_one", _q(F.conv2d(_qg(x, eb), _qw(state[p + ".conv_one.weight"], eb), state[p + ".conv_one.bias"], stride=1, padding=1), eb))
        x = pt(f"net.{slot}.relu_one", _q(F.relu(_batch_norm(state, p + ".batch_norm_one", x, training, True)), eb))
        x = pt(f"net.{slot}.conv_two", _q(F.conv2d(x, _qw(state[p + ".conv_two.weight"], eb), state[p + ".conv_two.bias"], stride=1, padding=1), eb))
        x = pt(f"net.{slot}", _q(F.relu(_batch_norm(state, p + ".batch_norm_two", x, training, True)), eb))
        x = pt(f"net.{slot + 1}", F.max_pool2d(x, kernel_size=pool))
    return F.linear(torch.flatten(x, 1), state[f"{enc}.net.5.weight"], state[f"{enc}.net.5.bias"])


def convblock_train_step(state: "OrderedDict[str, Tensor]", opt_state: Dict, A: Tensor, I: Tensor, labels: Tensor,
                         dropout_mask: Optional[Tensor] = None, dropout_p: float = 0.5, lr: float = 5e-4, weight_decay: float = 1e-4,
                         apply_update: bool = True, emulate_bf16: bool = False, forced: Optional[Dict[str, Tensor]] = None) -> Dict[str, object]:
    params = {k: v for k, v in state.items() if is_parameter(k)}
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    work = dict(state)
    work.update(leaves)
    logits = head_forward(work, convblock_encoder_forward(work, "audio_encoder", A, True, None, emulate_bf16, forced),
                          convblock_encoder_forward(work, "image_encoder", I, True, None, emulate_bf16, forced), dropout_mask, dropout_p)
    loss = total_loss(logits, labels)
    gl = torch.autograd.grad(loss, list(leaves.values()))
    grads = dict(zip(leaves.keys(), gl))
    for k in state:
        if k.endswith("num_batches_tracked"):
            state[k] = work[k]
    if apply_update:
        with torch.no_grad():
            adam_step(params, grads, opt_state, lr=lr, weight_decay=weight_decay)
    return {"loss": float(loss.item()), "logits": logits.detach(), "predictions": torch.softmax(logits.detach(), 1).argmax(1), "grads": grads}


# ----------------------------------------------------------------------------------------------
# e -- data parallel semantics: N replicas, per-replica BatchNorm, gradients averaged
# ----------------------------------------------------------------------------------------------
def data_parallel_grads(
    state: "OrderedDict[str, Tensor]", shards: Sequence[Tuple[Tensor, Tensor, Tensor, Optional[Tensor]]], dropout_p: float = 0.5
) -> Dict[str, Tensor]:
    """Mean over replicas of each replica's gradient on its own shard (no SyncBN)."""
    acc: Dict[str, Tensor] = {}
    for A, I, y, dm in shards:
        st = OrderedDict((k, v.clone()) for k, v in state.items())
        out = train_step(st, {}, A, I, y, dm, dropout_p, apply_update=False)
        for k, g in out["grads"].items():
            acc[k] = g.clone() if k not in acc else acc[k] + g
    return {k: v / len(shards) for k, v in acc.items()}


# ----------------------------------------------------------------------------------------------
# a13 -- FedAvg (no reference implementation exists: train_congruent_federated.py is 0 bytes;
# textbook McMahan et al. weighted average; parity UNPINNED by the reference)
# ----------------------------------------------------------------------------------------------
def fedavg(states: Sequence[Dict[str, Tensor]], num_samples: Sequence[float]) -> "OrderedDict[str, Tensor]":
    total = float(sum(num_samples))
    out: "OrderedDict[str, Tensor]" = OrderedDict()
    for k in states[0]:
        if k.endswith("num_batches_tracked"):
            out[k] = states[0][k].clone()
            continue
        acc = torch.zeros_like(states[0][k], dtype=torch.float32)
        for st, n in zip(states, num_samples):
            acc += st[k].float() * (float(n) / total)
        out[k] = acc
    return out


# ----------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------
def synthetic_batch(
    batch: int, seed: int, audio_hw: Tuple[int, int] = (112, 112), audio_missing_rate: float = 0.2, mask_seed: int = 7
) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    A = torch.rand(batch, *audio_hw, generator=g)
    I = torch.rand(batch, 1, 28, 28, generator=g)
    y = torch.randint(0, NUM_CLASSES, (batch,), generator=g)
    gm = torch.Generator().manual_seed(mask_seed)
    m_audio = torch.bernoulli(torch.full((batch,), 1.0 - audio_missing_rate), generator=gm)
    m_image = torch.ones(batch)
    drop = torch.bernoulli(torch.full((batch, 128), 0.5), generator=gm)
    return {"audio": A, "image": I, "labels": y, "audio_mask": m_audio, "image_mask": m_image, "dropout_mask": drop}
