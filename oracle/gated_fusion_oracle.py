"""CPU oracle for the MML_Suite MMIMDb gated late-fusion training step (BASELINE config 3).  TEST INFRASTRUCTURE ONLY.

Same rules as ``late_fusion_oracle.py``: a functional fp32 restatement over a flat ``state`` dict carrying exactly the
reference's ``state_dict()`` keys (38 entries); only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU legs may import
it.  Parity pinning: ``oracle/make_golden.py`` runs the UNMODIFIED reference ``MMIMDb`` (imported from /root/reference in
the build container) on seeded inputs, refuses to write fixtures unless this restatement reproduces it, and stores
``tests/golden/mmimdb_*.npz``; ``tests/test_oracle_golden.py`` re-checks the restatement against those vectors.

Reference files followed (paths relative to /root/reference/MML_Suite):
  models/mmimdb.py:63-92     MMIMDbModalityEncoder  = BatchNorm1d(in) -> Linear(in, out)
  models/mmimdb.py:20-60     MLPGenreClassifier     = BN1d, MaxOut, Dropout(.5), BN1d, MaxOut, Dropout(.5), BN1d, Linear
  models/mmimdb.py:166-245   MMIMDb.forward / train_step
  models/gates/gated_bimodal.py:40-60  GatedBiModalNetwork.forward (GMU, scalar gate per sample)
  models/maxout.py:27-41     MaxOut.forward (max over ``num_units`` bias-free Linear layers)
  experiment_utils/loss.py:52          "bce_with_logits" -> BCEWithLogitsLoss() (mean over batch x classes)
  data/mmimdb.py:73-77       missing patterns it / i / t  (x * m per modality, base_dataset.py:71)
  configs/mmimdb/centralised/mmimdb_baseline.yaml:14-46  sizes 4096/300 -> 512 -> 23, Adam lr 1e-5 wd 1e-3
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from late_fusion_oracle import (BN_EPS, BN_MOMENTUM, _bn_entries, _linear_params, _q, _qg, _qw, adam_step, apply_missing_mask)

Tensor = torch.Tensor
DROPOUT_P = 0.5  # fixed in MLPGenreClassifier (mmimdb.py:42,45)
PATTERNS = {"it": (1.0, 1.0), "i": (1.0, 0.0), "t": (0.0, 1.0)}  # (image mask, text mask), data/mmimdb.py:73-77


# ----------------------------------------------------------------------------------------------
# construction: same torch RNG draws, in the same order, as
#   MMIMDb(MMIMDbModalityEncoder(4096,512), MMIMDbModalityEncoder(300,512), GatedBiModalNetwork(512,512,512,512),
#          classifier=MLPGenreClassifier(512,23,512))   built in YAML order (image, text, gate, classifier)
# ----------------------------------------------------------------------------------------------
def _linear_nobias(out_f: int, in_f: int) -> Tensor:
    import math

    w = torch.empty(out_f, in_f)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    return w


def init_mmimdb_state(image_dim: int = 4096, text_dim: int = 300, embed: int = 512, hidden: int = 512, classes: int = 23) -> "OrderedDict[str, Tensor]":
    st: "OrderedDict[str, Tensor]" = OrderedDict()
    for prefix, d in (("image_model", image_dim), ("text_model", text_dim)):
        _bn_entries(st, prefix + ".net.0", d)
        st[prefix + ".net.1.weight"], st[prefix + ".net.1.bias"] = _linear_params(embed, d)
    st["fusion_module.fc_one.weight"] = _linear_nobias(embed, embed)
    st["fusion_module.fc_two.weight"] = _linear_nobias(embed, embed)
    st["fusion_module.hidden_sigmoid.weight"] = _linear_nobias(1, 2 * embed)
    _bn_entries(st, "mm_mlp.net.0", embed)
    st["mm_mlp.net.1.layers.0.weight"] = _linear_nobias(hidden, embed)
    st["mm_mlp.net.1.layers.1.weight"] = _linear_nobias(hidden, embed)
    _bn_entries(st, "mm_mlp.net.3", hidden)
    st["mm_mlp.net.4.layers.0.weight"] = _linear_nobias(hidden, hidden)
    st["mm_mlp.net.4.layers.1.weight"] = _linear_nobias(hidden, hidden)
    _bn_entries(st, "mm_mlp.net.6", hidden)
    st["mm_mlp.net.7.weight"], st["mm_mlp.net.7.bias"] = _linear_params(classes, hidden)
    return st


def init_mmimdb_pooling_state(pooling_type: str = "max", image_dim: int = 4096, text_dim: int = 300, embed: int = 512, hidden: int = 512,
                              classes: int = 23, pool_hidden: int = 512) -> "OrderedDict[str, Tensor]":
    """MMIMDb(..., multimodal_pooling={...}) (mmimdb.py:132-146, pooling.py:17-74).  RNG order of mmimdb_pooling.yaml: the two
    encoders and the classifier are built by the YAML loader, MultimodalPooling inside MMIMDb.__init__ afterwards; the
    state_dict order is registration order (image_model, text_model, fusion_module, mm_mlp)."""
    enc: "OrderedDict[str, Tensor]" = OrderedDict()
    for prefix, d in (("image_model", image_dim), ("text_model", text_dim)):
        _bn_entries(enc, prefix + ".net.0", d)
        enc[prefix + ".net.1.weight"], enc[prefix + ".net.1.bias"] = _linear_params(embed, d)
    mlp: "OrderedDict[str, Tensor]" = OrderedDict()
    _bn_entries(mlp, "mm_mlp.net.0", embed)
    mlp["mm_mlp.net.1.layers.0.weight"] = _linear_nobias(hidden, embed)
    mlp["mm_mlp.net.1.layers.1.weight"] = _linear_nobias(hidden, embed)
    _bn_entries(mlp, "mm_mlp.net.3", hidden)
    mlp["mm_mlp.net.4.layers.0.weight"] = _linear_nobias(hidden, hidden)
    mlp["mm_mlp.net.4.layers.1.weight"] = _linear_nobias(hidden, hidden)
    _bn_entries(mlp, "mm_mlp.net.6", hidden)
    mlp["mm_mlp.net.7.weight"], mlp["mm_mlp.net.7.bias"] = _linear_params(classes, hidden)
    fus: "OrderedDict[str, Tensor]" = OrderedDict()
    fus["fusion_module.proj_a.weight"], fus["fusion_module.proj_a.bias"] = _linear_params(embed, embed)
    fus["fusion_module.proj_b.weight"], fus["fusion_module.proj_b.bias"] = _linear_params(embed, embed)
    pt = pooling_type.lower()
    if pt == "attention":
        fus["fusion_module.attention_layer.0.weight"], fus["fusion_module.attention_layer.0.bias"] = _linear_params(pool_hidden, 2 * embed)
        fus["fusion_module.attention_layer.2.weight"], fus["fusion_module.attention_layer.2.bias"] = _linear_params(2, pool_hidden)
    elif pt == "gated":
        fus["fusion_module.gate_layer.0.weight"], fus["fusion_module.gate_layer.0.bias"] = _linear_params(pool_hidden, 2 * embed)
        fus["fusion_module.gate_layer.2.weight"], fus["fusion_module.gate_layer.2.bias"] = _linear_params(1, pool_hidden)
    st: "OrderedDict[str, Tensor]" = OrderedDict()
    for part in (enc, fus, mlp):
        st.update(part)
    return st


def pooling_fusion(st: Dict[str, Tensor], ei: Tensor, et: Tensor, pooling_type: str, training: bool,
                   pool_masks: Optional[Tuple[Tensor, Tensor]], pool_p: float, q: bool) -> Tensor:
    """MultimodalPooling.forward (pooling.py:76-126)."""
    a = torch.tanh(_qg(_q(F.linear(ei, _qw(st["fusion_module.proj_a.weight"], q)), q), q) + st["fusion_module.proj_a.bias"])
    b = torch.tanh(_qg(_q(F.linear(et, _qw(st["fusion_module.proj_b.weight"], q)), q), q) + st["fusion_module.proj_b.bias"])
    if training and pool_masks is not None and pool_p > 0:
        a = a * pool_masks[0] / (1.0 - pool_p)
        b = b * pool_masks[1] / (1.0 - pool_p)
    pt = pooling_type.lower()
    if pt == "max":
        return torch.max(a, b)
    if pt in ("avg", "average"):
        return (a + b) / 2
    if pt == "sum":
        return a + b
    combined = torch.cat([a, b], dim=1)
    if pt == "attention":
        h = torch.tanh(F.linear(combined, st["fusion_module.attention_layer.0.weight"], st["fusion_module.attention_layer.0.bias"]))
        att = torch.softmax(F.linear(h, st["fusion_module.attention_layer.2.weight"], st["fusion_module.attention_layer.2.bias"]), dim=1)
        return att[:, 0:1] * a + att[:, 1:2] * b
    if pt == "gated":
        h = torch.tanh(F.linear(combined, st["fusion_module.gate_layer.0.weight"], st["fusion_module.gate_layer.0.bias"]))
        gate = torch.sigmoid(F.linear(h, st["fusion_module.gate_layer.2.weight"], st["fusion_module.gate_layer.2.bias"]))
        return gate * a + (1 - gate) * b
    raise ValueError(f"Unknown pooling type: {pooling_type}")


def is_parameter(key: str) -> bool:
    return not key.endswith(("running_mean", "running_var", "num_batches_tracked"))


# ----------------------------------------------------------------------------------------------
# forward (mmimdb.py:166-200)
# ----------------------------------------------------------------------------------------------
def _bn1d(st: Dict[str, Tensor], prefix: str, x: Tensor, training: bool) -> Tensor:
    y = F.batch_norm(x, st[prefix + ".running_mean"], st[prefix + ".running_var"], st[prefix + ".weight"], st[prefix + ".bias"],
                     training, BN_MOMENTUM, BN_EPS)
    if training:
        st[prefix + ".num_batches_tracked"] = st[prefix + ".num_batches_tracked"] + 1
    return y


def gated_fusion_forward(st: Dict[str, Tensor], I: Tensor, T: Tensor, training: bool,
                         dropout_masks: Optional[Tuple[Tensor, Tensor]] = None, emulate_bf16: bool = False,
                         taps: Optional[Dict[str, Tensor]] = None, pooling_type: Optional[str] = None,
                         pool_masks: Optional[Tuple[Tensor, Tensor]] = None, pool_p: float = 0.0) -> Tensor:
    """logits [B, classes].  ``dropout_masks`` = two {0,1} keep-masks [B, hidden] (train mode; None = no dropout).
    ``emulate_bf16`` rounds where the B200 path stores bf16 (GEMM operands and outputs); debugging aid only."""
    q = emulate_bf16
    tap = (lambda k, v: taps.__setitem__(k, v.detach())) if taps is not None else (lambda k, v: None)
    # encoders: BatchNorm1d -> Linear
    xi = _q(_bn1d(st, "image_model.net.0", I, training), q)
    xt = _q(_bn1d(st, "text_model.net.0", T, training), q)
    ei = _qg(_q(F.linear(xi, _qw(st["image_model.net.1.weight"], q), _qw(st["image_model.net.1.bias"], q)), q), q)
    et = _qg(_q(F.linear(xt, _qw(st["text_model.net.1.weight"], q), _qw(st["text_model.net.1.bias"], q)), q), q)
    tap("image_embedding", ei)
    tap("text_embedding", et)
    if "fusion_module.proj_a.weight" in st:  # multimodal_pooling variant (mmimdb.py:132-146)
        z = pooling_fusion(st, ei, et, pooling_type or "max", training, pool_masks, pool_p, q)
        return _mlp_tail(st, z, training, dropout_masks, q, tap)
    # GMU (gated_bimodal.py:40-60): one scalar gate per sample
    h1 = torch.tanh(_qg(_q(F.linear(ei, _qw(st["fusion_module.fc_one.weight"], q)), q), q))
    h2 = torch.tanh(_qg(_q(F.linear(et, _qw(st["fusion_module.fc_two.weight"], q)), q), q))
    gate = torch.sigmoid(F.linear(torch.cat([h1, h2], dim=1), st["fusion_module.hidden_sigmoid.weight"]))
    z = gate.view(-1, 1) * h1 + (1 - gate).view(-1, 1) * h2
    tap("gate", gate)
    return _mlp_tail(st, z, training, dropout_masks, q, tap)


def _mlp_tail(st, z, training, dropout_masks, q, tap) -> Tensor:
    tap("fused", z)
    # MaxOut MLP (mmimdb.py:37-46)
    x = z
    for bn, mo, mi in (("mm_mlp.net.0", "mm_mlp.net.1", 0), ("mm_mlp.net.3", "mm_mlp.net.4", 1)):
        x = _qg(_q(_bn1d(st, bn, x, training), q), q)
        a = _qg(_q(F.linear(x, _qw(st[mo + ".layers.0.weight"], q)), q), q)
        b = _qg(_q(F.linear(x, _qw(st[mo + ".layers.1.weight"], q)), q), q)
        x = torch.max(a, b)
        if training and dropout_masks is not None:
            x = x * dropout_masks[mi] / (1.0 - DROPOUT_P)
    x = _qg(_bn1d(st, "mm_mlp.net.6", x, training), q)
    return F.linear(x, st["mm_mlp.net.7.weight"], st["mm_mlp.net.7.bias"])


def total_loss(logits: Tensor, labels: Tensor) -> Tensor:
    """LossFunctionGroup({"bce": bce_with_logits x 1.0}): BCEWithLogitsLoss() defaults = mean over all B x classes."""
    return F.binary_cross_entropy_with_logits(logits, labels) * 1.0


def train_step(st: "OrderedDict[str, Tensor]", opt_state: Dict, I: Tensor, T: Tensor, labels: Tensor,
               dropout_masks: Optional[Tuple[Tensor, Tensor]] = None, lr: float = 1e-5, weight_decay: float = 1e-3,
               apply_update: bool = True, emulate_bf16: bool = False, threshold: float = 0.5, pooling_type: Optional[str] = None,
               pool_masks: Optional[Tuple[Tensor, Tensor]] = None, pool_p: float = 0.0) -> Dict[str, object]:
    """mmimdb.py:202-245 body: zero_grad, forward (train mode), BCE, backward, Adam.step, sigmoid > threshold."""
    params = {k: v for k, v in st.items() if is_parameter(k)}
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    work = dict(st)
    work.update(leaves)
    logits = gated_fusion_forward(work, I, T, True, dropout_masks, emulate_bf16, pooling_type=pooling_type, pool_masks=pool_masks, pool_p=pool_p)
    loss = total_loss(logits, labels)
    gl = torch.autograd.grad(loss, list(leaves.values()))
    grads = dict(zip(leaves.keys(), gl))
    for k in st:
        if k.endswith("num_batches_tracked"):
            st[k] = work[k]
    if apply_update:
        with torch.no_grad():
            adam_step(params, grads, opt_state, lr=lr, weight_decay=weight_decay)
    preds = (torch.sigmoid(logits.detach()) > threshold).to(torch.int64)
    return {"loss": float(loss.item()), "logits": logits.detach(), "predictions": preds, "grads": grads}


@torch.no_grad()
def validation_step(st: Dict[str, Tensor], I: Tensor, T: Tensor, labels: Tensor, threshold: float = 0.5,
                    pooling_type: Optional[str] = None) -> Dict[str, object]:
    logits = gated_fusion_forward(dict(st), I, T, False, pooling_type=pooling_type)
    return {"loss": float(total_loss(logits, labels).item()), "logits": logits, "predictions": (torch.sigmoid(logits) > threshold).to(torch.int64)}


# ----------------------------------------------------------------------------------------------
# monomodal pre-training of ONE MMIMDb encoder (train_monomodal.py:64-92,224-232 with configs/mmimdb/mono/*.yaml):
#   MonomodalEncoder(MMIMDbModalityEncoder(in, 512), 512, 23) = BatchNorm1d -> Linear -> Linear(512, 23), BCEWithLogitsLoss on the
#   multi-hot genre vector, Adam(lr 1e-5, weight_decay 1e-3), predictions = sigmoid > 0.5 (:243).  CUDA path: mml_b200/mono.py.
# ----------------------------------------------------------------------------------------------
def init_mono_vector_state(in_dim: int, embed: int = 512, classes: int = 23) -> "OrderedDict[str, Tensor]":
    """Same RNG draws, in the same order, as ``MonomodalEncoder(MMIMDbModalityEncoder(in_dim, embed), embed, classes)``."""
    st: "OrderedDict[str, Tensor]" = OrderedDict()
    _bn_entries(st, "encoder.net.0", in_dim)
    st["encoder.net.1.weight"], st["encoder.net.1.bias"] = _linear_params(embed, in_dim)
    st["classifier.weight"], st["classifier.bias"] = _linear_params(classes, embed)
    return st


def mono_vector_forward(st: Dict[str, Tensor], x: Tensor, training: bool, emulate_bf16: bool = False) -> Tensor:
    q = emulate_bf16
    xn = _q(_bn1d(st, "encoder.net.0", x, training), q)
    emb = _qg(_q(F.linear(xn, _qw(st["encoder.net.1.weight"], q), _qw(st["encoder.net.1.bias"], q)), q), q)
    return F.linear(emb.reshape(emb.shape[0], -1), st["classifier.weight"], st["classifier.bias"])


def mono_vector_train_step(st: "OrderedDict[str, Tensor]", opt_state: Dict, x: Tensor, labels: Tensor, lr: float = 1e-5,
                           weight_decay: float = 1e-3, apply_update: bool = True, emulate_bf16: bool = False,
                           threshold: float = 0.5) -> Dict[str, object]:
    params = {k: v for k, v in st.items() if is_parameter(k)}
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    work = dict(st)
    work.update(leaves)
    logits = mono_vector_forward(work, x, True, emulate_bf16)
    loss = total_loss(logits, labels)
    gl = torch.autograd.grad(loss, list(leaves.values()))
    grads = dict(zip(leaves.keys(), gl))
    for k in st:
        if k.endswith("num_batches_tracked"):
            st[k] = work[k]
    if apply_update:
        with torch.no_grad():
            adam_step(params, grads, opt_state, lr=lr, weight_decay=weight_decay)
    preds = (torch.sigmoid(logits.detach()) > threshold).to(torch.int64)
    return {"loss": float(loss.item()), "logits": logits.detach(), "predictions": preds, "grads": grads}


@torch.no_grad()
def mono_vector_validation_step(st: Dict[str, Tensor], x: Tensor, labels: Tensor, threshold: float = 0.5) -> Dict[str, object]:
    logits = mono_vector_forward(dict(st), x, False)
    return {"loss": float(total_loss(logits, labels).item()), "logits": logits, "predictions": (torch.sigmoid(logits) > threshold).to(torch.int64)}


# ----------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------
def synthetic_batch(batch: int, seed: int, image_dim: int = 4096, text_dim: int = 300, hidden: int = 512, classes: int = 23,
                    rates: Sequence[float] = (0.5, 0.25, 0.25)) -> Dict[str, Tensor]:
    """I ~ |N(0,1)| (VGG fc features are >= 0), T ~ N(0, 0.3), labels ~ Bernoulli(0.1) fp32, pattern per sample drawn from
    it / i / t with probabilities ``rates``; masks are what base_dataset.py:46-59 would have drawn for those patterns."""
    g = torch.Generator().manual_seed(seed)
    I = torch.randn(batch, image_dim, generator=g).abs()
    T = torch.randn(batch, text_dim, generator=g) * 0.3
    y = (torch.rand(batch, classes, generator=g) < 0.1).float()
    pat = torch.multinomial(torch.tensor(list(rates)), batch, replacement=True, generator=g)
    names = [("it", "i", "t")[int(p)] for p in pat]
    mi = torch.tensor([PATTERNS[n][0] for n in names])
    mt = torch.tensor([PATTERNS[n][1] for n in names])
    d1 = (torch.rand(batch, hidden, generator=g) < 0.5).float()
    d2 = (torch.rand(batch, hidden, generator=g) < 0.5).float()
    pa = (torch.rand(batch, hidden, generator=g) >= 0.1).float()  # MultimodalPooling dropout 0.1 (mmimdb_pooling.yaml), one mask per branch
    pb = (torch.rand(batch, hidden, generator=g) >= 0.1).float()
    return {"pool_masks": (pa, pb),"image": I, "text": T, "labels": y, "pattern_name": names, "image_mask": mi, "text_mask": mt, "dropout_masks": (d1, d2),
            "image_masked": apply_missing_mask(I, mi), "text_masked": apply_missing_mask(T, mt)}
