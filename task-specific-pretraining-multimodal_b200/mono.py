"""Monomodal encoder pre-training -- drop-in for ``MonomodalEncoder`` (MML_Suite/train_monomodal.py:64-420) on a B200.

``MonomodalEncoder(encoder, output_dim, num_classes)`` wraps ONE modality encoder and a ``Linear(output_dim, num_classes)``;
its ``train_step`` is encoder forward -> classifier -> CrossEntropy -> backward -> Adam on the ORIGINAL (un-masked) modality
tensor.  Here that is the same fused machinery as the late-fusion step with one encoder: the tcgen05 conv / BatchNorm
schedule of ``engine.EncoderPlan`` plus a small tail (encoder fc, classifier, softmax-CE; ``mml_mono_head_fwd/bwd``), one
CUDA graph per (batch, input size).  ``get_encoder().state_dict()`` is what ``train_monomodal.py:790-801`` saves and
``train_multimodal.py:186-187`` loads into the fusion model -- same names, shapes and dtypes as the reference.

Two encoder families are built: the ResNet encoders of ``mml_b200.resnet`` (AVMNIST; single-label CrossEntropy) and the
``MMIMDbModalityEncoder`` of ``mml_b200.mmimdb`` (``configs/mmimdb/mono/*.yaml``: BatchNorm1d -> Linear on a feature vector,
multi-hot genre targets, ``bce_with_logits``, predictions ``sigmoid > 0.5``, train_monomodal.py:243) -- the latter reuses the
MMIMDb step's kernels (fused BatchNorm1d, the tcgen05 GEMM with the bias as an extra input column, the BCE head).  Anything else
raises.  The MOSI encoders of ``mml_b200.utt_fusion`` (``configs/mosi/mono/*.yaml``: ``LSTMEncoder`` on [B, T, 5 | 20] sequences,
``TextCNN`` on [B, T, 768]; cross entropy over 3 classes, no gradient clip) run on the UttFusion step's kernels (LSTM forward / BPTT,
tcgen05 convolutions over time, fused ReLU + max-over-time + dropout, dense layers).  File-path batches (the reference loads ``.pt`` paths inside ``train_step``,
:138-160) are host I/O and are not accepted -- pass tensors.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .avmnist import _copy_in
from .engine import ALIGN, BF16, BN_EPS, BN_MOMENTUM, EncoderPlan, FlatState, _round_up

_SKIP = ("labels", "label", "genres", "imdb_ids", "pattern_name", "missing_masks", "sample_idx")


class _MonoPlan:
    def __init__(self, eng: "MonoEngine", B: int, H: int, W: int):
        self.eng, self.B = eng, B
        fs, dev, model = eng.fs, eng.device, eng.model
        self.enc = EncoderPlan(fs, model.encoder, "encoder.", B, H, W, train=True)
        self.enc.wgrad_stream = torch.cuda.Stream(device=dev)
        params = dict(model.named_parameters())

        def par(flat, name):
            return fs.flat_slice(flat, name).view(params[name].shape)

        self.fc_w, self.fc_b = par(fs.P, "encoder.fc.weight"), par(fs.P, "encoder.fc.bias")
        self.cls_w, self.cls_b = par(fs.P, "classifier.weight"), par(fs.P, "classifier.bias")
        self.d_fc_w, self.d_fc_b = par(fs.G, "encoder.fc.weight"), par(fs.G, "encoder.fc.bias")
        self.d_cls_w, self.d_cls_b = par(fs.G, "classifier.weight"), par(fs.G, "classifier.bias")
        E, NC = self.fc_w.shape[0], self.cls_w.shape[0]
        if self.cls_w.shape[1] != E:
            raise ValueError(f"classifier expects {self.cls_w.shape[1]} features but the encoder produces {E}")
        self.labels = torch.zeros(B, device=dev, dtype=torch.int64)
        self.emb, self.demb = torch.zeros(B, E, device=dev), torch.zeros(B, E, device=dev)
        self.logits, self.dlogits = torch.zeros(B, NC, device=dev), torch.zeros(B, NC, device=dev)
        self.row_loss, self.loss = torch.zeros(B, device=dev), torch.zeros(1, device=dev)
        self.pred = torch.zeros(B, device=dev, dtype=torch.int32)
        self.h_loss = torch.zeros(1).pin_memory()
        self.h_pred = torch.zeros(B, dtype=torch.int32).pin_memory()
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.eager_steps = 0
        self.launches_per_step = 0

    def _head_fwd(self, with_loss: bool, with_grad: bool) -> None:
        ops.mono_head_fwd(self.enc.pooled, self.fc_w, self.fc_b, self.cls_w, self.cls_b, self.labels if with_loss else None, self.emb, self.logits,
                          self.dlogits if with_grad else None, self.row_loss if with_loss else None, self.loss if with_loss else None, self.pred, 1.0)

    def run_train(self) -> None:
        fs = self.eng.fs
        fs.G.zero_()
        self.enc.stat_arena.zero_()
        for op in self.enc.fwd_train:
            op()
        self._head_fwd(True, True)
        ops.mono_head_bwd(self.enc.pooled, self.emb, self.dlogits, self.fc_w, self.cls_w, self.d_fc_w, self.d_fc_b, self.d_cls_w, self.d_cls_b,
                          self.demb, self.enc.dpooled)
        for op in self.enc.bwd:
            op()
        fs.NBT += 1
        if self.eng.allreduce is not None:
            self.eng.allreduce(self, 0, update=lambda: fs.adam(0, fs.total, True))
        else:
            fs.adam(0, fs.total, True)

    def train_step(self) -> None:
        eng = self.eng
        if getattr(self, "_range_version", None) != eng.fs.range_version:
            self.graph, self._range_version = None, eng.fs.range_version
        if not eng.use_graphs:
            return self.run_train()
        if self.graph is None:
            if self.eager_steps < 2:
                before = ops.launch_count(eng.device.index)
                self.run_train()
                self.launches_per_step = ops.launch_count(eng.device.index) - before
                self.eager_steps += 1
                return
            torch.cuda.synchronize(eng.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.run_train()
        self.graph.replay()

    def run_forward(self, training: bool, with_loss: bool) -> None:
        if training:
            self.enc.stat_arena.zero_()
        for op in (self.enc.fwd_train if training else self.enc.fwd_eval):
            op()
        self._head_fwd(with_loss, False)
        if training:
            self.eng.fs.NBT += 1


class MonoEngine:
    def __init__(self, model: nn.Module, device: torch.device):
        self.model, self.device = model, device
        self.fs = FlatState(model, device)
        self.plans: Dict[Tuple[int, int, int], _MonoPlan] = {}
        self.world = 1
        self.allreduce = None
        self.use_graphs = True

    def plan_for(self, B: int, H: int, W: int) -> _MonoPlan:
        key = (B, H, W)
        plan = self.plans.get(key)
        if plan is None:
            plan = self.plans[key] = _MonoPlan(self, B, H, W)
        return plan


class _VecMonoPlan:
    """MonomodalEncoder(MMIMDbModalityEncoder(D, E), E, NC) for a fixed batch size: BatchNorm1d -> Linear -> Linear -> BCE.

    Layout as in ``gated_engine``: the encoder Linear is stored augmented ([E][ld], ld = D + 1 rounded up to 64, bias in column D)
    and its input rows carry a constant 1 there, so the tensor-core GEMM adds the bias and its wgrad yields the bias gradient."""

    def __init__(self, eng: "VecMonoEngine", B: int):
        self.eng, self.B = eng, B
        fs, dev, model = eng.fs, eng.device, eng.model
        params = dict(model.named_parameters())
        D, E, NC = params["encoder.net.0.weight"].numel(), params["encoder.net.1.weight"].shape[0], params["classifier.weight"].shape[0]
        if E % 64:
            raise NotImplementedError("the encoder's output width must be a multiple of 64 for the tensor-core GEMM")
        if params["classifier.weight"].shape[1] != E:
            raise ValueError(f"classifier expects {params['classifier.weight'].shape[1]} features but the encoder produces {E}")
        self.D, self.E, self.NC = D, E, NC
        L = _round_up(D + 1, ALIGN)

        def f32(*shape):
            return torch.zeros(*shape, device=dev)

        def b16(*shape):
            return torch.zeros(*shape, device=dev, dtype=BF16)

        def par(flat, name):
            return fs.flat_slice(flat, name).view(params[name].shape)

        self.x, self.labels = f32(B, D), f32(B, NC)
        self.xn, self.xh, self.inv = b16(B, L), f32(B, D), f32(D)
        self.xn[:, D] = 1.0  # the bias column
        self.e, self.ef = b16(B, E), f32(B, E)
        self.logits, self.dlogits, self.loss = f32(B, NC), f32(B, NC), f32(1)
        self.pred = torch.zeros(B, NC, device=dev, dtype=torch.uint8)
        self.scratch = f32(ops.bce_head_scratch_floats(B))
        self.de, self.dxn = b16(B, E), b16(B, L)
        self.h_loss = torch.zeros(1).pin_memory()
        self.h_pred = torch.zeros(B, NC, dtype=torch.uint8).pin_memory()
        om, ov = fs.buf_offsets["encoder.net.0.running_mean"], fs.buf_offsets["encoder.net.0.running_var"]
        gamma, beta = par(fs.P, "encoder.net.0.weight"), par(fs.P, "encoder.net.0.bias")
        self.f_bn = ops.bn1d_fwd_desc(ops.BN1D_INPUT, B, D, gamma, beta, fs.S[om:om + D], fs.S[ov:ov + D], x=self.x, xhat=self.xh, invstd=self.inv,
                                      y_bf16=self.xn, momentum=BN_MOMENTUM, eps=BN_EPS)
        self.b_bn = ops.bn1d_bwd_desc(ops.BN1D_INPUT, B, D, self.dxn, self.xh, gamma, self.inv, par(fs.G, "encoder.net.0.weight"),
                                      par(fs.G, "encoder.net.0.bias"))
        self.gem = (ops.make_geom(B, 1, 1, L, E, 1, 1, 1, 0), fs.aug_matrix(fs.Wb, "encoder.net.1.weight"), fs.aug_matrix(fs.G, "encoder.net.1.weight"))
        self.cls_w, self.cls_b = par(fs.P, "classifier.weight"), par(fs.P, "classifier.bias")
        self.d_cls_w, self.d_cls_b = par(fs.G, "classifier.weight"), par(fs.G, "classifier.bias")
        self.wgrad_ws = ops.WgradScratch(dev)
        self.threshold = 0.5  # train_monomodal.py:243
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.eager_steps = 0
        self.launches_per_step = 0

    def run_forward(self, training: bool, with_loss: bool, with_grad: bool = False) -> None:
        ops.bn1d_fwd(self.f_bn, training)
        ops.conv_fprop(self.gem[0], self.xn, self.gem[1], self.e, None)
        ops.cast_bf16_f32(self.e, self.ef)
        ops.bce_head_fwd(self.ef, self.cls_w, self.cls_b, self.labels if with_loss else None, self.logits, self.loss if with_loss else None,
                         self.dlogits if with_grad else None, self.pred, self.scratch, self.threshold, 1.0)
        if training:
            self.eng.fs.NBT += 1

    def run_train(self) -> None:
        fs = self.eng.fs
        fs.G.zero_()
        self.run_forward(True, True, True)
        ops.bce_head_bwd(self.dlogits, self.ef, self.cls_w, self.d_cls_w, self.d_cls_b, self.de)
        ops.conv_wgrad(self.gem[0], self.xn, self.de, self.gem[2], self.wgrad_ws)
        ops.conv_dgrad(self.gem[0], self.de, self.gem[1], self.dxn)
        ops.bn1d_bwd(self.b_bn)
        if self.eng.allreduce is not None:
            self.eng.allreduce(self, 0, update=lambda: fs.adam(0, fs.total, True))
        else:
            fs.adam(0, fs.total, True)

    train_step = _MonoPlan.train_step  # same eager-twice-then-graph protocol


class _SeqMonoPlan:
    """MonomodalEncoder(LSTMEncoder | TextCNN, 64, NC) for a fixed (batch, sequence length): encoder -> Linear -> cross entropy.

    The schedule is the matching slice of the UttFusion step (utt_fusion._UttPlan): ``mml_lstm_fwd / mml_lstm_bwd`` with h_T as the
    embedding, or three ``Conv2d(1, C, (k, 768))`` on the tcgen05 conv path + fused ReLU / max-over-time / dropout + the embd Linear."""

    def __init__(self, eng: "SeqMonoEngine", B: int, T: int):
        self.eng, self.B, self.T = eng, B, T
        fs, dev, enc = eng.fs, eng.device, eng.model.encoder
        params = dict(eng.model.named_parameters())

        def par(flat, name):
            return fs.flat_slice(flat, name).view(params[name].shape)

        f32 = lambda *sh: torch.zeros(*sh, device=dev)
        self.kind = eng.kind
        self.D = enc.input_size
        E = enc.hidden_size
        self.E = E
        self.NC = params["classifier.weight"].shape[0]
        if params["classifier.weight"].shape[1] != E:
            raise ValueError(f"classifier expects {params['classifier.weight'].shape[1]} features but the encoder produces {E}")
        self.x = f32(B, T, self.D)
        self.labels = torch.zeros(B, device=dev, dtype=torch.int64)
        self.emb, self.demb = f32(B, E), f32(B, E)
        self.logits, self.dlogits = f32(B, self.NC), f32(B, self.NC)
        self.row_loss, self.loss = f32(B), f32(1)
        self.pred = torch.zeros(B, device=dev, dtype=torch.int32)
        self.h_loss = torch.zeros(1).pin_memory()
        self.h_pred = torch.zeros(B, dtype=torch.int32).pin_memory()
        self.cls = dict(w=par(fs.P, "classifier.weight"), b=par(fs.P, "classifier.bias"), dw=par(fs.G, "classifier.weight"), db=par(fs.G, "classifier.bias"))
        if self.kind == "lstm":
            names = [f"encoder.rnn.{n}" for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
            self.w, self.dw = [par(fs.P, n) for n in names], [par(fs.G, n) for n in names]
            self.gates, self.cs, self.hs = f32(B, T, 4 * E), f32(B, T, E), f32(B, T, E)
            self.p_drop = 0.0
        else:
            C_ = enc.out_channels
            self.C = C_
            self.x16 = torch.zeros(B, T, self.D, device=dev, dtype=BF16)  # bf16 NHWC [B][T][1][768]
            self.convs = []
            self.wgrad_ws = ops.WgradScratch(dev)
            for i, k in enumerate(enc.kernel_heights):
                name = f"encoder.conv{i + 1}"
                P_ = T - k + 1
                if P_ < 1:
                    raise ValueError(f"sequence length {T} shorter than the TextCNN kernel height {k}")
                self.convs.append(dict(geom=ops.make_geom(B, T, 1, self.D, C_, k, 1, 1, 0), w16=fs.flat_slice(fs.Wb, name + ".weight").view(C_, k, 1, self.D),
                                       dw=fs.flat_slice(fs.G, name + ".weight").view(C_, k, 1, self.D), bias=par(fs.P, name + ".bias"),
                                       dbias=par(fs.G, name + ".bias"), out=torch.zeros(B, P_, C_, device=dev, dtype=BF16),
                                       dout=torch.zeros(B, P_, C_, device=dev, dtype=BF16)))
            NP = len(self.convs) * C_
            self.pooled, self.dpooled = f32(B, NP), f32(B, NP)
            self.arg = torch.zeros(B, NP, device=dev, dtype=torch.int32)
            self.keep = torch.ones(B, NP, device=dev, dtype=torch.uint8)
            self.p_drop = float(enc.dropout.p)
            self.embd = dict(w=par(fs.P, "encoder.embd.0.weight"), b=par(fs.P, "encoder.embd.0.bias"), dw=par(fs.G, "encoder.embd.0.weight"),
                             db=par(fs.G, "encoder.embd.0.bias"))
        self.graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        self.eager_steps = 0
        self.launches_per_step = 0

    def run_forward(self, training: bool, with_loss: bool, with_grad: bool = False) -> None:
        B, E = self.B, self.E
        if self.kind == "lstm":
            ops.lstm_fwd(self.x, *self.w, self.gates, self.cs, self.hs, self.emb)
        else:
            drop = training and self.p_drop > 0
            ops.cast_f32_bf16(self.x, self.x16)
            for i, cv in enumerate(self.convs):
                ops.conv_fprop(cv["geom"], self.x16, cv["w16"], cv["out"], None)
                ops.relumax_fwd(cv["out"], cv["bias"], self.keep if drop else None, 1.0 / (1.0 - self.p_drop) if drop else 1.0, self.pooled, self.arg, i * self.C)
            ops.dense_fwd(self.pooled, self.pooled.shape[1], self.embd["w"], self.embd["b"], None, 1.0, True, self.emb, E, B)
        ops.dense_fwd(self.emb, E, self.cls["w"], self.cls["b"], None, 1.0, False, self.logits, self.NC, B)
        ops.softmax_ce(self.logits, self.labels if with_loss else None, self.dlogits if with_grad else None, self.row_loss if with_loss else None,
                       self.loss if with_loss else None, self.pred, 1.0)

    def run_train(self, own_dropout: bool = True) -> None:
        eng, fs, B, E = self.eng, self.eng.fs, self.B, self.E
        fs.G.zero_()  # the BPTT kernel accumulates its weight gradients with atomics
        if self.p_drop > 0 and own_dropout:
            ops.dropout_mask(self.keep, self.p_drop, eng.seed, fs.step)
        self.run_forward(True, True, True)
        ops.dense_bwd(self.dlogits, self.logits, self.NC, None, 1.0, False, self.emb, E, self.cls["w"], self.demb, E, self.cls["dw"], self.cls["db"], B)
        if self.kind == "lstm":
            ops.lstm_bwd(self.x, self.w[1], self.gates, self.cs, self.hs, self.demb, *self.dw)
        else:
            drop = self.p_drop > 0
            e = self.embd
            ops.dense_bwd(self.demb, self.emb, E, None, 1.0, True, self.pooled, self.pooled.shape[1], e["w"], self.dpooled, self.dpooled.shape[1], e["dw"], e["db"], B)
            for i, cv in enumerate(self.convs):
                ops.relumax_bwd(self.dpooled, self.arg, self.keep if drop else None, 1.0 / (1.0 - self.p_drop) if drop else 1.0, cv["dout"], cv["dbias"], i * self.C)
                ops.conv_wgrad(cv["geom"], self.x16, cv["dout"], cv["dw"], self.wgrad_ws)
        if eng.allreduce is not None:
            eng.allreduce(self, 0, update=lambda: fs.adam(0, fs.total, True))
        else:
            fs.adam(0, fs.total, True)

    def train_step(self, given_dropout: bool = False) -> None:
        eng = self.eng
        key = "train_given" if given_dropout else "train"
        if getattr(self, "_range_version", None) != eng.fs.range_version:
            self.graphs.clear()
            self._range_version = eng.fs.range_version
        if not eng.use_graphs:
            return self.run_train(not given_dropout)
        g = self.graphs.get(key)
        if g is None:
            if self.eager_steps < 2:
                before = ops.launch_count(eng.device.index)
                self.run_train(not given_dropout)
                self.launches_per_step = ops.launch_count(eng.device.index) - before
                self.eager_steps += 1
                return
            torch.cuda.synchronize(eng.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.run_train(not given_dropout)
            self.graphs[key] = g
        g.replay()


class SeqMonoEngine:
    def __init__(self, model: nn.Module, device: torch.device, kind: str):
        self.model, self.device, self.kind = model, device, kind
        self.seed = ops.engine_seed(int(getattr(model, "_mml_client_id", 0)))
        self.fs = FlatState(model, device)
        self.plans: Dict[Tuple[int, int], _SeqMonoPlan] = {}
        self.world = 1
        self.allreduce = None
        self.use_graphs = True

    def plan_for(self, B: int, T: int) -> _SeqMonoPlan:
        plan = self.plans.get((B, T))
        if plan is None:
            plan = self.plans[(B, T)] = _SeqMonoPlan(self, B, T)
        return plan


class VecMonoEngine:
    def __init__(self, model: nn.Module, device: torch.device):
        self.model, self.device = model, device
        self.fs = FlatState(model, device, augment={"encoder.net.1.weight": "encoder.net.1.bias"})
        self.plans: Dict[int, _VecMonoPlan] = {}
        self.world = 1
        self.allreduce = None
        self.use_graphs = True

    def plan_for(self, B: int) -> _VecMonoPlan:
        plan = self.plans.get(B)
        if plan is None:
            plan = self.plans[B] = _VecMonoPlan(self, B)
        return plan


class MonomodalEncoder(nn.Module):
    def __init__(self, encoder: nn.Module, output_dim: int, num_classes: int):
        super().__init__()
        from .mmimdb import MMIMDbModalityEncoder
        from .resnet import ResNetEncoder
        from .utt_fusion import LSTMEncoder, TextCNN

        if not isinstance(encoder, (ResNetEncoder, MMIMDbModalityEncoder, LSTMEncoder, TextCNN)):
            raise NotImplementedError("mml_b200.MonomodalEncoder wraps the ResNet encoders of mml_b200.resnet, the MMIMDbModalityEncoder of "
                                      f"mml_b200.mmimdb and the LSTMEncoder / TextCNN of mml_b200.utt_fusion; got {type(encoder).__name__}")
        self._vector = isinstance(encoder, MMIMDbModalityEncoder)  # feature-vector encoder, multi-label targets
        self._seq = "lstm" if isinstance(encoder, LSTMEncoder) else ("textcnn" if isinstance(encoder, TextCNN) else None)  # [B, T, D] sequences
        self.encoder = encoder
        self.classifier = nn.Linear(output_dim, num_classes)
        self._engine: Optional[MonoEngine] = None
        self._dp = None
        self.world_size = 1

    # ---- plumbing ------------------------------------------------------------------------------------------------------
    def train(self, mode: bool = True):
        super().train(mode)
        self._uniform_mode = bool(mode)
        return self

    def _set_mode(self, training: bool) -> None:
        if getattr(self, "_uniform_mode", None) is not training or self.training is not training:
            self.train(training)

    def _get_engine(self, device) -> MonoEngine:
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("mml_b200.MonomodalEncoder runs on a B200 GPU only: there is no CPU / PyTorch fallback path")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        eng = self._engine
        if eng is None or eng.device != device:
            import weakref
            if self._seq is not None:
                eng = self._engine = SeqMonoEngine(self, device, self._seq)
            elif self._vector:
                eng = self._engine = VecMonoEngine(self, device)
                self.encoder._mml_owner = (weakref.ref(self), "mono")  # encoder(x) outside the step -> self.encode
            else:
                eng = self._engine = MonoEngine(self, device)
                self.encoder._mml_owner = (weakref.ref(eng), "encoder.")  # stand-alone encoder calls share this flat storage
            if self._dp is not None:
                self._dp.attach(eng)
        eng.fs.ensure_fresh()
        return eng

    def enable_data_parallel(self, dp) -> None:
        self._dp, self.world_size = dp, dp.world_size
        if self._engine is not None:
            dp.attach(self._engine)

    def get_encoder(self) -> nn.Module:
        return self.encoder

    # ---- batch handling (train_monomodal.py:96-135, 196-222) ---------------------------------------------------------------
    @staticmethod
    def _unpack(batch: Dict[Any, Any], config=None):
        want = None
        name = getattr(getattr(config, "experiment", None), "name", "") or ""
        if "AVMNIST_Image_Encoder" in name:
            want = "IMAGE"
        elif "AVMNIST_Audio_Encoder" in name:
            want = "AUDIO"
        key = None
        for k in batch.keys():
            ks = str(k)
            if ks in _SKIP or ks.endswith(("_missing_index", "_reverse", "_original")):
                continue
            key = k
            if want is not None and want in ks.upper():
                break
        if key is None:
            raise ValueError(f"No modality data found in batch. Available keys: {list(batch.keys())}")
        data = batch[f"{key}_original"] if f"{key}_original" in batch else batch[key]
        if isinstance(data, (list, tuple)):
            if len(data) and isinstance(data[0], str):
                raise NotImplementedError("file-path batches are host I/O: load them in the dataset / collate function and pass tensors")
            data = torch.stack([torch.as_tensor(t) for t in data])
        labels = next((batch[k] for k in ("label", "labels", "genres") if k in batch), None)
        if labels is None:
            raise ValueError(f"No labels found in batch. Available keys: {list(batch.keys())}")
        labels = torch.as_tensor(labels)
        return key, data, labels

    def _stage(self, eng: MonoEngine, x: torch.Tensor, labels: Optional[torch.Tensor]) -> _MonoPlan:
        if x.dim() == 4:
            if x.shape[1] != 1:
                raise ValueError("expected a 1-channel tensor")
            x = x[:, 0]
        if x.dim() != 3:
            raise ValueError(f"expected [B,H,W] or [B,1,H,W], got {tuple(x.shape)}")
        plan = eng.plan_for(x.shape[0], x.shape[1], x.shape[2])
        _copy_in(plan.enc.x, x)
        plan.enc.mask.fill_(1.0)
        if labels is not None:
            ops.check_class_labels(labels, plan.logits.shape[1])
            _copy_in(plan.labels, labels.reshape(-1))
        return plan

    # ---- forward / steps --------------------------------------------------------------------------------------------------
    def _stage_vec(self, eng: "VecMonoEngine", x: torch.Tensor, labels: Optional[torch.Tensor]) -> _VecMonoPlan:
        x = x.reshape(x.shape[0], -1)
        plan = eng.plan_for(x.shape[0])
        if x.shape[1] != plan.D:
            raise ValueError(f"expected [B, {plan.D}] features, got {tuple(x.shape)}")
        _copy_in(plan.x, x)
        if labels is not None:
            if labels.dim() != 2 or labels.shape[1] != plan.NC:
                raise ValueError(f"multi-label targets must be [B, {plan.NC}] (bce_with_logits), got {tuple(labels.shape)}")
            _copy_in(plan.labels, labels.to(torch.float32))
        return plan

    def _stage_seq(self, eng: "SeqMonoEngine", x: torch.Tensor, labels: Optional[torch.Tensor]) -> _SeqMonoPlan:
        if x.dim() != 3:
            raise ValueError(f"expected a [B, T, D] sequence batch, got {tuple(x.shape)}")
        plan = eng.plan_for(x.shape[0], x.shape[1])
        if x.shape[2] != plan.D:
            raise ValueError(f"expected {plan.D} features per step, got {x.shape[2]}")
        _copy_in(plan.x, x)
        if labels is not None:
            labels = labels.reshape(-1)
            ops.check_class_labels(labels, plan.NC)
            _copy_in(plan.labels, labels)
        return plan

    def _plan(self, eng, x, labels):
        from .data import note_inputs_consumed

        if self._seq is not None:
            plan = self._stage_seq(eng, x, labels)
        else:
            plan = self._stage_vec(eng, x, labels) if self._vector else self._stage(eng, x, labels)
        note_inputs_consumed(eng.device)  # a prefetcher may overwrite the batch's device buffers from here on
        return plan

    def encode(self, which: str, x: torch.Tensor) -> torch.Tensor:
        """``self.encoder(x)`` outside the step (vector encoders): BatchNorm1d -> Linear, fp32 [B, E]."""
        eng = self._get_engine(x.device if x.is_cuda else next(self.parameters()).device)
        plan = self._stage_vec(eng, x, None)
        plan.run_forward(self.training, with_loss=False)
        return plan.ef.clone()

    def forward(self, x) -> torch.Tensor:
        if isinstance(x, list):
            x = torch.stack(x)
        eng = self._get_engine(x.device if x.is_cuda else next(self.parameters()).device)
        if self._vector or self._seq is not None:
            plan = self._plan(eng, x, None)
            plan.run_forward(self.training, with_loss=False)
            return plan.logits.clone()
        plan = self._stage(eng, x, None)
        plan.run_forward(self.training, with_loss=False)
        return plan.logits.clone()

    def train_step(self, batch, optimizer, loss_functions, device, metric_recorder=None, config=None, **kwargs) -> Dict[str, Any]:
        from .avmnist import AVMNIST

        eng = self._get_engine(device)
        self._check_loss(loss_functions)
        key, x, labels = self._unpack(batch, config)
        self._set_mode(True)
        fs = eng.fs
        fs.adopt_optimizer(optimizer)
        fs.sync_hyper(optimizer, 1.0 / self.world_size)
        plan = self._plan(eng, x, labels)
        given = kwargs.get("dropout_mask") if self._seq == "textcnn" else None
        if given is not None:  # tests: replay a given TextCNN dropout mask
            plan.keep.copy_(torch.as_tensor(given).reshape(plan.keep.shape).to(torch.uint8), non_blocking=True)
            plan.train_step(given_dropout=True)
        else:
            plan.train_step()
        fs._host_step += 1
        return self._finish(eng, plan, key, labels, metric_recorder)

    def validation_step(self, batch, loss_functions, device, metric_recorder=None, config=None, **kwargs) -> Dict[str, Any]:
        from .avmnist import AVMNIST

        eng = self._get_engine(device)
        self._check_loss(loss_functions)
        key, x, labels = self._unpack(batch, config)
        self._set_mode(False)
        plan = self._plan(eng, x, labels)
        plan.run_forward(False, with_loss=True)
        return self._finish(eng, plan, key, labels, metric_recorder)

    def _check_loss(self, loss_functions) -> None:
        """ResNet encoders: one cross_entropy term; MMIMDb encoders: one bce_with_logits term (configs/mmimdb/mono/*.yaml)."""
        if self._vector:
            from .mmimdb import MMIMDb
            MMIMDb._check_loss(loss_functions)
        else:
            from .avmnist import AVMNIST
            AVMNIST._check_loss(loss_functions)

    def _finish(self, eng, plan, key, labels, metric_recorder) -> Dict[str, Any]:
        plan.h_loss.copy_(plan.loss, non_blocking=True)
        plan.h_pred.copy_(plan.pred, non_blocking=True)
        torch.cuda.current_stream(eng.device).synchronize()
        loss = float(plan.h_loss[0])
        if self._vector:  # multi-label: predictions = sigmoid(logits) > 0.5, no accuracy entry (train_monomodal.py:240-258)
            preds = plan.h_pred.to(torch.bool)
            targets = labels.detach().cpu()
            if metric_recorder is not None:
                for group_name in metric_recorder.config.groups:
                    metric_recorder.update_group(group_name=group_name, predictions=preds, targets=targets, modality=str(key))
            return {"loss": loss, "metrics": {"loss": loss}}
        preds = plan.h_pred.to(torch.int64)
        targets = labels.detach().cpu().reshape(-1)
        acc = float((preds == targets).float().mean())
        if metric_recorder is not None:
            for group_name in metric_recorder.config.groups:
                metric_recorder.update_group(group_name=group_name, predictions=preds, targets=targets, modality=str(key))
        return {"loss": loss, "metrics": {"loss": loss, "accuracy": acc}}
