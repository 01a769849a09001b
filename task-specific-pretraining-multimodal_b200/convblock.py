"""AVMNIST's ConvBlock encoders on a B200: ``ConvBlockArgs`` / ``ConvBlock`` (MML_Suite/models/conv.py:7-59), ``MNISTAudio`` /
``MNISTImage`` (MML_Suite/models/avmnist.py:34-185) and the fused step for ``AVMNIST(MNISTAudio, MNISTImage, hidden_dim)`` --
the model of configs/avmnist/centralised/train_avmnist.yaml.

Same constructors, sub-module names and ``state_dict()`` entries as the reference (``net.0.conv_one.weight`` ..
``net.5.bias``).  The arithmetic runs in libmml_b200.so:

  * the first convolution of each encoder (1 input channel, + the missing-modality mask) is a SIMT kernel writing NHWC bf16
    padded to 64 channels (csrc/convblock.cu); every other convolution is the 64-channel tcgen05 implicit GEMM of the ResNet
    path -- 32-channel layers are stored zero-padded to 64 (``FlatState(pad=...)``), which keeps padded activations, gradients
    and Adam moments exactly zero;
  * Conv2d biases sit in front of a BatchNorm2d: they cancel in train mode and are folded into the running mean / the eval
    coefficients (``mml_bn_conv_bias_fold``); their gradient is identically zero and is stored as such (the reference's is
    fp32 rounding noise around 1e-9), so Adam moves them through weight decay only, like the reference;
  * MaxPool2d(k) writes the fp32 ``nn.Flatten`` layout directly for the last pool; Flatten->Linear, the concat head and the
    cross entropy are the small dense kernels.
There is no CPU / PyTorch fallback.
"""
from __future__ import annotations

import os
import weakref
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import ops
from .engine import BN_EPS, BN_MOMENTUM, EncoderPlan, FlatState, _StepPlan

CP = 64  # padded channel count of every activation inside a ConvBlock encoder


# =====================================================================================================================
# modules (parameter containers with the reference's names)
# =====================================================================================================================
@dataclass
class ConvBlockArgs:
    conv_one_in: int
    conv_one_out: int
    conv_one_kernel_size: Union[int, Tuple[int, int]] = (3, 3)
    conv_one_stride: Union[int, Tuple[int, int]] = (1, 1)
    conv_one_padding: Union[int, Tuple[int, int]] = (1, 1)


def _pair(v) -> Tuple[int, int]:
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


class ConvBlock(nn.Module):
    """conv -> BN -> ReLU -> conv -> BN -> ReLU (conv.py:16-59).  Holds parameters; the arithmetic is the encoder plan's."""

    def __init__(self, conv_block_one_args: ConvBlockArgs, conv_block_two_args: ConvBlockArgs, batch_norm: bool = True) -> None:
        super().__init__()
        if not batch_norm:
            raise NotImplementedError("mml_b200 ConvBlock implements the batch_norm=True configuration of the reference's YAMLs")
        for a in (conv_block_one_args, conv_block_two_args):
            if _pair(a.conv_one_kernel_size) != (3, 3) or _pair(a.conv_one_stride) != (1, 1) or _pair(a.conv_one_padding) != (1, 1):
                raise NotImplementedError("mml_b200 ConvBlock implements 3x3 / stride 1 / padding 1 convolutions (the reference's YAMLs)")
            if a.conv_one_out > CP or (a.conv_one_in != 1 and a.conv_one_in > CP):
                raise NotImplementedError(f"mml_b200 ConvBlock supports at most {CP} channels per layer")
        a, b = conv_block_one_args, conv_block_two_args
        self.conv_one = nn.Conv2d(a.conv_one_in, a.conv_one_out, kernel_size=3, stride=1, padding=1)
        self.conv_two = nn.Conv2d(b.conv_one_in, b.conv_one_out, kernel_size=3, stride=1, padding=1)
        self.relu = nn.ReLU()
        self.do_batch_norm = batch_norm
        self.batch_norm_one = nn.BatchNorm2d(a.conv_one_out)
        self.batch_norm_two = nn.BatchNorm2d(b.conv_one_out)

    def forward(self, tensor):
        raise NotImplementedError("a ConvBlock runs inside MNISTAudio / MNISTImage (fused encoder schedule); it has no stand-alone forward")


class _ConvBlockEncoder(nn.Module):
    FLAT: int = 0
    INPUT_HW: Tuple[int, int] = (0, 0)

    def _make(self, args, hidden_dim: int, conv_batch_norm: bool, pools) -> None:
        one = ConvBlock(args[0], args[1], batch_norm=conv_batch_norm)
        two = ConvBlock(args[2], args[3], batch_norm=conv_batch_norm)
        if args[0].conv_one_in != 1 or args[0].conv_one_out not in (8, 16, 32, 64):
            raise NotImplementedError("the first convolution must map 1 channel to 8 / 16 / 32 / 64 channels")
        if args[3].conv_one_out != CP:
            raise NotImplementedError(f"the last convolution must have {CP} output channels (the Flatten layout is written without padding)")
        self.pool_k = []
        for p in pools:
            kh, kw = _pair(p)
            if kh != kw:
                raise NotImplementedError("square MaxPool2d kernels only")
            self.pool_k.append(kh)
        self.hidden_dim = hidden_dim
        self.net = nn.Sequential(one, nn.MaxPool2d(kernel_size=pools[0]), two, nn.MaxPool2d(kernel_size=pools[1]), nn.Flatten(),
                                 nn.Linear(self.FLAT, hidden_dim))
        self._standalone = None

    def get_embedding_size(self) -> int:
        return self.hidden_dim

    def _encode(self, x: torch.Tensor) -> torch.Tensor:
        """Encoder-only forward ([B, hidden_dim] fp32, no autograd graph): train() -> batch statistics + running-stat update."""
        if x.dim() == 4:
            if x.shape[1] != 1:
                raise ValueError("expected a 1-channel tensor")
            x = x[:, 0]
        if x.dim() != 3:
            raise ValueError(f"expected [B,H,W] or [B,1,H,W], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("mml_b200 ConvBlock encoders run on a B200 GPU only: there is no CPU / PyTorch fallback path")
        if self._standalone is None:
            self._standalone = StandaloneConvBlockEncoder(self)
        return self._standalone.forward(x.float().contiguous(), self.training)

    def __str__(self) -> str:
        return str(self.net)


class MNISTAudio(_ConvBlockEncoder):
    """avmnist.py:34-118 (Flatten width 4800 = 64 x 5 x 15 for 32 x 94 spectrograms, pools 2 and 3)."""
    FLAT = 4800

    def __init__(self, conv_block_one_one_args: ConvBlockArgs, conv_block_one_two_args: ConvBlockArgs, conv_block_two_one_args: ConvBlockArgs,
                 conv_block_two_two_args: ConvBlockArgs, hidden_dim: int, *, conv_batch_norm: bool = True,
                 max_pool_one_kernel_size: Union[int, Tuple[int, int]] = (2, 2),
                 max_pool_two_kernel_size: Union[int, Tuple[int, int]] = (3, 3)) -> None:
        super().__init__()
        self._make((conv_block_one_one_args, conv_block_one_two_args, conv_block_two_one_args, conv_block_two_two_args), hidden_dim,
                   conv_batch_norm, (max_pool_one_kernel_size, max_pool_two_kernel_size))

    def forward(self, audio: torch.Tensor) -> torch.Tensor:
        return self._encode(audio)  # the reference unsqueezes the channel dimension here (avmnist.py:117)


class MNISTImage(_ConvBlockEncoder):
    """avmnist.py:121-185 (Flatten width 3136 = 64 x 7 x 7 for 28 x 28 images, both pools 2)."""
    FLAT = 3136

    def __init__(self, conv_block_one_one_args: ConvBlockArgs, conv_block_one_two_args: ConvBlockArgs, conv_block_two_one_args: ConvBlockArgs,
                 conv_block_two_two_args: ConvBlockArgs, hidden_dim: int, *, conv_batch_norm: bool = True,
                 max_pool_kernel_size: Union[int, Tuple[int, int]] = (2, 2)) -> None:
        super().__init__()
        self._make((conv_block_one_one_args, conv_block_one_two_args, conv_block_two_one_args, conv_block_two_two_args), hidden_dim,
                   conv_batch_norm, (max_pool_kernel_size, max_pool_kernel_size))

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        return self._encode(image)


def pad_map(enc: nn.Module, prefix: str) -> Dict[str, int]:
    """FlatState ``pad`` entries of one ConvBlock encoder: every conv weight with more than one input channel and every
    BatchNorm vector / running statistic is stored with 64 channels."""
    out: Dict[str, int] = {}
    for slot in (0, 2):
        blk = enc.net[slot]
        for conv, bn in (("conv_one", "batch_norm_one"), ("conv_two", "batch_norm_two")):
            if getattr(blk, conv).in_channels != 1:
                out[f"{prefix}net.{slot}.{conv}.weight"] = CP
            for leaf in ("weight", "bias", "running_mean", "running_var"):
                out[f"{prefix}net.{slot}.{bn}.{leaf}"] = CP
    return out


# =====================================================================================================================
# per-encoder plan
# =====================================================================================================================
class ConvBlockEncoderPlan(EncoderPlan):
    """Buffers + forward / backward closures of one MNISTAudio / MNISTImage for a fixed (B, H, W).  The embedding is written to
    ``emb`` ([B, ld] fp32 at column ``emb_off``: the concat of the fusion head is a column offset) and its gradient read from
    ``demb`` at the same place."""

    def __init__(self, fs: FlatState, enc: nn.Module, prefix: str, B: int, H: int, W: int, train: bool, emb: torch.Tensor, demb: torch.Tensor,
                 emb_off: int):
        self.fs, self.enc, self.prefix, self.B, self.H, self.W = fs, enc, prefix, B, H, W
        dev = fs.device
        self.fwd_train: List[Callable[[], None]] = []
        self.fwd_eval: List[Callable[[], None]] = []
        self.bwd: List[Callable[[], None]] = []
        self.x = torch.zeros(B, H, W, device=dev)
        self.mask = torch.ones(B, device=dev)
        self.taps: Dict[str, torch.Tensor] = {}
        self.wgrad_stream: Optional[torch.cuda.Stream] = None
        self.wgrad_ws = ops.WgradScratch(dev)
        self.stat_arena = torch.zeros(4 * 4 * ops.bn_stat_slots(CP) * CP, device=dev, dtype=torch.float64)  # 4 BatchNorms x (fwd, bwd)
        self._stat_off = 0
        self.emb, self.demb, self.emb_off = emb, demb, emb_off
        k1, k2 = enc.pool_k
        if (H // k1) // k2 < 1 or (W // k1) // k2 < 1 or CP * ((H // k1) // k2) * ((W // k1) // k2) != enc.FLAT:
            raise ValueError(f"{type(enc).__name__}: a {H}x{W} input does not flatten to {enc.FLAT} features (avmnist.py: conv_block_out_dim)")
        self._build_convblock(train)

    def _build_convblock(self, train: bool) -> None:
        fs, B, dev, pre, enc = self.fs, self.B, self.fs.device, self.prefix, self.enc
        F, E, Bk = self.fwd_train, self.fwd_eval, self.bwd
        x, mask = self.x, self.mask
        k1, k2 = enc.pool_k
        H0, W0 = self.H, self.W
        H1, W1 = H0 // k1, W0 // k1
        H2, W2 = H1 // k2, W1 // k2
        K1 = enc.net[0].conv_one.out_channels
        P_, G_, Wb = fs.P, fs.G, fs.Wb

        def names(slot, conv, bn):
            return f"{pre}net.{slot}.{conv}", f"net.{slot}.{bn}"

        # ---- buffers
        act = self._act
        raw1, a1, raw2, a2 = (act(B, H0, W0, CP) for _ in range(4))
        pool1 = act(B, H1, W1, CP)
        amax1 = torch.zeros(B, H1, W1, CP, device=dev, dtype=torch.uint8)
        raw3, a3, raw4, a4 = (act(B, H1, W1, CP) for _ in range(4))
        amax2 = torch.zeros(B, H2, W2, CP, device=dev, dtype=torch.uint8)
        flat = torch.zeros(B, enc.FLAT, device=dev)
        self.flat = flat
        self.taps.update({"net.0.conv_one": raw1, "net.0.relu_one": a1, "net.0.conv_two": raw2, "net.0": a2, "net.1": pool1,
                          "net.2.conv_one": raw3, "net.2.relu_one": a3, "net.2.conv_two": raw4, "net.2": a4})
        c1, b1n = names(0, "conv_one", "batch_norm_one")
        c2, b2n = names(0, "conv_two", "batch_norm_two")
        c3, b3n = names(2, "conv_one", "batch_norm_one")
        c4, b4n = names(2, "conv_two", "batch_norm_two")
        bn1, bn2, bn3, bn4 = (self._bn(n, CP) for n in (b1n, b2n, b3n, b4n))
        w1 = fs.flat_slice(P_, c1 + ".weight")                      # fp32 [K1][9]
        w2, w3, w4 = (fs.flat_slice(Wb, c + ".weight") for c in (c2, c3, c4))  # bf16 [64][3][3][64]
        bias = [fs.flat_slice(P_, c + ".bias") for c in (c1, c2, c3, c4)]
        g_a = ops.make_geom(B, H0, W0, CP, CP, 3, 3, 1, 1)
        g_b = ops.make_geom(B, H1, W1, CP, CP, 3, 3, 1, 1)
        rows_a, rows_b = B * H0 * W0, B * H1 * W1
        fc_w = fs.flat_slice(P_, pre + "net.5.weight").view(enc.hidden_dim, enc.FLAT)
        fc_b = fs.flat_slice(P_, pre + "net.5.bias")
        emb_view = self.emb[:, self.emb_off:]
        ld_e = self.emb.shape[1]

        def bn_relu_train(raw, bn, out, rows, cb):
            return [lambda: ops.bn_train_fwd(raw, bn, None, None, out, rows, CP, True, BN_MOMENTUM, BN_EPS),
                    lambda: ops.bn_conv_bias_fold(cb, BN_MOMENTUM, running_mean=bn.rmean)]

        def bn_relu_eval(raw, bn, out, rows, cb):
            return [lambda: ops.bn_eval_coeffs(CP, bn.gamma, bn.beta, bn.rmean, bn.rvar, BN_EPS, bn.scale, bn.shift),
                    lambda: ops.bn_conv_bias_fold(cb, 0.0, scale=bn.scale, shift=bn.shift),
                    lambda: ops.bn_act_fwd(raw, bn.scale, bn.shift, None, None, None, out, rows, CP, True)]

        for L, train_mode in ((F, True), (E, False)):
            tail = bn_relu_train if train_mode else bn_relu_eval
            st = (lambda bn: bn.stats) if train_mode else (lambda bn: None)
            L.append(lambda s=st(bn1): ops.conv3x3_c1_fprop(x, mask, w1, raw1, s, K1))
            L.extend(tail(raw1, bn1, a1, rows_a, bias[0]))
            L.append(lambda s=st(bn2): ops.conv_fprop(g_a, a1, w2, raw2, s))
            L.extend(tail(raw2, bn2, a2, rows_a, bias[1]))
            L.append(lambda: ops.maxpool_k_fwd(a2, pool1, None, amax1, k1))
            L.append(lambda s=st(bn3): ops.conv_fprop(g_b, pool1, w3, raw3, s))
            L.extend(tail(raw3, bn3, a3, rows_b, bias[2]))
            L.append(lambda s=st(bn4): ops.conv_fprop(g_b, a3, w4, raw4, s))
            L.extend(tail(raw4, bn4, a4, rows_b, bias[3]))
            L.append(lambda: ops.maxpool_k_fwd(a4, None, flat, amax2, k2))
            L.append(lambda: ops.dense_fwd(flat, enc.FLAT, fc_w, fc_b, None, 1.0, False, emb_view, ld_e, B))
        if not train:
            return
        # ---- backward
        d_flat = torch.zeros(B, enc.FLAT, device=dev)
        d_a4, d_raw4, d_a3, d_raw3, d_pool1 = (act(B, H1, W1, CP) for _ in range(5))
        d_a2, d_raw2, d_a1, d_raw1 = (act(B, H0, W0, CP) for _ in range(4))
        dw1 = fs.flat_slice(G_, c1 + ".weight")
        dw2, dw3, dw4 = (fs.flat_slice(G_, c + ".weight") for c in (c2, c3, c4))
        d_fc_w = fs.flat_slice(G_, pre + "net.5.weight").view(enc.hidden_dim, enc.FLAT)
        d_fc_b = fs.flat_slice(G_, pre + "net.5.bias")
        demb_view = self.demb[:, self.emb_off:]
        ws1 = torch.zeros(max(ops.conv3x3_c1_wgrad_workspace(x, K1) // 4, 4), device=dev)

        def bn_relu_bwd(dy, out, raw, bn, d_raw, rows):
            ops.bn_bwd_reduce(dy, None, out, raw, bn.mean, bn.invstd, bn.bstat, dy, rows, CP, True)  # g overwrites dy in place
            ops.bn_bwd_apply(dy, raw, bn.mean, bn.invstd, bn.gamma, bn.bstat, bn.dgamma, bn.dbeta, d_raw, rows, CP)

        def bwd_fc():
            ops.dense_bwd(demb_view, emb_view, ld_e, None, 1.0, False, flat, enc.FLAT, fc_w, d_flat, enc.FLAT, d_fc_w, d_fc_b, B, lddy=ld_e)
            ops.maxpool_k_bwd(None, d_flat, amax2, d_a4, k2)

        def bwd_block_two():
            bn_relu_bwd(d_a4, a4, raw4, bn4, d_raw4, rows_b)
            self._offload(lambda: ops.conv_wgrad(g_b, a3, d_raw4, dw4, self.wgrad_ws))
            ops.conv_dgrad(g_b, d_raw4, w4, d_a3)
            bn_relu_bwd(d_a3, a3, raw3, bn3, d_raw3, rows_b)
            self._offload(lambda: ops.conv_wgrad(g_b, pool1, d_raw3, dw3, self.wgrad_ws))
            ops.conv_dgrad(g_b, d_raw3, w3, d_pool1)
            ops.maxpool_k_bwd(d_pool1, None, amax1, d_a2, k1)

        def bwd_block_one():
            bn_relu_bwd(d_a2, a2, raw2, bn2, d_raw2, rows_a)
            self._offload(lambda: ops.conv_wgrad(g_a, a1, d_raw2, dw2, self.wgrad_ws))
            ops.conv_dgrad(g_a, d_raw2, w2, d_a1)
            bn_relu_bwd(d_a1, a1, raw1, bn1, d_raw1, rows_a)
            ops.conv3x3_c1_wgrad(x, mask, d_raw1, dw1, ws1, K1)
            self.join_offload()

        Bk.extend([bwd_fc, bwd_block_two, bwd_block_one])
        self.bwd_names = ["fc", "net.2", "net.0"]
        self.grad_taps = {"net.0.conv_one": d_raw1, "net.0.conv_two": d_raw2, "net.2.conv_one": d_raw3, "net.2.conv_two": d_raw4}


class StandaloneConvBlockEncoder:
    """MNISTAudio.forward / MNISTImage.forward outside the fusion model (``get_embeddings``, avmnist.py:362-401).  Shares the
    owning fusion engine's storage when there is one."""

    def __init__(self, enc: nn.Module):
        self.enc = enc
        self.fs: Optional[FlatState] = None
        self.plans: Dict[Tuple, ConvBlockEncoderPlan] = {}

    def forward(self, x: torch.Tensor, training: bool) -> torch.Tensor:
        owner = getattr(self.enc, "_mml_owner", None)
        eng = owner[0]() if owner is not None else None
        if eng is not None and eng.device == x.device:
            fs, prefix = eng.fs, owner[1]
        else:
            if self.fs is None or self.fs.device != x.device:
                self.fs = FlatState(self.enc, x.device, pad=pad_map(self.enc, ""))
                self.plans.clear()
            fs, prefix = self.fs, ""
        fs.ensure_fresh()
        key = (id(fs),) + tuple(x.shape)
        plan = self.plans.get(key)
        if plan is None:
            emb = torch.zeros(x.shape[0], self.enc.hidden_dim, device=x.device)
            plan = self.plans[key] = ConvBlockEncoderPlan(fs, self.enc, prefix, x.shape[0], x.shape[1], x.shape[2], False, emb, emb, 0)
        plan.x.copy_(x)
        plan.mask.fill_(1.0)
        if training:
            plan.stat_arena.zero_()
        for op in (plan.fwd_train if training else plan.fwd_eval):
            op()
        if training:
            for i, name in enumerate(fs.nbt_names):
                if name.startswith(prefix):
                    fs.NBT[i] += 1
        return plan.emb.clone()


# =====================================================================================================================
# the fused step of AVMNIST(MNISTAudio, MNISTImage, hidden_dim)
# =====================================================================================================================
class ConvBlockFusionEngine:
    """Same surface as engine.LateFusionEngine (``fs``, ``plan_for``, ``device``, data-parallel hooks)."""

    def __init__(self, model: nn.Module, device: torch.device, dropout_p: float, seed: Optional[int] = None):
        self.model, self.device, self.dropout_p = model, device, float(dropout_p)
        self.client_id = int(getattr(model, "_mml_client_id", 0))
        self.seed = ops.engine_seed(self.client_id) if seed is None else seed
        self.fwd_calls = 0
        pad = {}
        pad.update(pad_map(model.audio_encoder, "audio_encoder."))
        pad.update(pad_map(model.image_encoder, "image_encoder."))
        self.fs = FlatState(model, device, pad=pad)
        self.plans: Dict[Tuple, "_ConvBlockStepPlan"] = {}
        self.world = 1
        self.allreduce = None
        self.allreduce_range = None
        self.use_graphs = True

    def plan_for(self, B: int, aH: int, aW: int, iH: int, iW: int) -> "_ConvBlockStepPlan":
        key = (B, aH, aW, iH, iW)
        plan = self.plans.get(key)
        if plan is None:
            plan = self.plans[key] = _ConvBlockStepPlan(self, B, aH, aW, iH, iW)
        return plan


class _ConvBlockStepPlan(_StepPlan):
    """mask -> both encoders (audio on the current stream, image on a side stream) -> concat -> net.0/3/5 -> CE -> backward ->
    [all-reduce] -> Adam.  Reuses _StepPlan's stream fork / join, eager-then-graph execution and Adam ranges."""

    def __init__(self, eng: ConvBlockFusionEngine, B: int, aH: int, aW: int, iH: int, iW: int):
        self.eng, self.B = eng, B
        fs, dev, model = eng.fs, eng.device, eng.model
        EA, EI = model.audio_encoder.hidden_dim, model.image_encoder.hidden_dim
        self.EA, self.EI = EA, EI
        self.concat = torch.zeros(B, EA + EI, device=dev)
        self.dconcat = torch.zeros(B, EA + EI, device=dev)
        self.audio = ConvBlockEncoderPlan(fs, model.audio_encoder, "audio_encoder.", B, aH, aW, True, self.concat, self.dconcat, 0)
        self.image = ConvBlockEncoderPlan(fs, model.image_encoder, "image_encoder.", B, iH, iW, True, self.concat, self.dconcat, EA)
        params = dict(model.named_parameters())
        self.W = {n: fs.flat_slice(fs.P, n).view(params[n].shape) for n in ("net.0.weight", "net.3.weight", "net.5.weight")}
        self.Bv = {n: fs.flat_slice(fs.P, n) for n in ("net.0.bias", "net.3.bias", "net.5.bias")}
        self.dW = {n: fs.flat_slice(fs.G, n).view(params[n].shape) for n in self.W}
        self.dB = {n: fs.flat_slice(fs.G, n) for n in self.Bv}
        self.H1, self.H2, self.NC = (params[n].shape[0] for n in ("net.0.weight", "net.3.weight", "net.5.weight"))
        z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
        self.h1, self.h2, self.dh1, self.dh2 = z(B, self.H1), z(B, self.H2), z(B, self.H1), z(B, self.H2)
        self.labels = torch.zeros(B, device=dev, dtype=torch.int64)
        self.logits, self.dlogits, self.row_loss = z(B, self.NC), z(B, self.NC), z(B)
        self.loss = z(1)
        self.pred = torch.zeros(B, device=dev, dtype=torch.int32)
        self.drop_mask = torch.ones(B, self.H1, device=dev, dtype=torch.uint8)
        self.h_loss = torch.zeros(1).pin_memory()
        self.h_pred = torch.zeros(B, dtype=torch.int32).pin_memory()
        self.graph_train = self.graph_train_nodrop = self.graph_eval = None
        self.loss_ready = torch.cuda.Event()
        self.want_pred = False
        self.eager_steps = 0
        self.launches_per_step = 0
        self.side_stream = None
        self.mid_stream = None
        self.tune = {"adam_split": False, "head_side": False, "skip": "", "side_prio": os.environ.get("MML_SIDE_PRIO", "-1")}
        self.reserve_sms = 0  # both encoders launch full grids here; nothing to reserve
        self.pdl_mode = os.environ.get("MML_PDL_MODE", "none")
        ops.set_pdl(dev.index, self.pdl_mode != "none")
        if os.environ.get("MML_WGRAD_STREAMS", "1") == "1":
            self.audio.wgrad_stream = torch.cuda.Stream(device=dev)
            self.image.wgrad_stream = torch.cuda.Stream(device=dev)
        self.param_split = 0
        self.audio_mid = 0

    # -- head ------------------------------------------------------------------------------------------------------
    def _head_fwd(self, dm, scale: float) -> None:
        B, E = self.B, self.EA + self.EI
        ops.dense_fwd(self.concat, E, self.W["net.0.weight"], self.Bv["net.0.bias"], dm, scale, True, self.h1, self.H1, B)
        ops.dense_fwd(self.h1, self.H1, self.W["net.3.weight"], self.Bv["net.3.bias"], None, 1.0, True, self.h2, self.H2, B)
        ops.dense_fwd(self.h2, self.H2, self.W["net.5.weight"], self.Bv["net.5.bias"], None, 1.0, False, self.logits, self.NC, B)

    def _head_bwd(self, dm, scale: float) -> None:
        B, E = self.B, self.EA + self.EI
        ops.dense_bwd(self.dlogits, self.logits, self.NC, None, 1.0, False, self.h2, self.H2, self.W["net.5.weight"], self.dh2, self.H2,
                      self.dW["net.5.weight"], self.dB["net.5.bias"], B)
        ops.dense_bwd(self.dh2, self.h2, self.H2, None, 1.0, True, self.h1, self.H1, self.W["net.3.weight"], self.dh1, self.H1,
                      self.dW["net.3.weight"], self.dB["net.3.bias"], B)
        ops.dense_bwd(self.dh1, self.h1, self.H1, dm, scale, True, self.concat, E, self.W["net.0.weight"], self.dconcat, E,
                      self.dW["net.0.weight"], self.dB["net.0.bias"], B)

    def _drop(self):
        p = self.eng.dropout_p
        return (self.drop_mask, 1.0 / (1.0 - p)) if p > 0.0 else (None, 1.0)

    # -- schedules -------------------------------------------------------------------------------------------------
    def run_train_fwd(self, own_dropout: bool) -> None:
        eng, fs = self.eng, self.eng.fs
        self.audio.stat_arena.zero_()
        self.image.stat_arena.zero_()
        if self._use_dropout() and own_dropout:
            ops.dropout_mask(self.drop_mask, eng.dropout_p, eng.seed, fs.step)
        self._both_encoders(self.audio.fwd_train, self.image.fwd_train)
        dm, scale = self._drop()
        self._head_fwd(dm, scale)
        ops.softmax_ce(self.logits, self.labels, self.dlogits, self.row_loss, self.loss, self.pred)

    def run_train_bwd(self) -> None:
        dm, scale = self._drop()
        self._head_bwd(dm, scale)
        self._both_encoders(self.audio.bwd, self.image.bwd)
        self.eng.fs.NBT += 1

    def run_update(self) -> None:
        eng, fs = self.eng, self.eng.fs
        if eng.allreduce_range is not None:
            eng.allreduce_range(0, fs.total, [torch.cuda.current_stream(eng.device)], update=lambda: self._adam_range(0, fs.total, True), join=True)
        else:
            self._adam_range(0, fs.total, True)

    def run_eval(self, with_loss: bool) -> None:
        self._both_encoders(self.audio.fwd_eval, self.image.fwd_eval)
        self._head_fwd(None, 1.0)
        ops.softmax_ce(self.logits, self.labels, self.dlogits, self.row_loss, self.loss, self.pred)

    def run_forward_train_mode(self) -> None:
        eng, fs = self.eng, self.eng.fs
        self.audio.stat_arena.zero_()
        self.image.stat_arena.zero_()
        if self._use_dropout():
            eng.fwd_calls += 1
            ops.dropout_mask(self.drop_mask, eng.dropout_p, ops.engine_seed(eng.client_id, eng.fwd_calls) ^ eng.seed, fs.step)
        self._both_encoders(self.audio.fwd_train, self.image.fwd_train)
        dm, scale = self._drop()
        self._head_fwd(dm, scale)
        fs.NBT += 1


def make_engine(model: nn.Module, device: torch.device, dropout_p: float) -> ConvBlockFusionEngine:
    eng = ConvBlockFusionEngine(model, device, dropout_p)
    model.audio_encoder._mml_owner = (weakref.ref(eng), "audio_encoder.")
    model.image_encoder._mml_owner = (weakref.ref(eng), "image_encoder.")
    return eng
