"""MOSI / UttFusion model -- drop-in for ``MML_Suite/models/msa/utt_fusion.py`` (+ ``networks/lstm.py``, ``networks/textcnn.py``,
``networks/classifier.py``) on a B200.

Same classes and constructors as the YAML tags build them (configs/mosi/centralised/utt_fusion_base_training.yaml:10-46):
``LSTMEncoder(input_size, hidden_size, embd_method="last")``, ``TextCNN(input_size, embd_size, in_channels, out_channels,
kernel_heights, dropout)``, ``FcClassifier(input_dim, layers, output_dim, dropout=, use_bn=False)`` and
``UttFusionModel(netA, netV, netT, netC, clip=, pretrained_path=)``; same sub-module names, hence the same 24-entry
``state_dict()``; same ``forward(A, V, T)`` / ``train_step`` / ``validation_step`` / ``get_encoder`` / ``flatten_parameters``.
The torch modules inside are parameter CONTAINERS (reference initialisation and names); the arithmetic of a step is one fused
schedule in libmml_b200.so: LSTM forward / BPTT kernels, the TextCNN convolutions on the tcgen05 conv path (the Conv2d weight
[128,1,k,768] is a K,R,S,C tensor with C = 768), fused ReLU + max-over-time, small dense kernels, softmax-CE, the gradient-norm
clip folded into the fused Adam.  Unsupported requests (attention / maxpool LSTM embeddings, batch-norm classifier, embeddings
as inputs, other optimizers or losses) raise -- nothing falls back to PyTorch.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .avmnist import AVMNIST, _copy_in, _find
from .engine import BF16, FlatState

_FUSED_ONLY = "is evaluated inside the fused UttFusionModel step only"


class LSTMEncoder(nn.Module):
    """lstm.py:8-64 (container).  Only ``embd_method="last"`` (h_T, the YAML's choice) is built."""

    def __init__(self, input_size: int, hidden_size: int, embd_method: str = "last"):
        super().__init__()
        if embd_method != "last":
            raise NotImplementedError(f"mml_b200 LSTMEncoder implements embd_method='last' (got '{embd_method}')")
        if hidden_size != 64 or not 1 <= input_size <= 32:
            raise NotImplementedError("mml_b200 LSTM kernels are built for hidden_size 64 and 1..32 input features")
        self.input_size, self.hidden_size, self.embd_method = input_size, hidden_size, embd_method
        self.rnn = nn.LSTM(input_size, hidden_size, batch_first=True)

    def forward(self, x):
        raise NotImplementedError("mml_b200.LSTMEncoder " + _FUSED_ONLY)


class TextCNN(nn.Module):
    """textcnn.py:10-69 (container)."""

    def __init__(self, input_size: int, embd_size: int = 128, in_channels: int = 1, out_channels: int = 128,
                 kernel_heights: List[int] = [3, 4, 5], dropout: float = 0.5) -> None:
        super().__init__()
        if in_channels != 1 or len(kernel_heights) != 3 or input_size % 64 or out_channels % 64 or max(kernel_heights) > 9:
            raise NotImplementedError("mml_b200 TextCNN: in_channels 1, three kernel heights <= 9, feature / channel widths multiples of 64")
        self.conv1 = nn.Conv2d(in_channels, out_channels, (kernel_heights[0], input_size), stride=1, padding=0)
        self.conv2 = nn.Conv2d(in_channels, out_channels, (kernel_heights[1], input_size), stride=1, padding=0)
        self.conv3 = nn.Conv2d(in_channels, out_channels, (kernel_heights[2], input_size), stride=1, padding=0)
        self.dropout = nn.Dropout(dropout)
        self.embd = nn.Sequential(nn.Linear(len(kernel_heights) * out_channels, embd_size), nn.ReLU(inplace=True))
        self.hidden_size = embd_size
        self.kernel_heights, self.input_size, self.out_channels = list(kernel_heights), input_size, out_channels

    def forward(self, frame_x):
        raise NotImplementedError("mml_b200.TextCNN " + _FUSED_ONLY)


class FcClassifier(nn.Module):
    """classifier.py:83-117 (container)."""

    def __init__(self, input_dim: int, layers: List[int], output_dim: int, *, dropout: float = 0.3, use_bn: bool = False) -> None:
        super().__init__()
        if use_bn or len(layers) == 0:
            raise NotImplementedError("mml_b200 FcClassifier: use_bn=False and at least one hidden layer")
        mods: List[nn.Module] = []
        self.layer_index: List[int] = []
        d = input_dim
        for width in layers:
            self.layer_index.append(len(mods))
            mods.append(nn.Linear(d, width))
            mods.append(nn.ReLU())
            if dropout > 0:
                mods.append(nn.Dropout(dropout))
            d = width
        self.module = nn.Sequential(*mods)
        self.fc_out = nn.Linear(layers[-1], output_dim)
        self.dropout_p = float(dropout)

    def forward(self, x):
        raise NotImplementedError("mml_b200.FcClassifier " + _FUSED_ONLY)


class _UttPlan:
    """Static buffers + schedule of one (batch, sequence length)."""

    def __init__(self, eng: "UttEngine", B: int, T: int):
        self.eng, self.B, self.T = eng, B, T
        fs, dev, m = eng.fs, eng.device, eng.model
        params = dict(m.named_parameters())

        def par(flat, name):
            return fs.flat_slice(flat, name).view(params[name].shape)

        f32 = lambda *s: torch.zeros(*s, device=dev)
        H = 64
        self.H = H
        self.DA, self.DV, self.DT = m.netA.input_size, m.netV.input_size, m.netT.input_size
        self.xA, self.xV, self.xT = f32(B, T, self.DA), f32(B, T, self.DV), f32(B, T, self.DT)
        self.mA, self.mV, self.mT = (torch.ones(B, device=dev) for _ in range(3))
        self.xmA, self.xmV = f32(B, T, self.DA), f32(B, T, self.DV)      # masked inputs (x * m, base_dataset.py:71)
        self.xmT = f32(B, T, self.DT)
        self.xT16 = torch.zeros(B, T, self.DT, device=dev, dtype=BF16)   # masked text, bf16 NHWC [B][T][1][768]
        self.labels = torch.zeros(B, device=dev, dtype=torch.int64)
        # LSTMs
        self.lstm = {}
        for key, net, x in (("A", "netA", self.xmA), ("V", "netV", self.xmV)):
            names = [f"{net}.rnn.{n}" for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
            self.lstm[key] = dict(x=x, w=[par(fs.P, n) for n in names], dw=[par(fs.G, n) for n in names],
                                  gates=f32(B, T, 4 * H), cs=f32(B, T, H), hs=f32(B, T, H), h_last=f32(B, H), dh=f32(B, H))
        # TextCNN
        C_ = m.netT.out_channels
        self.C = C_
        self.convs = []
        self.wgrad_ws = ops.WgradScratch(dev)  # the three TextCNN weight gradients run on one stream
        for i, k in enumerate(m.netT.kernel_heights):
            name = f"netT.conv{i + 1}"
            P_ = T - k + 1
            if P_ < 1:
                raise ValueError(f"sequence length {T} shorter than the TextCNN kernel height {k}")
            self.convs.append(dict(k=k, P=P_, geom=ops.make_geom(B, T, 1, self.DT, C_, k, 1, 1, 0),
                                   w16=fs.flat_slice(fs.Wb, name + ".weight").view(C_, k, 1, self.DT),
                                   dw=fs.flat_slice(fs.G, name + ".weight").view(C_, k, 1, self.DT),
                                   bias=par(fs.P, name + ".bias"), dbias=par(fs.G, name + ".bias"),
                                   out=torch.zeros(B, P_, C_, device=dev, dtype=BF16), dout=torch.zeros(B, P_, C_, device=dev, dtype=BF16)))
        NP = 3 * C_
        self.pooled, self.dpooled = f32(B, NP), f32(B, NP)
        self.arg = torch.zeros(B, NP, device=dev, dtype=torch.int32)
        self.keepT = torch.ones(B, NP, device=dev, dtype=torch.uint8)
        self.pT = float(m.netT.dropout.p)
        # fused features [a | v | t] and the dense chain (textcnn embd + classifier)
        E = m.netT.hidden_size
        self.fused, self.dfused = f32(B, 2 * H + E), f32(B, 2 * H + E)
        self.demb = f32(B, E)
        if m.netC.module[0].in_features != 2 * H + E:
            raise ValueError("classifier input width does not match the concatenated embeddings")
        self.embd = dict(w=par(fs.P, "netT.embd.0.weight"), b=par(fs.P, "netT.embd.0.bias"), dw=par(fs.G, "netT.embd.0.weight"),
                         db=par(fs.G, "netT.embd.0.bias"))
        self.pC = m.netC.dropout_p
        self.dense = []
        x, ldx = self.fused, self.fused.shape[1]
        for idx in m.netC.layer_index:
            n = f"netC.module.{idx}"
            N = params[n + ".weight"].shape[0]
            layer = dict(w=par(fs.P, n + ".weight"), b=par(fs.P, n + ".bias"), dw=par(fs.G, n + ".weight"), db=par(fs.G, n + ".bias"), x=x, ldx=ldx,
                         y=f32(B, N), dy=f32(B, N), keep=torch.ones(B, N, device=dev, dtype=torch.uint8) if self.pC > 0 else None)
            self.dense.append(layer)
            x, ldx = layer["y"], N
        self.NC = params["netC.fc_out.weight"].shape[0]
        self.out = dict(w=par(fs.P, "netC.fc_out.weight"), b=par(fs.P, "netC.fc_out.bias"), dw=par(fs.G, "netC.fc_out.weight"),
                        db=par(fs.G, "netC.fc_out.bias"), x=x, ldx=ldx)
        self.logits, self.dlogits = f32(B, self.NC), f32(B, self.NC)
        self.row_loss, self.loss = f32(B), f32(1)
        self.pred = torch.zeros(B, device=dev, dtype=torch.int32)
        self.partial = torch.zeros(256, device=dev, dtype=torch.float64)
        self.grad_norm = f32(1)
        self.h_loss = torch.zeros(1).pin_memory()
        self.h_pred = torch.zeros(B, dtype=torch.int32).pin_memory()
        self.graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        self.eager_steps = 0
        self.launches_per_step = 0
        import os as _os
        # the two LSTM recurrences (50 dependent steps each) and the text branch are independent: three streams (MML_UTT_STREAMS=0: one)
        multi = _os.environ.get("MML_UTT_STREAMS", "1") == "1"
        self.streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)] if multi else None

    def _parallel(self, main_ops, *side_chains) -> None:
        """``main_ops`` on the current stream, every chain of ``side_chains`` on its own stream; fork before, join after."""
        if self.streams is None:
            for chain in (main_ops,) + side_chains:
                for op in chain:
                    op()
            return
        main = torch.cuda.current_stream(self.eng.device)
        for st, chain in zip(self.streams, side_chains):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                for op in chain:
                    op()
        for op in main_ops:
            op()
        for st, _ in zip(self.streams, side_chains):
            main.wait_stream(st)

    # ---- schedule --------------------------------------------------------------------------------------------------------
    def _mask_inputs(self) -> None:
        # sample[mod] = original * mask (data/base_dataset.py:70-72): the library's mask kernel, like the AVMNIST / MMIMDb paths
        ops.mask_apply_into(self.xA, self.mA, self.xmA)
        ops.mask_apply_into(self.xV, self.mV, self.xmV)
        ops.mask_apply_into(self.xT, self.mT, self.xmT)
        ops.cast_f32_bf16(self.xmT, self.xT16)

    def run_forward(self, train: bool, with_loss: bool, with_grad: bool) -> None:
        B, H = self.B, self.H
        self._mask_inputs()
        dropT = train and self.pT > 0

        def lstm_chain(off, key):
            L = self.lstm[key]
            return [lambda: ops.lstm_fwd(L["x"], *L["w"], L["gates"], L["cs"], L["hs"], L["h_last"]),
                    lambda: self.fused[:, off:off + H].copy_(L["h_last"])]  # torch.cat([a, v, t]) (utt_fusion.py:147) is a column offset

        def text_chain():
            for i, cv in enumerate(self.convs):
                ops.conv_fprop(cv["geom"], self.xT16, cv["w16"], cv["out"], None)
                ops.relumax_fwd(cv["out"], cv["bias"], self.keepT if dropT else None, 1.0 / (1.0 - self.pT) if dropT else 1.0, self.pooled, self.arg,
                                i * self.C)
            ops.dense_fwd(self.pooled, self.pooled.shape[1], self.embd["w"], self.embd["b"], None, 1.0, True, self.fused[:, 2 * H:], self.fused.shape[1], B)

        self._parallel([text_chain], lstm_chain(0, "A"), lstm_chain(H, "V"))
        dropC = train and self.pC > 0
        for layer in self.dense:
            ops.dense_fwd(layer["x"], layer["ldx"], layer["w"], layer["b"], layer["keep"] if dropC else None, 1.0 / (1.0 - self.pC) if dropC else 1.0, True,
                          layer["y"], layer["y"].shape[1], B)
        o = self.out
        ops.dense_fwd(o["x"], o["ldx"], o["w"], o["b"], None, 1.0, False, self.logits, self.NC, B)
        ops.softmax_ce(self.logits, self.labels if with_loss else None, self.dlogits if with_grad else None, self.row_loss if with_loss else None,
                       self.loss if with_loss else None, self.pred, 1.0)

    def run_train(self, own_dropout: bool) -> None:
        eng, fs, B, H = self.eng, self.eng.fs, self.B, self.H
        fs.G.zero_()
        if own_dropout:
            if self.pT > 0:
                ops.dropout_mask(self.keepT, self.pT, eng.seed, fs.step)
            for i, layer in enumerate(self.dense):
                if layer["keep"] is not None:
                    ops.dropout_mask(layer["keep"], self.pC, eng.seed + 101 * (i + 1), fs.step)
        self.run_forward(True, True, True)
        # classifier backward (reverse), ending in d fused
        o = self.out
        scale_c = 1.0 / (1.0 - self.pC) if self.pC > 0 else 1.0
        last = self.dense[-1]
        ops.dense_bwd(self.dlogits, self.logits, self.NC, None, 1.0, False, o["x"], o["ldx"], o["w"], last["dy"], last["dy"].shape[1], o["dw"], o["db"], B)
        for li in range(len(self.dense) - 1, -1, -1):
            layer = self.dense[li]
            dx, lddx = (self.dense[li - 1]["dy"], self.dense[li - 1]["dy"].shape[1]) if li > 0 else (self.dfused, self.dfused.shape[1])
            ops.dense_bwd(layer["dy"], layer["y"], layer["y"].shape[1], layer["keep"], scale_c, True, layer["x"], layer["ldx"], layer["w"], dx, lddx,
                          layer["dw"], layer["db"], B)
        # text branch: embd Linear+ReLU -> dropout / max over time -> conv weight gradients
        e = self.embd
        dropT = self.pT > 0

        def text_bwd():
            self.demb.copy_(self.dfused[:, 2 * H:])
            ops.dense_bwd(self.demb, self.fused[:, 2 * H:], self.fused.shape[1], None, 1.0, True, self.pooled, self.pooled.shape[1], e["w"], self.dpooled,
                          self.dpooled.shape[1], e["dw"], e["db"], B)
            for i, cv in enumerate(self.convs):
                ops.relumax_bwd(self.dpooled, self.arg, self.keepT if dropT else None, 1.0 / (1.0 - self.pT) if dropT else 1.0, cv["dout"], cv["dbias"],
                                i * self.C)
                ops.conv_wgrad(cv["geom"], self.xT16, cv["dout"], cv["dw"], self.wgrad_ws)

        def lstm_bwd_chain(off, key):  # BPTT from d h_T = this encoder's columns of d fused
            L = self.lstm[key]
            return [lambda: L["dh"].copy_(self.dfused[:, off:off + H]),
                    lambda: ops.lstm_bwd(L["x"], L["w"][1], L["gates"], L["cs"], L["hs"], L["dh"], *L["dw"])]

        self._parallel([text_bwd], lstm_bwd_chain(0, "A"), lstm_bwd_chain(H, "V"))

    def run_update(self) -> None:
        eng, fs = self.eng, self.eng.fs

        def update():
            if eng.clip is not None:
                ops.clip_grad_scale(fs.G, eng.clip, 1.0 / eng.world, fs.hyper, len(fs.hyper_host), self.partial, self.grad_norm)
            fs.adam(0, fs.total, True)

        if eng.allreduce is not None:
            eng.allreduce(self, 0, update=update)
        else:
            update()

    def train_step(self, given_dropout: bool) -> None:
        eng = self.eng
        key = "train_given" if given_dropout else "train"
        if getattr(self, "_range_version", None) != eng.fs.range_version:
            self.graphs.clear()
            self._range_version = eng.fs.range_version
        if not eng.use_graphs:
            self.run_train(not given_dropout)
            return self.run_update()
        g = self.graphs.get(key)
        if g is None:
            if self.eager_steps < 2:
                before = ops.launch_count(eng.device.index)
                self.run_train(not given_dropout)
                self.run_update()
                self.launches_per_step = ops.launch_count(eng.device.index) - before
                self.eager_steps += 1
                return
            torch.cuda.synchronize(eng.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.run_train(not given_dropout)
                self.run_update()
            self.graphs[key] = g
        g.replay()


class UttEngine:
    def __init__(self, model: nn.Module, device: torch.device, seed: Optional[int] = None):
        self.model, self.device = model, device
        self.client_id = int(getattr(model, "_mml_client_id", 0))
        self.seed = ops.engine_seed(self.client_id) if seed is None else seed
        self.fs = FlatState(model, device)
        self.plans: Dict[Tuple[int, int], _UttPlan] = {}
        self.world = 1
        self.allreduce = None
        self.use_graphs = True
        self.clip = model.clip

    def plan_for(self, B: int, T: int) -> _UttPlan:
        plan = self.plans.get((B, T))
        if plan is None:
            plan = self.plans[(B, T)] = _UttPlan(self, B, T)
        return plan


class UttFusionModel(nn.Module):
    def __init__(self, netA: LSTMEncoder, netV: LSTMEncoder, netT: TextCNN, netC: FcClassifier, *, clip: Optional[float] = None,
                 pretrained_path: Optional[str] = None) -> None:
        super().__init__()
        self.netA, self.netV, self.netT, self.netC = netA, netV, netT, netC
        self.clip = clip
        self.pretrained_path = pretrained_path
        self._engine: Optional[UttEngine] = None
        self._dp = None
        self.world_size = 1

    # ---- plumbing ----------------------------------------------------------------------------------------------------------
    def train(self, mode: bool = True):
        super().train(mode)
        self._uniform_mode = bool(mode)
        return self

    def _set_mode(self, training: bool) -> None:
        if getattr(self, "_uniform_mode", None) is not training or self.training is not training:
            self.train(training)

    def _get_engine(self, device) -> UttEngine:
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("mml_b200.UttFusionModel runs on a B200 GPU only: there is no CPU / PyTorch fallback path")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        eng = self._engine
        if eng is None or eng.device != device:
            eng = self._engine = UttEngine(self, device)
            if self._dp is not None:
                self._dp.attach(eng)
        eng.clip = self.clip
        eng.fs.ensure_fresh()
        return eng

    def enable_data_parallel(self, dp) -> None:
        self._dp, self.world_size = dp, dp.world_size
        if self._engine is not None:
            dp.attach(self._engine)

    def flatten_parameters(self) -> None:
        p = next(self.parameters())
        if p.is_cuda:
            self._get_engine(p.device)

    def get_encoder(self, modality) -> nn.Module:
        name = str(modality).lower().split(".")[-1]
        if name == "audio":
            return self.netA
        if name == "video":
            return self.netV
        if name == "text":
            return self.netT
        raise ValueError(f"Unknown modality: {modality}")

    # ---- staging -------------------------------------------------------------------------------------------------------------
    def _stage(self, eng: UttEngine, A, V, T, masks=(None, None, None), labels=None) -> _UttPlan:
        if A.dim() != 3 or V.dim() != 3 or T.dim() != 3 or not (A.shape[:2] == V.shape[:2] == T.shape[:2]):
            raise ValueError(f"expected aligned [B,T,F] sequences, got {tuple(A.shape)} / {tuple(V.shape)} / {tuple(T.shape)}")
        plan = eng.plan_for(A.shape[0], A.shape[1])
        if (A.shape[2], V.shape[2], T.shape[2]) != (plan.DA, plan.DV, plan.DT):
            raise ValueError("feature widths do not match the encoders")
        for dst, src in ((plan.xA, A), (plan.xV, V), (plan.xT, T)):
            _copy_in(dst, src if src.dtype == dst.dtype else src.float())
        for dst, m in zip((plan.mA, plan.mV, plan.mT), masks):
            if m is None:
                dst.fill_(1.0)
            else:
                _copy_in(dst, torch.as_tensor(m).reshape(-1).float())
        if labels is not None:
            ops.check_class_labels(labels, plan.logits.shape[1])
            _copy_in(plan.labels, torch.as_tensor(labels).reshape(-1))
        from .data import note_inputs_consumed
        note_inputs_consumed(eng.device)  # a prefetcher may overwrite the batch's device buffers from here on
        return plan

    def _unpack(self, batch: Dict[Any, Any]):
        A, V, T = _find(batch, "audio"), _find(batch, "video"), _find(batch, "text")
        masks = [None, None, None]
        for i, name in enumerate(("audio", "video", "text")):
            if f"{name}_original" in batch and f"{name}_missing_index" in batch:
                masks[i] = batch[f"{name}_missing_index"]
        if masks[0] is not None:
            A = batch["audio_original"]
        if masks[1] is not None:
            V = batch["video_original"]
        if masks[2] is not None:
            T = batch["text_original"]
        if A is None or V is None or T is None:
            raise KeyError("batch needs audio, video and text tensors (Modality keys or *_original + *_missing_index)")
        return A, V, T, masks, batch["label"], batch.get("pattern_name")

    # ---- forward / steps ---------------------------------------------------------------------------------------------------------
    def forward(self, A=None, V=None, T=None, *, is_embd_A: bool = False, is_embd_V: bool = False, is_embd_T: bool = False) -> torch.Tensor:
        assert not all((A is None, V is None, T is None)), "At least one of A, V, T must be provided"
        assert not all([is_embd_A, is_embd_V, is_embd_T]), "Cannot have all embeddings as True"
        if is_embd_A or is_embd_V or is_embd_T or A is None or V is None or T is None:
            raise NotImplementedError("mml_b200.UttFusionModel.forward needs the three raw modalities (embeddings / missing inputs are outside the hot path)")
        eng = self._get_engine(A.device if A.is_cuda else next(self.parameters()).device)
        plan = self._stage(eng, A, V, T)
        if self.training:
            eng_seed_step = eng.fs.step
            if plan.pT > 0:
                ops.dropout_mask(plan.keepT, plan.pT, eng.seed, eng_seed_step)
            for i, layer in enumerate(plan.dense):
                if layer["keep"] is not None:
                    ops.dropout_mask(layer["keep"], plan.pC, eng.seed + 101 * (i + 1), eng_seed_step)
        plan.run_forward(self.training, False, False)
        return plan.logits.clone()

    def train_step(self, batch, optimizer, loss_functions, device, metric_recorder=None, **kwargs) -> Dict[str, Any]:
        """One fused training step; returns {"loss": float} like utt_fusion.py:151-198."""
        eng = self._get_engine(device)
        AVMNIST._check_loss(loss_functions)
        A, V, T, masks, labels, miss_type = self._unpack(batch)
        self._set_mode(True)
        fs = eng.fs
        fs.adopt_optimizer(optimizer)
        fs.sync_hyper(optimizer, 1.0 / self.world_size)
        plan = self._stage(eng, A, V, T, masks, labels)
        given = kwargs.get("dropout_masks")
        if given is not None:
            plan.keepT.copy_(torch.as_tensor(given[0]).reshape(plan.keepT.shape).to(torch.uint8), non_blocking=True)
            for layer, km in zip(plan.dense, given[1:]):
                layer["keep"].copy_(torch.as_tensor(km).reshape(layer["keep"].shape).to(torch.uint8), non_blocking=True)
        plan.train_step(given_dropout=given is not None)
        fs._host_step += 1
        return self._finish(eng, plan, labels, miss_type, metric_recorder, False)

    def validation_step(self, batch, loss_functions, device, metric_recorder=None, return_test_info: bool = False, **kwargs) -> Dict[str, Any]:
        eng = self._get_engine(device)
        AVMNIST._check_loss(loss_functions)
        A, V, T, masks, labels, miss_type = self._unpack(batch)
        self._set_mode(False)
        plan = self._stage(eng, A, V, T, masks, labels)
        plan.run_forward(False, True, False)
        return self._finish(eng, plan, labels, miss_type, metric_recorder, return_test_info)

    def _finish(self, eng, plan, labels, miss_type, metric_recorder, return_test_info) -> Dict[str, Any]:
        plan.h_loss.copy_(plan.loss, non_blocking=True)
        if metric_recorder is not None or return_test_info:
            plan.h_pred.copy_(plan.pred, non_blocking=True)
        torch.cuda.current_stream(eng.device).synchronize()
        loss = float(plan.h_loss[0])
        if metric_recorder is None and not return_test_info:
            return {"loss": loss}
        predictions = plan.h_pred.numpy().astype(np.int64)
        targets = torch.as_tensor(labels).detach().cpu().reshape(-1).numpy()
        mt = np.array(miss_type)
        if metric_recorder is not None:
            metric_recorder.update_group_all("classification", predictions=predictions, targets=targets, m_types=mt)
        if return_test_info:
            return {"loss": loss, "predictions": predictions, "labels": targets, "miss_types": mt}
        return {"loss": loss}
