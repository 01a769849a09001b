"""The fused MMIMDb gated late-fusion step (BASELINE config 3; MML_Suite/models/mmimdb.py:166-245).

Host-side replacement for autograd + ~150 ATen launches: one static schedule per batch size, captured into a CUDA graph
after two eager steps.  Layout in HBM:

  * parameters: the same ``FlatState`` as the AVMNIST engine (fp32 master ``P``, gradients ``G``, Adam ``M``/``V``, bf16
    shadow ``Wb``).  nn.Linear's [out][in] weight is already the K-major B operand of the tensor-core GEMM.  The two
    encoder Linears carry a bias: they are stored "augmented" -- [out][ld], ld = in+1 rounded up to 64, bias in column
    ``in`` -- and their input rows carry a constant 1 there, so the GEMM adds the bias and its wgrad produces the bias
    gradient (pad columns are 0 in weights and activations and stay 0 under Adam).  The two Linear units of a MaxOut are
    consecutive in ``P`` and are read as ONE [2*hidden][in] matrix: one GEMM gives both candidates side by side.
  * activations: bf16 rows [B][C] (GEMM operands / outputs), fp32 for the saved normalised values (xhat), the GMU
    branches and the final normalised features.

Schedule (train): 2x bn1d(INPUT, mask fused) -> 2x encoder GEMM -> 2x GMU GEMM -> gmu_fwd -> bn1d(GATED) -> GEMM ->
bn1d(MAXOUT) -> GEMM -> bn1d(MAXOUT) -> bce_head_fwd | bce_head_bwd -> bn1d_bwd/dgrad/wgrad chain -> Adam.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .engine import ALIGN, BF16, BN_EPS, BN_MOMENTUM, FlatState, _round_up

DROPOUT_P = 0.5  # MLPGenreClassifier hard-codes Dropout(p=0.5) (mmimdb.py:42,45)


class GatedFusionEngine:
    def __init__(self, model: nn.Module, device: torch.device, seed: Optional[int] = None):
        self.model, self.device = model, device
        self.client_id = int(getattr(model, "_mml_client_id", 0))
        self.seed = ops.engine_seed(self.client_id) if seed is None else seed
        self.fs = FlatState(model, device, augment={"image_model.net.1.weight": "image_model.net.1.bias",
                                                    "text_model.net.1.weight": "text_model.net.1.bias"})
        self.plans: Dict[int, "_GatedPlan"] = {}
        self.world = 1
        self.allreduce: Optional[Callable] = None
        self.use_graphs = True
        self.dropout_p = DROPOUT_P

    def plan_for(self, B: int) -> "_GatedPlan":
        plan = self.plans.get(B)
        if plan is None:
            plan = self.plans[B] = _GatedPlan(self, B)
        return plan


class _GatedPlan:
    def __init__(self, eng: GatedFusionEngine, B: int):
        self.eng, self.B = eng, B
        fs, dev, model = eng.fs, eng.device, eng.model
        params = dict(model.named_parameters())
        DI = params["image_model.net.0.weight"].numel()
        DT = params["text_model.net.0.weight"].numel()
        E = params["image_model.net.1.weight"].shape[0]
        H = params["mm_mlp.net.1.layers.0.weight"].shape[0]
        NC = params["mm_mlp.net.7.weight"].shape[0]
        self.pooling = getattr(model, "fusion_type", "gated") == "pooling"
        wa, wb = ("fusion_module.proj_a.weight", "fusion_module.proj_b.weight") if self.pooling else \
            ("fusion_module.fc_one.weight", "fusion_module.fc_two.weight")
        if params[wa].shape != (E, E) or params[wb].shape != (E, E) \
                or params["mm_mlp.net.1.layers.0.weight"].shape[1] != E or params["text_model.net.1.weight"].shape[0] != E:
            raise NotImplementedError("mml_b200 gated fusion expects equal embedding / gate widths (mmimdb_baseline.yaml: 512)")
        self.pool_type = model.fusion_module.pooling_type if self.pooling else None
        self.pool_p = float(model.fusion_module.dropout) if self.pooling else 0.0
        self.pool_net = {"attention": "fusion_module.attention_layer", "gated": "fusion_module.gate_layer"}.get(self.pool_type)
        if self.pool_net is not None:
            self.Hd = params[self.pool_net + ".0.weight"].shape[0]
            if self.Hd % 64:
                raise NotImplementedError("the pooling network's hidden width must be a multiple of 64 for the tensor-core GEMM")
        if E % 64 or H % 64:
            raise NotImplementedError("embedding and hidden widths must be multiples of 64 for the tensor-core GEMMs")
        if len(model.mm_mlp.net[1].layers) != 2 or len(model.mm_mlp.net[4].layers) != 2:
            raise NotImplementedError("MaxOut with num_units != 2")
        self.DI, self.DT, self.E, self.H, self.NC = DI, DT, E, H, NC
        LI, LT = _round_up(DI + 1, ALIGN), _round_up(DT + 1, ALIGN)

        def f32(*shape):
            return torch.zeros(*shape, device=dev)

        def b16(*shape):
            return torch.zeros(*shape, device=dev, dtype=BF16)

        # ---- inputs
        self.xI, self.xT = f32(B, DI), f32(B, DT)
        self.mI, self.mT = torch.ones(B, device=dev), torch.ones(B, device=dev)
        self.labels = f32(B, NC)
        # ---- activations
        self.xnI, self.xnT = b16(B, LI), b16(B, LT)
        self.xnI[:, DI] = 1.0  # the bias column
        self.xnT[:, DT] = 1.0
        self.xhI, self.xhT = f32(B, DI), f32(B, DT)
        self.eI, self.eT, self.h1p, self.h2p = b16(B, E), b16(B, E), b16(B, E), b16(B, E)
        self.h1, self.h2, self.gate = f32(B, E), f32(B, E), f32(B)
        self.xh0, self.xn0 = f32(B, E), b16(B, E)
        self.pre1, self.xh1, self.xn1 = b16(B, 2 * H), f32(B, H), b16(B, H)
        self.pre2, self.xh2, self.xn2 = b16(B, 2 * H), f32(B, H), f32(B, H)
        self.keep = torch.ones(2, B, H, device=dev, dtype=torch.uint8)  # both dropout masks: one generator launch
        self.keep1, self.keep2 = self.keep[0], self.keep[1]
        self.keepAB = torch.ones(2, B, E, device=dev, dtype=torch.uint8)  # MultimodalPooling dropout, one mask per branch
        self.keepA, self.keepB = self.keepAB[0], self.keepAB[1]
        self.inv = {k: f32(n) for k, n in (("I", DI), ("T", DT), ("0", E), ("1", H), ("2", H))}
        self.logits, self.dlogits = f32(B, NC), f32(B, NC)
        self.loss = f32(1)
        self.pred = torch.zeros(B, NC, device=dev, dtype=torch.uint8)
        self.scratch = f32(ops.bce_head_scratch_floats(B))
        # ---- gradients of activations
        self.dxn2, self.dpre2, self.dxn1, self.dpre1, self.dxn0 = b16(B, H), b16(B, 2 * H), b16(B, H), b16(B, 2 * H), b16(B, E)
        self.dz = f32(B, E)
        self.dh1p, self.dh2p, self.deI, self.deT = b16(B, E), b16(B, E), b16(B, E), b16(B, E)
        self.dxnI, self.dxnT = b16(B, LI), b16(B, LT)
        if self.pool_net is not None:  # attention / gated pooling: [a | b] -> Linear -> tanh -> Linear -> softmax / sigmoid
            self.comb, self.dcomb = b16(B, 2 * E), b16(B, 2 * E)
            self.hidp, self.dhidp, self.tpool = b16(B, self.Hd), b16(B, self.Hd), f32(B, self.Hd)
        # ---- pinned host mirrors
        self.h_loss = torch.zeros(1).pin_memory()
        self.h_pred = torch.zeros(B, NC, dtype=torch.uint8).pin_memory()
        self.threshold = 0.5
        self.graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        import os as _os
        # schedule: the text branch runs next to the image branch on a side stream, weight gradients (needed by Adam only)
        # on a third one; MML_GATED_STREAMS=0 serialises everything on one stream (A/B timing)
        multi = _os.environ.get("MML_GATED_STREAMS", "1") == "1"
        self.side = torch.cuda.Stream(device=dev) if multi else None
        self.wstream = torch.cuda.Stream(device=dev) if multi else None
        # every weight-gradient launch of a step runs on ONE stream (wstream, or the only stream when not multi): one scratch
        self.wgrad_ws = self.wgrad_ws_main = ops.WgradScratch(dev)
        self.eager_steps = 0
        self.launches_per_step = 0
        self._build(params, LI, LT)

    # ----------------------------------------------------------------------------------------------------------------
    def _build(self, params, LI: int, LT: int) -> None:
        fs, B, E, H, NC = self.eng.fs, self.B, self.E, self.H, self.NC
        P, G, Wb = fs.P, fs.G, fs.Wb

        def par(flat, name):
            return fs.flat_slice(flat, name).view(params[name].shape)

        def bn(prefix):
            n = params[prefix + ".weight"].numel()
            om, ov = fs.buf_offsets[prefix + ".running_mean"], fs.buf_offsets[prefix + ".running_var"]
            return (par(P, prefix + ".weight"), par(P, prefix + ".bias"), fs.S[om:om + n], fs.S[ov:ov + n],
                    par(G, prefix + ".weight"), par(G, prefix + ".bias"))

        gI, bI, rmI, rvI, dgI, dbI = bn("image_model.net.0")
        gT, bT, rmT, rvT, dgT, dbT = bn("text_model.net.0")
        g0, b0, rm0, rv0, dg0, db0 = bn("mm_mlp.net.0")
        g1, b1, rm1, rv1, dg1, db1 = bn("mm_mlp.net.3")
        g2, b2, rm2, rv2, dg2, db2 = bn("mm_mlp.net.6")
        scale = 1.0 / (1.0 - DROPOUT_P)
        kw = dict(momentum=BN_MOMENTUM, eps=BN_EPS)
        self.f_bnI = ops.bn1d_fwd_desc(ops.BN1D_INPUT, B, self.DI, gI, bI, rmI, rvI, x=self.xI, mask=self.mI, xhat=self.xhI, invstd=self.inv["I"],
                                       y_bf16=self.xnI, **kw)
        self.f_bnT = ops.bn1d_fwd_desc(ops.BN1D_INPUT, B, self.DT, gT, bT, rmT, rvT, x=self.xT, mask=self.mT, xhat=self.xhT, invstd=self.inv["T"],
                                       y_bf16=self.xnT, **kw)
        if not self.pooling:
            self.f_bn0 = ops.bn1d_fwd_desc(ops.BN1D_GATED, B, E, g0, b0, rm0, rv0, h1=self.h1, h2=self.h2, gate=self.gate, xhat=self.xh0,
                                           invstd=self.inv["0"], y_bf16=self.xn0, **kw)
        else:  # pooling.py:100-126: max | (a + b) / 2 | a + b | att_a a + att_b b | g a + (1 - g) b
            self.mix = {"max": (1.0, 1.0), "avg": (0.5, 0.5), "average": (0.5, 0.5), "sum": (1.0, 1.0)}.get(self.pool_type, (0.0, 0.0))
            mode0 = ops.BN1D_MAX2 if self.pool_type == "max" else ops.BN1D_GATED
            self.f_bn0 = ops.bn1d_fwd_desc(mode0, B, E, g0, b0, rm0, rv0, h1=self.h1, h2=self.h2, gate=self.gate if self.pool_net else None,
                                           xhat=self.xh0, invstd=self.inv["0"], y_bf16=self.xn0, mix_a=self.mix[0], mix_b=self.mix[1], **kw)
        self.f_bn1 = ops.bn1d_fwd_desc(ops.BN1D_MAXOUT, B, H, g1, b1, rm1, rv1, pre=self.pre1, keep=self.keep1, keep_scale=scale, xhat=self.xh1,
                                       invstd=self.inv["1"], y_bf16=self.xn1, **kw)
        self.f_bn2 = ops.bn1d_fwd_desc(ops.BN1D_MAXOUT, B, H, g2, b2, rm2, rv2, pre=self.pre2, keep=self.keep2, keep_scale=scale, xhat=self.xh2,
                                       invstd=self.inv["2"], y_f32=self.xn2, **kw)
        self.b_bn2 = ops.bn1d_bwd_desc(ops.BN1D_MAXOUT, B, H, self.dxn2, self.xh2, g2, self.inv["2"], dg2, db2, pre=self.pre2, keep=self.keep2,
                                       keep_scale=scale, dpre=self.dpre2)
        self.b_bn1 = ops.bn1d_bwd_desc(ops.BN1D_MAXOUT, B, H, self.dxn1, self.xh1, g1, self.inv["1"], dg1, db1, pre=self.pre1, keep=self.keep1,
                                       keep_scale=scale, dpre=self.dpre1)
        self.b_bn0 = ops.bn1d_bwd_desc(ops.BN1D_GATED, B, E, self.dxn0, self.xh0, g0, self.inv["0"], dg0, db0, dz=self.dz)  # also serves MAX2
        self.b_bnI = ops.bn1d_bwd_desc(ops.BN1D_INPUT, B, self.DI, self.dxnI, self.xhI, gI, self.inv["I"], dgI, dbI)
        self.b_bnT = ops.bn1d_bwd_desc(ops.BN1D_INPUT, B, self.DT, self.dxnT, self.xhT, gT, self.inv["T"], dgT, dbT)
        # ---- GEMMs: (geometry, weights bf16, weight gradient fp32)
        def gemm(n_in, n_out):
            return ops.make_geom(B, 1, 1, n_in, n_out, 1, 1, 1, 0)

        self.gemI = (gemm(LI, E), fs.aug_matrix(Wb, "image_model.net.1.weight"), fs.aug_matrix(G, "image_model.net.1.weight"))
        self.gemT = (gemm(LT, E), fs.aug_matrix(Wb, "text_model.net.1.weight"), fs.aug_matrix(G, "text_model.net.1.weight"))
        n1, n2 = ("fusion_module.proj_a.weight", "fusion_module.proj_b.weight") if self.pooling else \
            ("fusion_module.fc_one.weight", "fusion_module.fc_two.weight")
        self.gem1 = (gemm(E, E), par(Wb, n1), par(G, n1))
        self.gem2 = (gemm(E, E), par(Wb, n2), par(G, n2))

        def maxout(prefix, n_in):
            a, b = prefix + ".layers.0.weight", prefix + ".layers.1.weight"
            o = fs.offsets[a]
            if fs.offsets[b] != o + H * n_in:
                raise RuntimeError("MaxOut units are not adjacent in the flat parameter buffer")
            return gemm(n_in, 2 * H), Wb[o:o + 2 * H * n_in].view(2 * H, n_in), G[o:o + 2 * H * n_in].view(2 * H, n_in)

        self.gemM1 = maxout("mm_mlp.net.1", E)
        self.gemM2 = maxout("mm_mlp.net.4", H)
        if self.pooling:
            self.pb = (par(P, "fusion_module.proj_a.bias"), par(P, "fusion_module.proj_b.bias"))
            self.dpb = (par(G, "fusion_module.proj_a.bias"), par(G, "fusion_module.proj_b.bias"))
            if self.pool_net is not None:
                n0, n2 = self.pool_net + ".0", self.pool_net + ".2"
                self.gemA = (gemm(2 * E, self.Hd), par(Wb, n0 + ".weight"), par(G, n0 + ".weight"))
                self.att = (par(P, n0 + ".bias"), par(P, n2 + ".weight"), par(P, n2 + ".bias"))
                self.datt = (par(G, n0 + ".bias"), par(G, n2 + ".weight"), par(G, n2 + ".bias"))
        else:
            self.wz, self.dwz = fs.flat_slice(P, "fusion_module.hidden_sigmoid.weight"), fs.flat_slice(G, "fusion_module.hidden_sigmoid.weight")
        self.w7, self.b7 = par(P, "mm_mlp.net.7.weight"), par(P, "mm_mlp.net.7.bias")
        self.dw7, self.db7 = par(G, "mm_mlp.net.7.weight"), par(G, "mm_mlp.net.7.bias")

    # ----------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _fprop(gem, x, y):
        ops.conv_fprop(gem[0], x, gem[1], y, None)

    def _bprop(self, gem, x, dy, dx):
        """dgrad on the current stream (critical path), wgrad on the weight-gradient stream."""
        if self.wstream is None:
            ops.conv_wgrad(gem[0], x, dy, gem[2], self.wgrad_ws_main)
        else:
            self.wstream.wait_stream(torch.cuda.current_stream(self.eng.device))
            with torch.cuda.stream(self.wstream):
                ops.conv_wgrad(gem[0], x, dy, gem[2], self.wgrad_ws)
        if dx is not None:
            ops.conv_dgrad(gem[0], dy, gem[1], dx)

    def _fork(self, main_ops, side_ops) -> None:
        """two independent kernel chains: ``main_ops`` on the current stream, ``side_ops`` on the side stream; joined."""
        if self.side is None:
            for op in main_ops + side_ops:
                op()
            return
        main = torch.cuda.current_stream(self.eng.device)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            for op in side_ops:
                op()
        for op in main_ops:
            op()
        main.wait_stream(self.side)

    def run_forward(self, train: bool, dropout: bool, with_loss: bool, with_grad: bool) -> None:
        self._fork([lambda: ops.bn1d_fwd(self.f_bnI, train), lambda: self._fprop(self.gemI, self.xnI, self.eI),
                    lambda: self._fprop(self.gem1, self.eI, self.h1p)],
                   [lambda: ops.bn1d_fwd(self.f_bnT, train), lambda: self._fprop(self.gemT, self.xnT, self.eT),
                    lambda: self._fprop(self.gem2, self.eT, self.h2p)])
        if self.pooling:
            drop = train and dropout and self.pool_p > 0
            ops.pool_fwd(self.h1p, self.h2p, self.pb[0], self.pb[1], self.keepA if drop else None, self.keepB if drop else None,
                         1.0 / (1.0 - self.pool_p), self.h1, self.h2, self.comb if self.pool_net else None)
            if self.pool_net is not None:
                self._fprop(self.gemA, self.comb, self.hidp)
                ops.att_fwd(self.hidp, self.att[0], self.att[1], self.att[2], self.tpool, self.gate)
        else:
            ops.gmu_fwd(self.h1p, self.h2p, self.wz, self.h1, self.h2, self.gate)
        ops.bn1d_fwd(self.f_bn0, train)
        self._fprop(self.gemM1, self.xn0, self.pre1)
        ops.bn1d_fwd(self.f_bn1, train, use_keep=dropout)
        self._fprop(self.gemM2, self.xn1, self.pre2)
        ops.bn1d_fwd(self.f_bn2, train, use_keep=dropout)
        ops.bce_head_fwd(self.xn2, self.w7, self.b7, self.labels if with_loss else None, self.logits, self.loss if with_loss else None,
                         self.dlogits if with_grad else None, self.pred, self.scratch, self.threshold, 1.0)

    def run_train(self, own_dropout: bool) -> None:
        eng, fs = self.eng, self.eng.fs
        fs.G.zero_()
        if own_dropout:
            ops.dropout_mask(self.keep, DROPOUT_P, eng.seed, fs.step)
            if self.pooling and self.pool_p > 0:
                ops.dropout_mask(self.keepAB, self.pool_p, eng.seed ^ 0x51ED, fs.step)
        self.run_forward(True, True, True, True)
        ops.bce_head_bwd(self.dlogits, self.xn2, self.w7, self.dw7, self.db7, self.dxn2)
        ops.bn1d_bwd(self.b_bn2)
        self._bprop(self.gemM2, self.xn1, self.dpre2, self.dxn1)
        ops.bn1d_bwd(self.b_bn1)
        self._bprop(self.gemM1, self.xn0, self.dpre1, self.dxn0)
        ops.bn1d_bwd(self.b_bn0)
        if self.pooling:
            drop = self.pool_p > 0
            if self.pool_net is not None:
                ops.att_bwd(self.dz, self.h1, self.h2, self.gate, self.tpool, self.att[1], self.datt[1], self.datt[2], self.datt[0], self.dhidp)
                self._bprop(self.gemA, self.comb, self.dhidp, self.dcomb)
            ops.pool_bwd(self.dz, self.h1, self.h2, self.keepA if drop else None, self.keepB if drop else None, 1.0 / (1.0 - self.pool_p),
                         0 if self.pool_type == "max" else 1, self.mix[0], self.mix[1], self.dh1p, self.dh2p, self.dpb[0], self.dpb[1],
                         gate=self.gate if self.pool_net else None, dcomb=self.dcomb if self.pool_net else None)
        else:
            ops.gmu_bwd(self.dz, self.h1, self.h2, self.gate, self.wz, self.dwz, self.dh1p, self.dh2p)
        self._fork([lambda: self._bprop(self.gem1, self.eI, self.dh1p, self.deI), lambda: self._bprop(self.gemI, self.xnI, self.deI, self.dxnI),
                    lambda: ops.bn1d_bwd(self.b_bnI)],
                   [lambda: self._bprop(self.gem2, self.eT, self.dh2p, self.deT), lambda: self._bprop(self.gemT, self.xnT, self.deT, self.dxnT),
                    lambda: ops.bn1d_bwd(self.b_bnT), lambda: fs.NBT.add_(1)])
        if self.wstream is not None:
            torch.cuda.current_stream(eng.device).wait_stream(self.wstream)

    def run_update(self) -> None:
        eng, fs = self.eng, self.eng.fs

        def adam():
            fs.adam(0, fs.total, True)

        if eng.allreduce is not None:
            eng.allreduce(self, 0, update=adam)
        else:
            adam()

    def train_step(self, given_dropout: bool) -> None:
        eng = self.eng
        if not eng.use_graphs:
            self.run_train(not given_dropout)
            self.run_update()
            return
        key = "train_given" if given_dropout else "train"
        if getattr(self, "_range_version", None) != eng.fs.range_version:  # new optimizer grouping: the Adam launches changed
            self.graphs.pop("train", None), self.graphs.pop("train_given", None)
            self._range_version = eng.fs.range_version
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

    def encode(self, which: str, x: torch.Tensor, training: bool) -> torch.Tensor:
        """MMIMDbModalityEncoder.forward (mmimdb.py:82-92) of one modality: BatchNorm1d -> Linear, fp32 [B, E]."""
        fs = self.eng.fs
        if which == "image":
            src, mask, desc, gem, xn, out, prefix = self.xI, self.mI, self.f_bnI, self.gemI, self.xnI, self.eI, "image_model"
        else:
            src, mask, desc, gem, xn, out, prefix = self.xT, self.mT, self.f_bnT, self.gemT, self.xnT, self.eT, "text_model"
        src.copy_(x, non_blocking=True)
        mask.fill_(1.0)
        ops.bn1d_fwd(desc, training)
        self._fprop(gem, xn, out)
        if training:
            fs.NBT[fs.nbt_names.index(prefix + ".net.0.num_batches_tracked")] += 1
        return out.float()

    def run_eval(self, with_loss: bool) -> None:
        self.run_forward(False, False, with_loss, False)

    def run_forward_train_mode(self) -> None:
        eng, fs = self.eng, self.eng.fs
        ops.dropout_mask(self.keep, DROPOUT_P, eng.seed, fs.step)
        if self.pooling and self.pool_p > 0:
            ops.dropout_mask(self.keepAB, self.pool_p, eng.seed ^ 0x51ED, fs.step)
        self.run_forward(True, True, False, False)
        fs.NBT += 1
