"""The fused late-fusion step: flat parameter storage, static activation buffers and the kernel schedule.

This is the host-side replacement for what autograd + ~400 ATen/cuDNN launches do in the reference's
``AVMNIST.train_step`` (MML_Suite/models/avmnist.py:269-310): zero_grad -> forward -> CrossEntropy -> backward ->
Adam.step -> argmax.  Everything numerical is a call into libmml_b200.so (``ops``); PyTorch provides device memory,
streams, CUDA-graph capture and (for N > 1) the NCCL process group.

Layout in HBM
  * ``FlatState``: every parameter lives in ONE fp32 buffer ``P`` (conv weights physically K,R,S,C == torch
    channels_last, exposed to PyTorch as OIHW views so ``state_dict()`` keeps the reference's 346 names/shapes/dtypes);
    gradients ``G``, Adam moments ``M``/``V`` mirror it element for element; ``Wb`` is the bf16 shadow the tensor-core
    kernels read (fprop as a K-major, dgrad as an MN-major B operand -- no transposed copy).  BatchNorm running
    statistics live in ``S``.
  * activations: NHWC bf16, one static buffer per conv output ("raw") and per block activation, kept for backward.
  * one schedule of closures per (batch, input size); after two eager steps it is captured into a CUDA graph.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops

BF16 = torch.bfloat16
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
ALIGN = 64  # elements; 256 B for fp32, 128 B for bf16 (TMA base alignment)


def _round_up(n: int, a: int) -> int:
    return (n + a - 1) // a * a


# =====================================================================================================================
# flat parameter / buffer storage
# =====================================================================================================================
class FlatState:
    def __init__(self, module: nn.Module, device: torch.device, augment: Optional[Dict[str, str]] = None,
                 pad: Optional[Dict[str, int]] = None):
        """``augment`` = {linear weight name: its bias name}: such a pair is stored as ONE [out][ld] matrix, ld = in + 1
        rounded up to 64, with the bias in column ``in`` (the tensor-core GEMM then adds the bias through a constant-1
        activation column and its wgrad yields the bias gradient); ``weight`` / ``bias`` stay visible as strided views.

        ``pad`` = {parameter or buffer name: channel count Cp}: a conv weight [K,C,R,S] is stored as [Cp,R,S,Cp] and a per-channel
        vector [C] as [Cp], zero outside the logical part (narrow layers run as Cp-channel tensor-core convolutions; padded
        weights, gradients and Adam moments are zero and stay zero); PyTorch sees the logical slice as a strided view."""
        self.module = module
        self.pad: Dict[str, int] = dict(pad or {})
        self.aug: Dict[str, Tuple[int, int, int, int]] = {}  # weight or bias name -> (offset, out, in, ld)
        augment = augment or {}
        aug_bias = set(augment.values())
        self.device = device
        self.param_names: List[str] = []
        self.offsets: Dict[str, int] = {}
        self.numels: Dict[str, int] = {}
        off = 0
        for name, p in module.named_parameters():
            self.param_names.append(name)
            if name in aug_bias:
                continue  # placed together with its weight
            if name in augment:
                n_out, n_in = p.shape
                ld = _round_up(n_in + 1, ALIGN)
                self.aug[name] = self.aug[augment[name]] = (off, n_out, n_in, ld)
                self.offsets[name], self.numels[name] = off, n_out * ld
                self.offsets[augment[name]], self.numels[augment[name]] = off + n_in, n_out
                off = _round_up(off + n_out * ld, ALIGN)
                continue
            self.offsets[name] = off
            self.numels[name] = self._padded_numel(name, p)
            off = _round_up(off + self.numels[name], ALIGN)
        self.total = off
        self.P = torch.zeros(off, device=device)
        self.G = torch.zeros(off, device=device)
        self.M = torch.zeros(off, device=device)
        self.V = torch.zeros(off, device=device)
        self.Wb = torch.zeros(off, device=device, dtype=BF16)
        # running statistics
        self.buf_offsets: Dict[str, int] = {}
        soff, nbt = 0, []
        for name, b in module.named_buffers():
            if name.endswith("num_batches_tracked"):
                nbt.append(name)
            else:
                self.buf_offsets[name] = soff
                soff = _round_up(soff + self._padded_numel(name, b), 4)
        self.S = torch.zeros(max(soff, 4), device=device)
        self.nbt_names = nbt
        self.NBT = torch.zeros(max(len(nbt), 1), device=device, dtype=torch.int64)
        self.hyper = torch.zeros(self.MAX_GROUPS, 8, device=device)
        self.hyper_host: List[Optional[Tuple[float, ...]]] = [None] * self.MAX_GROUPS
        self.ranges = None
        self.range_version = 0
        self.step = torch.zeros(1, device=device, dtype=torch.int64)
        self._versions = None
        self._sentinels = None
        self._plist = list(module.parameters())
        self.bind(copy_in=True)

    def _padded_numel(self, name: str, t: torch.Tensor) -> int:
        cp = self.pad.get(name)
        if cp is None:
            return t.numel()
        if t.dim() == 4:
            if t.shape[0] > cp or t.shape[1] > cp:
                raise ValueError(f"{name}: shape {tuple(t.shape)} exceeds the padded channel count {cp}")
            return cp * cp * t.shape[2] * t.shape[3]
        if t.dim() != 1 or t.numel() > cp:
            raise ValueError(f"{name}: only conv weights and per-channel vectors can be padded")
        return cp

    # -- views ------------------------------------------------------------------------------------------------
    def _view(self, flat: torch.Tensor, name: str, like: torch.Tensor) -> torch.Tensor:
        if name in self.aug:
            o, n_out, n_in, ld = self.aug[name]
            m = flat[o:o + n_out * ld].view(n_out, ld)
            return m[:, :n_in] if like.dim() == 2 else m[:, n_in]
        o, n = self.offsets[name], self.numels[name]
        v = flat[o:o + n]
        cp = self.pad.get(name)
        if like.dim() == 4:
            K, Cc, R, S = like.shape
            if cp is not None:
                return v.view(cp, R, S, cp).permute(0, 3, 1, 2)[:K, :Cc]
            return v.view(K, R, S, Cc).permute(0, 3, 1, 2)  # OIHW view of K,R,S,C storage (channels_last)
        if cp is not None:
            return v[:like.numel()]
        return v.view(like.shape)

    def flat_slice(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        o, n = self.offsets[name], self.numels[name]
        return flat[o:o + n]

    def aug_matrix(self, flat: torch.Tensor, weight_name: str) -> torch.Tensor:
        """[out][ld] weight-plus-bias-column matrix of an augmented Linear."""
        o, n_out, _, ld = self.aug[weight_name]
        return flat[o:o + n_out * ld].view(n_out, ld)

    def bind(self, copy_in: bool) -> None:
        """Point every parameter / buffer of the module at its slot (copying current values in first)."""
        with torch.no_grad():
            for name, p in self.module.named_parameters():
                v = self._view(self.P, name, p)
                if copy_in:
                    v.copy_(p.detach().to(self.device, torch.float32))
                p.data = v
                p.grad = self._view(self.G, name, p)
            bufs = dict(self.module.named_buffers())
            for name, o in self.buf_offsets.items():
                b = bufs[name]
                v = self.S[o:o + b.numel()].view(b.shape)  # a padded buffer's tail (zeros) follows in S
                if copy_in:
                    v.copy_(b.to(self.device, torch.float32))
                self._set_buffer(name, v)
            for i, name in enumerate(self.nbt_names):
                if copy_in:
                    self.NBT[i] = int(bufs[name])
                self._set_buffer(name, self.NBT[i])
        self.refresh_shadows()

    def _set_buffer(self, name: str, value: torch.Tensor) -> None:
        mod = self.module
        *path, leaf = name.split(".")
        for part in path:
            mod = getattr(mod, part)
        mod._buffers[leaf] = value

    def is_bound(self) -> bool:
        """Cheap per-step check (first / last parameter); ``bind`` re-homes everything if an outside ``.to()`` or
        ``param.data = ...`` moved them."""
        if self._sentinels is None:
            named = list(self.module.named_parameters())
            self._plist = [p for _, p in named]
            self._sentinels = [(p, self.P.data_ptr() + 4 * self.offsets[n]) for n, p in (named[0], named[len(named) // 2], named[-1])]
        return all(p.data_ptr() == addr for p, addr in self._sentinels)

    def param_versions(self) -> int:
        return sum(p._version for p in self._plist)

    def refresh_shadows(self) -> None:
        """fp32 master -> bf16 operand copies (after load_state_dict / external optimizer steps)."""
        ops.cast_f32_bf16(self.P, self.Wb)
        self._versions = self.param_versions()

    def ensure_fresh(self) -> None:
        if not self.is_bound():
            self.bind(copy_in=True)
        elif self._versions != self.param_versions():
            self.refresh_shadows()

    # -- optimizer ---------------------------------------------------------------------------------------------
    MAX_GROUPS = 8

    def adopt_optimizer(self, optimizer: torch.optim.Optimizer) -> None:
        """Validate that ``optimizer`` is the Adam the fused kernel implements, alias its state to M / V and map its param
        groups onto contiguous ranges of the flat buffers (one Adam launch per range, each with its own lr / betas / eps /
        weight_decay -- the reference's pretrained-encoder runs use per-encoder groups, train_multimodal.py:213-300)."""
        groups = optimizer.param_groups
        # per-step fast path: same optimizer object, same group shapes (count, first and last parameter of every group), state still aliased
        quick = tuple((len(g["params"]), id(g["params"][0]), id(g["params"][-1])) for g in groups if g["params"])
        if getattr(self, "_adopted", None) is optimizer and quick == getattr(self, "_adopted_quick", None) and self._state_aliased(optimizer):
            return
        if type(optimizer) is not torch.optim.Adam:
            raise NotImplementedError(
                f"mml_b200 fused train_step implements torch.optim.Adam (the optimizer of the reference's YAMLs); got {type(optimizer).__name__}. "
                "There is no silent fallback.")
        if len(groups) > self.MAX_GROUPS:
            raise NotImplementedError(f"at most {self.MAX_GROUPS} optimizer param groups are supported")
        for g in groups:
            if g["amsgrad"] or g["maximize"]:
                raise NotImplementedError("amsgrad / maximize are not supported by the fused Adam kernel")
        group_of = {id(p): gi for gi, g in enumerate(groups) for p in g["params"]}
        named = list(self.module.named_parameters())
        if set(group_of) != {id(p) for _, p in named} or sum(len(g["params"]) for g in groups) != len(named):
            raise NotImplementedError("the optimizer must hold exactly the parameters of the model")
        # contiguous flat ranges of equal group (alignment padding between tensors belongs to the range; it is zero and stays zero)
        spans = []
        for name, p in named:
            gi = group_of[id(p)]
            if name in self.aug:
                o, n_out, _, ld = self.aug[name]
                other = [q for q in self.aug if self.aug[q] == self.aug[name] and q != name]
                if other and group_of[id(dict(named)[other[0]])] != gi:
                    raise NotImplementedError(f"{name} and its bias are stored as one matrix and must share an optimizer param group")
                spans.append((o, _round_up(o + n_out * ld, ALIGN), gi))
            else:
                o = self.offsets[name]
                spans.append((o, _round_up(o + self.numels[name], ALIGN), gi))
        spans.sort()
        ranges: List[List[int]] = []
        for a0, b0, gi in spans:
            if ranges and ranges[-1][2] == gi and a0 <= ranges[-1][1]:
                ranges[-1][1] = max(ranges[-1][1], b0)
            elif ranges and a0 < ranges[-1][1]:
                continue  # the bias half of an augmented pair (same block, same group)
            else:
                ranges.append([a0, b0, gi])
        ranges[-1][1] = self.total
        new_ranges = [tuple(r) for r in ranges]
        if new_ranges != getattr(self, "ranges", None):
            self.ranges = new_ranges
            self.range_version = getattr(self, "range_version", 0) + 1  # captured graphs bake the launch ranges in
        # Whose moments are in M / V?  (1) this optimizer's own, already aliased -> keep; (2) tensors it brought along (a state restored
        # by optimizer.load_state_dict, or an optimizer that stepped elsewhere) -> copy them in; (3) an optimizer without state (fresh
        # torch.optim.Adam) -> start from zero moments and step 0 like torch does, NOT from a previous optimizer's buffers.
        same_optimizer = getattr(self, "_adopted", None) is optimizer
        host_step = torch.tensor(float(self.step.item())) if same_optimizer else torch.tensor(0.0)
        fresh = [name for name, p in named if "exp_avg" not in optimizer.state.get(p, {})]
        if len(fresh) == len(named) and not same_optimizer:
            self.M.zero_()
            self.V.zero_()
        for name, p in named:
            st = optimizer.state[p]
            mv, vv = self._view(self.M, name, p), self._view(self.V, name, p)
            if "exp_avg" in st and st["exp_avg"].data_ptr() != mv.data_ptr():
                mv.copy_(st["exp_avg"])
                vv.copy_(st["exp_avg_sq"])
                host_step = torch.as_tensor(st["step"]).detach().float().cpu().reshape(()).clone()
            elif "exp_avg" not in st and not same_optimizer and len(fresh) != len(named):
                mv.zero_()  # a parameter this optimizer has not seen yet
                vv.zero_()
            st["exp_avg"] = mv
            st["exp_avg_sq"] = vv
            st["step"] = host_step  # one shared host scalar, advanced by the engine
        self.step.fill_(int(host_step.item()))
        self._host_step = host_step
        self._adopted = optimizer
        self._adopted_quick = quick
        self.hyper_host = [None] * self.MAX_GROUPS

    def _state_aliased(self, optimizer: torch.optim.Optimizer) -> bool:
        """Cheap per-step sentinel (first / last parameter): is ``optimizer.state`` still the M / V views?  ``load_state_dict`` replaces
        the state tensors (the reference's CheckpointManager resume path), after which the fused Adam would keep updating buffers the
        optimizer no longer owns."""
        named = getattr(self, "_named_cache", None)
        if named is None or len(named) != len(self._plist):
            named = self._named_cache = list(self.module.named_parameters())
        for name, p in (named[0], named[-1]):
            st = optimizer.state.get(p)
            if not st or "exp_avg" not in st or st["exp_avg"].data_ptr() != self._view(self.M, name, p).data_ptr() or st.get("step") is not self._host_step:
                return False
        return True

    def sync_hyper(self, optimizer: torch.optim.Optimizer, grad_scale: float) -> None:
        for gi, g in enumerate(optimizer.param_groups):
            h = (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), float(grad_scale), 0.0, 0.0)
            if h != self.hyper_host[gi]:
                self.hyper[gi].copy_(torch.tensor(h, dtype=torch.float32), non_blocking=False)
                self.hyper_host[gi] = h

    def adam_ranges(self, a: int, b: int) -> List[Tuple[int, int, int]]:
        """(begin, end, group) pieces of [a, b) with uniform hyper-parameters."""
        rs = getattr(self, "ranges", None) or [(0, self.total, 0)]
        return [(max(a, ra), min(b, rb), g) for ra, rb, g in rs if max(a, ra) < min(b, rb)]

    def adam(self, a: int, b: int, advance: bool) -> None:
        """torch.optim.Adam.step over the flat range [a, b): one fused launch per param-group piece."""
        pieces = self.adam_ranges(a, b)
        for i, (ra, rb, g) in enumerate(pieces):
            ops.adam_step(self.P[ra:rb], self.G[ra:rb], self.M[ra:rb], self.V[ra:rb], self.Wb[ra:rb], self.hyper[g], self.step,
                          advance and i == len(pieces) - 1)


# =====================================================================================================================
# per-encoder plan
# =====================================================================================================================
class _BN(ops.BNBuffers):
    __slots__ = ("C", "scale", "shift", "bstat", "dgamma", "dbeta")

    def __init__(self):  # filled field by field in EncoderPlan._bn
        pass


class EncoderPlan:
    """Buffers + forward / backward closures of one ResNet encoder for a fixed (B, H, W)."""

    def __init__(self, fs: FlatState, enc: nn.Module, prefix: str, B: int, H: int, W: int, train: bool):
        self.fs, self.enc, self.prefix, self.B, self.H, self.W = fs, enc, prefix, B, H, W
        dev = fs.device
        self.fwd_train: List[Callable[[], None]] = []
        self.fwd_eval: List[Callable[[], None]] = []
        self.bwd: List[Callable[[], None]] = []
        self.x = torch.zeros(B, H, W, device=dev)        # original fp32 input (static)
        self.mask = torch.ones(B, device=dev)            # per-sample missing-modality mask
        self.pooled = torch.zeros(B, 512, device=dev)
        self.dpooled = torch.zeros(B, 512, device=dev)
        self.taps: Dict[str, torch.Tensor] = {}  # stored intermediates by name (NHWC bf16), for tests / inspection
        self.wgrad_stream: Optional[torch.cuda.Stream] = None  # set by the step plan
        self.wgrad_ws = ops.WgradScratch(dev)  # split partial sums of THIS encoder's weight-gradient launches (one stream)
        n_stat = sum(4 * ops.bn_stat_slots(m.num_features) * m.num_features for m in enc.modules() if isinstance(m, nn.BatchNorm2d))
        self.stat_arena = torch.zeros(n_stat, device=dev, dtype=torch.float64)  # zeroed once per step
        self._stat_off = 0
        self._build(train)

    # -- helpers -----------------------------------------------------------------------------------------------
    def _bn(self, name: str, C: int) -> _BN:
        fs, dev = self.fs, self.fs.device
        bn = _BN()
        bn.C = C
        bn.gamma = fs.flat_slice(fs.P, f"{self.prefix}{name}.weight")
        bn.beta = fs.flat_slice(fs.P, f"{self.prefix}{name}.bias")
        bn.dgamma = fs.flat_slice(fs.G, f"{self.prefix}{name}.weight")
        bn.dbeta = fs.flat_slice(fs.G, f"{self.prefix}{name}.bias")
        o = fs.buf_offsets[f"{self.prefix}{name}.running_mean"]
        bn.rmean = fs.S[o:o + C]
        o = fs.buf_offsets[f"{self.prefix}{name}.running_var"]
        bn.rvar = fs.S[o:o + C]
        bn.scale, bn.shift, bn.mean, bn.invstd = (torch.zeros(C, device=dev) for _ in range(4))
        # fp64 accumulators: forward (sum x, sum x^2) filled by the conv epilogue, backward (sum g, sum g*xhat)
        o = self._stat_off
        n = ops.bn_stat_slots(C) * 2 * C
        self._stat_off += 2 * n
        bn.stats = self.stat_arena[o:o + n]
        bn.bstat = self.stat_arena[o + n:o + 2 * n]
        return bn

    def _act(self, *shape) -> torch.Tensor:
        return torch.zeros(*shape, device=self.fs.device, dtype=BF16)

    def _offload(self, fn) -> None:
        """Run a weight-gradient kernel off the dependency chain.  dgrad -> BN backward -> dgrad is the critical chain of the
        backward pass; the wgrads only consume what the chain has already produced, so they go to a separate (normal
        priority) stream where their tensor-pipe work overlaps the HBM-bound BN kernels.  Joined by ``join_offload``."""
        ws = self.wgrad_stream
        if ws is None:
            fn()
            return
        ws.wait_stream(torch.cuda.current_stream(self.fs.device))
        with torch.cuda.stream(ws):
            fn()

    def join_offload(self) -> None:
        if self.wgrad_stream is not None:
            torch.cuda.current_stream(self.fs.device).wait_stream(self.wgrad_stream)

    def _build(self, train: bool) -> None:
        fs, B, dev, pre = self.fs, self.B, self.fs.device, self.prefix
        F, E, Bk = self.fwd_train, self.fwd_eval, self.bwd
        bwd_stack: List[Callable[[], None]] = []  # appended in forward order, executed reversed
        bwd_stack_names: List[str] = []
        # ---------------- stem: conv7x7/s2 (+mask) -> BN -> ReLU -> maxpool3/s2
        P0, Q0 = (self.H - 1) // 2 + 1, (self.W - 1) // 2 + 1
        P1, Q1 = (P0 - 1) // 2 + 1, (Q0 - 1) // 2 + 1
        w_stem = fs.flat_slice(fs.P, pre + "conv1.weight")
        dw_stem = fs.flat_slice(fs.G, pre + "conv1.weight")
        raw0 = self._act(B, P0, Q0, 64)  # the post-ReLU activation is never materialised (fused BN+ReLU+maxpool)
        pool, amax = self._act(B, P1, Q1, 64), torch.zeros(B, P1, Q1, 64, device=dev, dtype=torch.uint8)
        self.taps.update({"conv1": raw0, "maxpool": pool})
        bn0 = self._bn("bn1", 64)
        rows0 = B * P0 * Q0
        x, mask = self.x, self.mask
        F.append(lambda: ops.stem_fprop(x, mask, w_stem, raw0, bn0.stats))
        F.append(lambda: ops.stem_bn_pool_fwd(raw0, bn0, None, None, pool, amax, B, P0, Q0, 64, True, BN_MOMENTUM, BN_EPS))
        E.append(lambda: ops.stem_fprop(x, mask, w_stem, raw0, None))
        E.append(lambda: ops.bn_eval_coeffs(64, bn0.gamma, bn0.beta, bn0.rmean, bn0.rvar, BN_EPS, bn0.scale, bn0.shift))
        E.append(lambda: ops.stem_bn_pool_fwd(raw0, bn0, bn0.scale, bn0.shift, pool, amax, B, P0, Q0, 64, False))
        # backward of the stem is emitted last (see end of _build); needs the two gradients of `pool`
        cur, curH, curW, curC = pool, P1, Q1, 64
        grads_of_cur: List[Optional[torch.Tensor]] = [None, None]  # filled by the first block's backward
        stem_grad_slots = grads_of_cur
        # ---------------- residual stages
        for blk_i, (bname, blk) in enumerate(self._named_blocks()):
            inC, outC, stride = blk.conv1.in_channels, blk.conv1.out_channels, blk.conv1.stride[0]
            has_ds = blk.downsample is not None
            oH, oW = ops.conv_out_hw(curH, curW, 3, 3, stride, 1)
            g1 = ops.make_geom(B, curH, curW, inC, outC, 3, 3, stride, 1)
            g2 = ops.make_geom(B, oH, oW, outC, outC, 3, 3, 1, 1)
            rows = B * oH * oW
            n1, n2 = f"{pre}{bname}.conv1.weight", f"{pre}{bname}.conv2.weight"
            w1, w2 = fs.flat_slice(fs.Wb, n1), fs.flat_slice(fs.Wb, n2)
            w1t, w2t = w1, w2  # dgrad reads the same K,R,S,C bf16 weights (MN-major B operand)
            dw1, dw2 = fs.flat_slice(fs.G, n1), fs.flat_slice(fs.G, n2)
            raw1, a1, raw2, out = (self._act(B, oH, oW, outC) for _ in range(4))
            self.taps.update({f"{bname}.conv1": raw1, f"{bname}.relu1": a1, f"{bname}.conv2": raw2, bname: out})
            bn1, bn2 = self._bn(f"{bname}.bn1", outC), self._bn(f"{bname}.bn2", outC)
            xin = cur
            if has_ds:
                gd = ops.make_geom(B, curH, curW, inC, outC, 1, 1, stride, 0)
                nd = f"{pre}{bname}.downsample.0.weight"
                wd, dwd = fs.flat_slice(fs.Wb, nd), fs.flat_slice(fs.G, nd)
                wdt = wd
                rawd = self._act(B, oH, oW, outC)
                self.taps[f"{bname}.downsample"] = rawd
                bnd = self._bn(f"{bname}.downsample.1", outC)

            bnd_ = bnd if has_ds else None
            rawd_ = rawd if has_ds else None

            def fwd_train(g1=g1, g2=g2, xin=xin, w1=w1, w2=w2, raw1=raw1, a1=a1, raw2=raw2, out=out, bn1=bn1, bn2=bn2, rows=rows, outC=outC,
                          has_ds=has_ds, gd=gd if has_ds else None, wd=wd if has_ds else None, rawd=rawd_, bnd=bnd_):
                steps = [lambda: ops.conv_fprop(g1, xin, w1, raw1, bn1.stats),
                         lambda: ops.bn_train_fwd(raw1, bn1, None, None, a1, rows, outC, True, BN_MOMENTUM, BN_EPS),
                         lambda: ops.conv_fprop(g2, a1, w2, raw2, bn2.stats)]
                if has_ds:
                    steps.append(lambda: ops.conv_fprop(gd, xin, wd, rawd, bnd.stats))
                    steps.append(lambda: ops.bn_train_fwd(raw2, bn2, rawd, bnd, out, rows, outC, True, BN_MOMENTUM, BN_EPS))
                else:
                    steps.append(lambda: ops.bn_train_fwd(raw2, bn2, xin, None, out, rows, outC, True, BN_MOMENTUM, BN_EPS))
                return steps

            def fwd_eval(g1=g1, g2=g2, xin=xin, w1=w1, w2=w2, raw1=raw1, a1=a1, raw2=raw2, out=out, bn1=bn1, bn2=bn2, rows=rows, outC=outC,
                         has_ds=has_ds, gd=gd if has_ds else None, wd=wd if has_ds else None, rawd=rawd_, bnd=bnd_):
                steps = [lambda: ops.conv_fprop(g1, xin, w1, raw1, None),
                         lambda: ops.bn_eval_coeffs(outC, bn1.gamma, bn1.beta, bn1.rmean, bn1.rvar, BN_EPS, bn1.scale, bn1.shift),
                         lambda: ops.bn_act_fwd(raw1, bn1.scale, bn1.shift, None, None, None, a1, rows, outC, True),
                         lambda: ops.conv_fprop(g2, a1, w2, raw2, None),
                         lambda: ops.bn_eval_coeffs(outC, bn2.gamma, bn2.beta, bn2.rmean, bn2.rvar, BN_EPS, bn2.scale, bn2.shift)]
                if has_ds:
                    steps.append(lambda: ops.conv_fprop(gd, xin, wd, rawd, None))
                    steps.append(lambda: ops.bn_eval_coeffs(outC, bnd.gamma, bnd.beta, bnd.rmean, bnd.rvar, BN_EPS, bnd.scale, bnd.shift))
                    steps.append(lambda: ops.bn_act_fwd(raw2, bn2.scale, bn2.shift, rawd, bnd.scale, bnd.shift, out, rows, outC, True))
                else:
                    steps.append(lambda: ops.bn_act_fwd(raw2, bn2.scale, bn2.shift, xin, None, None, out, rows, outC, True))
                return steps

            F.extend(fwd_train())
            E.extend(fwd_eval())

            # ---------------- backward of this block (closures run in reverse block order)
            if train:
                d_raw2, d_a1, d_raw1 = (self._act(B, oH, oW, outC) for _ in range(3))
                d_x_main = self._act(B, curH, curW, inC)
                g_skip = self._act(B, oH, oW, outC)  # g = (dy1 + dy2) * (out > 0): gradient of the skip path AND input of bn2's pass 2
                out_grads: List[Optional[torch.Tensor]] = [None, None]  # set by the consumer (next block / avgpool)
                if has_ds:
                    d_rawd, d_x_ds = self._act(B, oH, oW, outC), self._act(B, curH, curW, inC)

                def bwd_block(g1=g1, g2=g2, xin=xin, a1=a1, raw1=raw1, raw2=raw2, out=out, bn1=bn1, bn2=bn2, rows=rows, outC=outC,
                              d_raw2=d_raw2, d_a1=d_a1, d_raw1=d_raw1, d_x_main=d_x_main, g_skip=g_skip,
                              out_grads=out_grads, w1t=w1t, w2t=w2t, dw1=dw1, dw2=dw2, has_ds=has_ds,
                              ds=(gd, wdt, dwd, rawd, bnd, d_rawd, d_x_ds) if has_ds else None):
                    dy1, dy2 = out_grads
                    # bn2 (+ residual add + ReLU): pass 1 stores g once; every later consumer reads g instead of (dy1, dy2, out)
                    ops.bn_bwd_reduce(dy1, dy2, out, raw2, bn2.mean, bn2.invstd, bn2.bstat, g_skip, rows, outC, True)
                    ops.bn_bwd_apply(g_skip, raw2, bn2.mean, bn2.invstd, bn2.gamma, bn2.bstat, bn2.dgamma, bn2.dbeta, d_raw2, rows, outC)
                    if has_ds:
                        gd_, wdt_, dwd_, rawd_, bnd_, d_rawd_, d_x_ds_ = ds
                        ops.bn_bwd_reduce(g_skip, None, None, rawd_, bnd_.mean, bnd_.invstd, bnd_.bstat, None, rows, outC, False)
                        ops.bn_bwd_apply(g_skip, rawd_, bnd_.mean, bnd_.invstd, bnd_.gamma, bnd_.bstat, bnd_.dgamma, bnd_.dbeta, d_rawd_, rows, outC)
                        self._offload(lambda: ops.conv_wgrad(gd_, xin, d_rawd_, dwd_, self.wgrad_ws))
                        ops.conv_dgrad(gd_, d_rawd_, wdt_, d_x_ds_)
                    self._offload(lambda: ops.conv_wgrad(g2, a1, d_raw2, dw2, self.wgrad_ws))
                    ops.conv_dgrad(g2, d_raw2, w2t, d_a1)
                    # bn1 + ReLU (g overwrites d_a1 in place)
                    ops.bn_bwd_reduce(d_a1, None, a1, raw1, bn1.mean, bn1.invstd, bn1.bstat, d_a1, rows, outC, True)
                    ops.bn_bwd_apply(d_a1, raw1, bn1.mean, bn1.invstd, bn1.gamma, bn1.bstat, bn1.dgamma, bn1.dbeta, d_raw1, rows, outC)
                    self._offload(lambda: ops.conv_wgrad(g1, xin, d_raw1, dw1, self.wgrad_ws))
                    ops.conv_dgrad(g1, d_raw1, w1t, d_x_main)

                bwd_stack.append(bwd_block)
                bwd_stack_names.append(bname)
                # the gradient of this block's INPUT arrives over two paths
                grads_of_cur[0] = d_x_main
                grads_of_cur[1] = d_x_ds if has_ds else g_skip
                grads_of_cur = out_grads
            cur, curH, curW, curC = out, oH, oW, outC
        # ---------------- global average pool
        HW = curH * curW
        last = cur
        pooled = self.pooled
        for L in (F, E):
            L.append(lambda: ops.avgpool_fwd(last, pooled, B, HW, 512))
        if train:
            d_last = self._act(B, curH, curW, 512)
            grads_of_cur[0] = d_last
            grads_of_cur[1] = None
            dpooled = self.dpooled
            Bk.append(lambda: ops.avgpool_bwd(dpooled, d_last, B, HW, 512))
            Bk.extend(reversed(bwd_stack))
            self.bwd_names = ["avgpool"] + list(reversed(bwd_stack_names)) + ["stem"]  # one name per closure of ``bwd``
            # stem backward
            d_raw0 = self._act(B, P0, Q0, 64)
            ws = torch.zeros(max(ops.stem_wgrad_workspace(x) // 4, 4), device=dev)

            import os as _os
            fold = _os.environ.get("MML_STEM_FOLD", "1") == "1"

            def bwd_stem():
                # The stem's dx has ONE consumer, its own weight gradient, and the stem output is linear in the input patches: pass 2 of the
                # BatchNorm backward (read g + raw, write dx: 306 MB at B = 256 / 112x112) is folded into the wgrad algebraically
                # (csrc/stem.cu); MML_STEM_FOLD=0 keeps the two-pass form for A/B runs
                ops.stem_bn_pool_bwd(stem_grad_slots[0], stem_grad_slots[1], amax, raw0, bn0, bn0.bstat, bn0.dgamma, bn0.dbeta, d_raw0, B, P0, Q0, 64,
                                     apply=not fold)
                if fold:
                    ops.stem_wgrad_bn(x, mask, d_raw0, w_stem, bn0, bn0.bstat, bn0.dgamma, bn0.dbeta, dw_stem, ws)
                else:
                    ops.stem_wgrad(x, mask, d_raw0, dw_stem, ws)
                self.join_offload()

            Bk.append(bwd_stem)

    def _named_blocks(self):
        for li, stage in enumerate((self.enc.layer1, self.enc.layer2, self.enc.layer3, self.enc.layer4)):
            for bi, blk in enumerate(stage):
                yield f"layer{li + 1}.{bi}", blk


# =====================================================================================================================
# stand-alone encoder (ResNetEncoder.forward outside the fusion model)
# =====================================================================================================================
class StandaloneEncoder:
    """Encoder-only forward.  When the encoder is a sub-module of a fusion model that already owns a FlatState
    (``owner`` = (engine, prefix)), that storage is shared instead of re-homing the parameters."""

    def __init__(self, enc: nn.Module):
        self.enc = enc
        self.fs: Optional[FlatState] = None
        self.plans: Dict[Tuple, EncoderPlan] = {}

    def forward(self, x: torch.Tensor, training: bool) -> torch.Tensor:
        owner = getattr(self.enc, "_mml_owner", None)
        eng = owner[0]() if owner is not None else None
        if eng is not None and eng.device == x.device:
            fs, prefix = eng.fs, owner[1]
        else:
            if self.fs is None or self.fs.device != x.device:
                self.fs = FlatState(self.enc, x.device)
                self.plans.clear()
            fs, prefix = self.fs, ""
        fs.ensure_fresh()
        key = (id(fs),) + tuple(x.shape)
        plan = self.plans.get(key)
        if plan is None:
            plan = self.plans[key] = EncoderPlan(fs, self.enc, prefix, x.shape[0], x.shape[1], x.shape[2], train=False)
            plan.emb = torch.zeros(x.shape[0], self.enc.hidden_dim, device=x.device)
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
        ops.linear_fwd(plan.pooled, fs.flat_slice(fs.P, prefix + "fc.weight").view(self.enc.hidden_dim, 512),
                       fs.flat_slice(fs.P, prefix + "fc.bias"), plan.emb)
        return plan.emb.clone()


# =====================================================================================================================
# the fused late-fusion step
# =====================================================================================================================
class LateFusionEngine:
    """forward / train step / eval step of AVMNIST(audio_encoder, image_encoder, head) on one GPU."""

    def __init__(self, model: nn.Module, device: torch.device, dropout_p: float, seed: Optional[int] = None):
        self.model = model
        self.device = device
        self.dropout_p = float(dropout_p)
        self.client_id = int(getattr(model, "_mml_client_id", 0))
        self.seed = ops.engine_seed(self.client_id) if seed is None else seed
        self.fwd_calls = 0  # forward() calls in train mode (they do not advance the optimizer step the kernel mixes into the seed)
        self.fs = FlatState(model, device)
        self.plans: Dict[Tuple, "_StepPlan"] = {}
        self.world = 1
        self.allreduce: Optional[Callable[["_StepPlan"], None]] = None
        self.allreduce_range: Optional[Callable] = None  # (a, b, producer streams, update, join) -- set by dist.DataParallel
        self.use_graphs = True

    def plan_for(self, B: int, aH: int, aW: int, iH: int, iW: int) -> "_StepPlan":
        key = (B, aH, aW, iH, iW)
        plan = self.plans.get(key)
        if plan is None:
            plan = self.plans[key] = _StepPlan(self, B, aH, aW, iH, iW)
        return plan


def _late_image_allreduce() -> bool:
    import os as _os

    return _os.environ.get("MML_IMAGE_AR_LATE", "1") == "1"


class _StepPlan:
    def __init__(self, eng: LateFusionEngine, B: int, aH: int, aW: int, iH: int, iW: int):
        self.eng, self.B = eng, B
        fs, dev, model = eng.fs, eng.device, eng.model
        self.audio = EncoderPlan(fs, model.audio_encoder, "audio_encoder.", B, aH, aW, train=True)
        self.image = EncoderPlan(fs, model.image_encoder, "image_encoder.", B, iH, iW, train=True)
        names = ["audio_encoder.fc.weight", "audio_encoder.fc.bias", "image_encoder.fc.weight", "image_encoder.fc.bias",
                 "net.0.weight", "net.0.bias", "net.3.weight", "net.3.bias", "net.5.weight", "net.5.bias"]
        params = dict(model.named_parameters())
        self.hp = ops.head_params(*[fs.flat_slice(fs.P, n).view(params[n].shape) for n in names])
        self.hg = ops.head_grads(*[fs.flat_slice(fs.G, n).view(params[n].shape) for n in names])
        self._keep = [fs.flat_slice(fs.P, n) for n in names]
        ps = ops.head_scratch_per_sample(self.hp)
        self.scratch = torch.zeros(B, ps, device=dev)
        self.H1 = params["net.0.weight"].shape[0]
        self.NC = params["net.5.weight"].shape[0]
        self.labels = torch.zeros(B, device=dev, dtype=torch.int64)
        self.logits = torch.zeros(B, self.NC, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.pred = torch.zeros(B, device=dev, dtype=torch.int32)
        self.drop_mask = torch.ones(B, self.H1, device=dev, dtype=torch.uint8)
        self.drop_given = False
        # pinned host mirrors for the per-step D2H of (loss, predictions)
        self.h_loss = torch.zeros(1).pin_memory()
        self.h_pred = torch.zeros(B, dtype=torch.int32).pin_memory()
        self.h_logits = torch.zeros(B, self.NC).pin_memory()
        self.graph_train = None          # (forward graph, backward + update graph)
        self.graph_train_nodrop = None
        self.graph_eval: Optional[torch.cuda.CUDAGraph] = None
        self.loss_ready = torch.cuda.Event()
        self.want_pred = False
        self.eager_steps = 0
        self.launches_per_step = 0
        self.side_stream: Optional[torch.cuda.Stream] = None
        import os as _os
        # schedule variants (A/B-tested on B200): defaults are the measured best
        self.tune = {"adam_split": _os.environ.get("MML_ADAM_SPLIT", "1") == "1", "head_side": _os.environ.get("MML_HEAD_SIDE", "0") == "1",
                     "skip": _os.environ.get("MML_SKIP_ENCODER", ""), "side_prio": _os.environ.get("MML_SIDE_PRIO", "-1")}
        # the audio encoder's persistent conv kernels leave a few SMs to the image encoder's stream (measured on the round-2 build: 0 -> 2.64 ms,
        # 16 -> 2.57, 32 -> 2.53, 40 / 48 -> 2.53)
        self.reserve_sms = int(_os.environ.get("MML_RESERVE_SMS", "32"))
        self.pdl_mode = _os.environ.get("MML_PDL_MODE", "none")  # none | image | audio | all
        ops.set_pdl(dev.index, self.pdl_mode != "none")
        if _os.environ.get("MML_WGRAD_STREAMS", "1") == "1":
            self.audio.wgrad_stream = torch.cuda.Stream(device=dev)
            self.image.wgrad_stream = torch.cuda.Stream(device=dev)
        img = [o for n, o in fs.offsets.items() if n.startswith("image_encoder.")]
        self.param_split = min(img) if img else 0  # flat ranges: [0, split) audio encoder, [split, total) image encoder + head
        # [audio_mid, split) = audio layer3 .. fc: 94 % of the audio encoder's parameters, complete once layer3's backward is
        # done; its all-reduce / Adam then runs under the backward of layer2 / layer1 / stem (MML_AUDIO_MID=0: one audio range)
        mid_name = "audio_encoder.layer3.0.conv1.weight"
        self.audio_mid = fs.offsets.get(mid_name, 0) if _os.environ.get("MML_AUDIO_MID", "1") == "1" and self.tune["adam_split"] else 0
        self.mid_stream: Optional[torch.cuda.Stream] = None
        # same idea on the image side: ResNet34's layer4 + fc + the head are 2/3 of that range and their gradients are complete after
        # the FIRST three blocks of the image backward, so their all-reduce starts ~0.5 ms earlier (MML_IMAGE_MID=0: one image range)
        self.image_mid = fs.offsets.get("image_encoder.layer4.0.conv1.weight", 0) if _os.environ.get("MML_IMAGE_MID", "1") == "1" and self.tune["adam_split"] else 0

    # -- schedules -----------------------------------------------------------------------------------------------
    def _use_dropout(self) -> bool:
        return self.eng.dropout_p > 0.0

    def _side(self) -> torch.cuda.Stream:
        if self.side_stream is None:
            # High priority: the image encoder is a long serial chain of tiny kernels (a few CTAs each); letting them jump
            # the queue of the audio encoder's large grids keeps that chain off the critical path at almost no cost.
            prio = int(self.tune.get("side_prio", -1))
            self.side_stream = torch.cuda.Stream(device=self.eng.device, priority=prio)
        return self.side_stream

    def _both_encoders(self, audio_ops, image_ops, after_image=None, after_image_late=None) -> None:
        """Audio encoder on the current stream, image encoder concurrently on a side stream (fork / join).

        ResNet34 on 28x28 images is ~360 tiny latency-bound launches (1..128 CTAs each); run alone they cost more
        wall time than the 7x bigger audio encoder.  Forked onto a second stream they fill SMs the audio kernels leave idle.
        """
        skip = self.tune.get("skip", "")  # timing experiments only (results are then meaningless)
        if skip == "image":
            image_ops = []
        elif skip == "audio":
            audio_ops = []
        main = torch.cuda.current_stream(self.eng.device)
        side = self._side()
        side.wait_stream(main)
        idx = self.eng.device.index
        # programmatic dependent launch per stream (measured on B200): it shortens the image encoder's chain of small
        # dependent kernels (-7.5 % alone), but pre-launched CTAs of the audio encoder's large grids hold SMs the image stream needs
        pdl_image, pdl_audio = self.pdl_mode in ("image", "all"), self.pdl_mode in ("audio", "all")
        try:
            ops.set_pdl(idx, pdl_image)
            with torch.cuda.stream(side):
                for op in image_ops:
                    op()
                if after_image is not None:
                    after_image()
            ops.set_pdl(idx, pdl_audio)
            if self.reserve_sms > 0 and image_ops:
                ops.set_sm_budget(idx, ops._ctx_sm_count(idx) - self.reserve_sms)
            for op in audio_ops:
                op()
            if after_image_late is not None:
                # issued AFTER the audio encoder's work: collectives of one communicator run in issue order, and the all-reduce of the
                # audio mid range (ready early) must not queue behind one that waits for the end of the image backward
                with torch.cuda.stream(side):
                    after_image_late()
        finally:
            ops.set_sm_budget(idx, 0)
            ops.set_pdl(idx, self.pdl_mode != "none")
        main.wait_stream(side)

    def run_train(self, own_dropout: bool) -> None:
        self.run_train_fwd(own_dropout)
        self.run_train_bwd()

    def run_train_fwd(self, own_dropout: bool) -> None:
        """Forward half of the step: mask -> both encoders -> head -> loss / predictions (everything ``train_step`` returns)."""
        eng, fs = self.eng, self.eng.fs
        p = eng.dropout_p
        # no fs.G.zero_(): every producer of a gradient (conv / stem weight gradients, BatchNorm backward, head) STORES its
        # result; the alignment padding between tensors is never written and stays zero
        self.audio.stat_arena.zero_()
        self.image.stat_arena.zero_()
        if self._use_dropout() and own_dropout:
            ops.dropout_mask(self.drop_mask, p, eng.seed, fs.step)
        self._both_encoders(self.audio.fwd_train, self.image.fwd_train)
        dm = self.drop_mask if self._use_dropout() else None
        scale = 1.0 / (1.0 - p) if self._use_dropout() else 1.0
        ops.head_fwd(self.hp, self.audio.pooled, self.image.pooled, self.labels, dm, scale, self.scratch, self.logits, self.loss, self.pred)

    def run_train_bwd(self) -> None:
        eng, fs = self.eng, self.eng.fs
        p = eng.dropout_p
        dm = self.drop_mask if self._use_dropout() else None
        scale = 1.0 / (1.0 - p) if self._use_dropout() else 1.0
        ops.head_bwd(self.hp, self.hg, self.audio.pooled, self.image.pooled, self.labels, dm, scale, self.scratch, 1.0,
                     self.audio.dpooled, self.image.dpooled, phases=1)

        def head_weight_grads():  # needs only the per-sample deltas: runs on the side stream, off the audio critical path
            ops.head_bwd(self.hp, self.hg, self.audio.pooled, self.image.pooled, self.labels, dm, scale, self.scratch, 1.0,
                         self.audio.dpooled, self.image.dpooled, phases=2)
        # The image encoder holds 2/3 of the parameters but few FLOPs.  As soon as its backward is done (side stream) its
        # gradient range (image encoder + head) is all-reduced (DP) and its Adam update runs -- all under the audio
        # encoder's backward.  The audio range follows on the main stream and advances the step counter.
        split = self.param_split

        image_bwd = self.image.bwd
        img_hi = self.image_mid if (self.image_mid > split and eng.allreduce_range is not None) else 0

        def finish_image_range():
            if img_hi:
                # what is left of the image range: conv1 .. layer3 (producers: this stream + the image wgrad stream, joined by the stem)
                eng.allreduce_range(split, img_hi, [torch.cuda.current_stream(eng.device)], update=lambda: self._adam_range(split, img_hi, False))
            elif eng.allreduce is not None:
                eng.allreduce(self, 0, update=lambda: self._adam_range(split, fs.total, False))
            else:
                self._adam_range(split, fs.total, False)

        if img_hi:
            def finish_image_hi():
                # layer4 + fc + head: BatchNorm gradients on this stream; conv / head weight gradients on the image wgrad stream
                cur = torch.cuda.current_stream(eng.device)
                producers = [st for st in (cur, self.image.wgrad_stream) if st is not None]
                eng.allreduce_range(img_hi, fs.total, producers, update=lambda: self._adam_range(img_hi, fs.total, False))

            k = self.image.bwd_names.index("layer4.0") + 1
            image_bwd = self.image.bwd[:k] + [finish_image_hi] + self.image.bwd[k:]

        audio_bwd = self.audio.bwd
        if self.audio_mid > 0 and eng.allreduce_range is None:
            self.audio_mid = 0  # single GPU: nothing to hide (measured: no gain, and one more stream to alias with the copies)
        if self.audio_mid > 0:
            mid_a = self.audio_mid

            def finish_audio_mid():
                # producers of that range's gradients: BatchNorm gradients on this stream, conv weight gradients on the audio
                # wgrad stream, the audio fc gradients from the head's weight-gradient launch (image wgrad stream)
                cur = torch.cuda.current_stream(eng.device)
                producers = [st for st in (cur, self.audio.wgrad_stream, self.image.wgrad_stream) if st is not None]
                if eng.allreduce_range is not None:
                    eng.allreduce_range(mid_a, split, producers, update=lambda: self._adam_range(mid_a, split, False))
                else:
                    if self.mid_stream is None:
                        self.mid_stream = torch.cuda.Stream(device=eng.device)
                    for st in producers:
                        self.mid_stream.wait_stream(st)
                    with torch.cuda.stream(self.mid_stream):
                        self._adam_range(mid_a, split, False)

            k = self.audio.bwd_names.index("layer3.0") + 1
            audio_bwd = self.audio.bwd[:k] + [finish_audio_mid] + self.audio.bwd[k:]
        if self.tune["head_side"]:
            self._both_encoders(audio_bwd, [head_weight_grads] + image_bwd, after_image=finish_image_range if self.tune["adam_split"] else None)
        else:
            # the head's weight gradients are off the chain too: they share the image encoder's wgrad stream, which is joined
            # before that range's all-reduce / Adam
            if self.tune.get("skip", "") == "image":
                head_weight_grads()  # timing experiment without the image encoder: nothing would join its wgrad stream
            else:
                self.image._offload(head_weight_grads)
            late = eng.allreduce_range is not None and _late_image_allreduce()
            self._both_encoders(audio_bwd, image_bwd, after_image=finish_image_range if self.tune["adam_split"] and not late else None,
                                after_image_late=finish_image_range if self.tune["adam_split"] and late else None)
        fs.NBT += 1

    def _adam_range(self, a: int, b: int, advance: bool) -> None:
        self.eng.fs.adam(a, b, advance)

    def run_update(self) -> None:
        eng, fs = self.eng, self.eng.fs
        if not self.tune["adam_split"]:
            if eng.allreduce is not None:
                eng.allreduce(self, 0)
                eng.allreduce(self, 1, update=lambda: self._adam_range(0, fs.total, True))
            else:
                self._adam_range(0, fs.total, True)
            return
        last = self.audio_mid if self.audio_mid > 0 else self.param_split  # what is left of the audio encoder: [0, last)
        if eng.allreduce_range is not None:
            eng.allreduce_range(0, last, [torch.cuda.current_stream(eng.device)], update=lambda: self._adam_range(0, last, True), join=True)
        else:
            if self.mid_stream is not None:
                torch.cuda.current_stream(eng.device).wait_stream(self.mid_stream)
            self._adam_range(0, last, True)

    def run_eval(self, with_loss: bool) -> None:
        self._both_encoders(self.audio.fwd_eval, self.image.fwd_eval)
        ops.head_fwd(self.hp, self.audio.pooled, self.image.pooled, self.labels if with_loss else None, None, 1.0, self.scratch, self.logits,
                     self.loss if with_loss else None, self.pred)

    def run_forward_train_mode(self) -> None:
        """forward() in train() mode without a step: batch statistics, running stats updated, dropout applied."""
        eng, fs = self.eng, self.eng.fs
        self.audio.stat_arena.zero_()
        self.image.stat_arena.zero_()
        if self._use_dropout():
            eng.fwd_calls += 1
            ops.dropout_mask(self.drop_mask, eng.dropout_p, ops.engine_seed(eng.client_id, eng.fwd_calls) ^ eng.seed, fs.step)
        self._both_encoders(self.audio.fwd_train, self.image.fwd_train)
        dm = self.drop_mask if self._use_dropout() else None
        scale = 1.0 / (1.0 - eng.dropout_p) if self._use_dropout() else 1.0
        ops.head_fwd(self.hp, self.audio.pooled, self.image.pooled, None, dm, scale, self.scratch, self.logits, None, self.pred)
        fs.NBT += 1

    # -- execution (eager for the first steps, CUDA graphs afterwards) ----------------------------------------------------
    def publish(self) -> None:
        """Loss (and predictions, when a metric recorder wants them) to pinned host memory + an event, enqueued right after the FORWARD
        half.  ``train_step`` of the model waits for this event only, so the host returns -- and stages / launches the next step --
        while the GPU is still in the backward + optimizer half: the host work between two steps (~0.2 ms: prefetcher, checks, staging
        copies, graph launch) no longer leaves the GPU idle.  Everything later on the stream stays ordered behind the whole step."""
        self.h_loss.copy_(self.loss, non_blocking=True)
        if self.want_pred:
            self.h_pred.copy_(self.pred, non_blocking=True)
        self.loss_ready.record(torch.cuda.current_stream(self.eng.device))

    def train_step(self, given_dropout: bool) -> None:
        eng = self.eng
        if not eng.use_graphs:
            self.run_train_fwd(not given_dropout)
            self.publish()
            self.run_train_bwd()
            self.run_update()
            return
        attr = "graph_train_nodrop" if given_dropout else "graph_train"
        if getattr(self, "_range_version", None) != eng.fs.range_version:  # new optimizer grouping: the Adam launches changed
            self.graph_train = self.graph_train_nodrop = None
            self._range_version = eng.fs.range_version
        g = getattr(self, attr)
        if g is None:
            if self.eager_steps < 2:
                before = ops.launch_count(eng.device.index)
                self.run_train_fwd(not given_dropout)
                self.publish()
                self.run_train_bwd()
                self.run_update()
                self.launches_per_step = ops.launch_count(eng.device.index) - before
                self.eager_steps += 1
                return
            torch.cuda.synchronize(eng.device)
            g_fwd, g_bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_fwd):
                self.run_train_fwd(not given_dropout)
            with torch.cuda.graph(g_bwd):
                self.run_train_bwd()
                self.run_update()
            g = (g_fwd, g_bwd)
            setattr(self, attr, g)
        g[0].replay()
        self.publish()
        g[1].replay()

    def eval_step(self, with_loss: bool = True) -> None:
        self.run_eval(with_loss)
