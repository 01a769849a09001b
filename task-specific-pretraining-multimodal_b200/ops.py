"""Thin torch-tensor wrappers over the C ABI (include/mml_b200.h).

PyTorch is used here for device memory and streams only: every function checks dtype / device / contiguity, passes
raw device pointers and the CURRENT torch stream to libmml_b200.so and returns.  No function here computes anything
in PyTorch; if the library is missing the import of ``_lib`` raises -- there is no fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from ._lib import BN1dBwdDesc, BN1dDesc, ConvGeom, Context, HeadGrads, HeadParams, MMLError

BF16 = torch.bfloat16


def _ctx(t: torch.Tensor) -> Context:
    if not t.is_cuda:
        raise MMLError("mml_b200 ops need CUDA tensors (no CPU path)")
    cur = torch.cuda.current_device()
    idx = t.device.index if t.device.index is not None else cur
    if idx != cur:  # kernels launch on the CURRENT device; a stream / pointer of another GPU fails obscurely (invalid resource handle)
        raise MMLError(f"tensor lives on cuda:{idx} but the current device is cuda:{cur}: run the call under torch.cuda.device({idx})")
    return Context.get(idx)


def check_class_labels(labels, num_classes: int) -> None:
    """CrossEntropyLoss contract of the fused heads: every label in [0, num_classes).  The kernels implement neither ``ignore_index``
    rows (torch would drop them from the mean) nor torch's device-side assert for other out-of-range labels, so labels that arrive on
    the host (the DataLoader case) are validated here; device-resident labels are the caller's responsibility."""
    t = torch.as_tensor(labels)
    if t.is_cuda or t.numel() == 0:
        return
    lo, hi = t.aminmax()
    if int(lo) < 0 or int(hi) >= num_classes:
        raise ValueError(f"labels must lie in [0, {num_classes}) (got min {int(lo)}, max {int(hi)}); ignore_index rows are not supported by the fused loss")


def _stream(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t: Optional[torch.Tensor], dtype=None) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    if dtype is not None and t.dtype != dtype:
        raise MMLError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise MMLError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def bn_stat_slots(Cn: int) -> int:
    """Slots of a BatchNorm statistics accumulator: fp64 [slots, C, 2] (mml_bn_stat_slots: clamp(1024 / C, 2, 16))."""
    s = 1024 // Cn
    return 2 if s < 2 else (16 if s > 16 else s)


def bn_stats_buffer(Cn: int, device) -> torch.Tensor:
    return torch.zeros(bn_stat_slots(Cn), Cn, 2, device=device, dtype=torch.float64)


def conv_out_hw(H: int, W: int, R: int, S: int, stride: int, pad: int) -> Tuple[int, int]:
    return (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1


def make_geom(N, H, W, Cin, K, R, S, stride, pad) -> ConvGeom:
    return ConvGeom(int(N), int(H), int(W), int(Cin), int(K), int(R), int(S), int(stride), int(pad))


# ---- a1 ------------------------------------------------------------------------------------------------------
def mask_apply(x: torch.Tensor, mask: torch.Tensor, want_reverse: bool = False):
    """sample = original * mask  (data/base_dataset.py:71); x fp32 [B, ...], mask fp32 [B]."""
    ctx = _ctx(x)
    B = x.shape[0]
    per = x.numel() // max(B, 1)
    y = torch.empty_like(x)
    yr = torch.empty_like(x) if want_reverse else None
    ctx.check(ctx.lib.mml_mask_apply_f32(ctx.handle, _p(x, torch.float32), _p(mask, torch.float32), _p(y), _p(yr), B, per, _stream(x)), "mask_apply")
    return (y, yr) if want_reverse else y


def mask_apply_into(x: torch.Tensor, mask: torch.Tensor, y: torch.Tensor) -> None:
    """``mask_apply`` into a pre-allocated output (static buffers of a captured schedule)."""
    ctx = _ctx(x)
    B = x.shape[0]
    ctx.check(ctx.lib.mml_mask_apply_f32(ctx.handle, _p(x, torch.float32), _p(mask, torch.float32), _p(y, torch.float32), _p(None), B,
                                         x.numel() // max(B, 1), _stream(x)), "mask_apply")


# ---- f4: input staging (staging.cu) --------------------------------------------------------------------------
def missing_mask_draw(p_present: torch.Tensor, num_samples: int, seed: int, stream_id: int = 0, first_sample: int = 0,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [n_modalities, num_samples] of 0/1: sample ``first_sample + j`` of modality m is present iff its Philox4x32-10 uniform
    is below ``p_present[m]`` (base_dataset.py:46-59; bit pattern fixed by (seed, stream_id, m, sample) -- include/mml_b200.h)."""
    ctx = _ctx(p_present)
    n_mod = p_present.numel()
    if out is None:
        out = torch.empty(n_mod, num_samples, device=p_present.device)
    if out.shape != (n_mod, num_samples):
        raise MMLError(f"missing_mask_draw: out must be [{n_mod}, {num_samples}]")
    ctx.check(ctx.lib.mml_missing_mask_draw(ctx.handle, _p(p_present, torch.float32), _p(out, torch.float32), n_mod, int(first_sample),
                                            int(num_samples), int(num_samples), int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_id) & 0xFFFFFFFF,
                                            _stream(p_present)), "missing_mask_draw")
    return out


def missing_mask_gather(masks: torch.Tensor, sample_idx: torch.Tensor, out: Optional[torch.Tensor] = None,
                        bad_flag: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[m, b] = masks[m, sample_idx[b]] (the per-item mask lookup of data/avmnist.py:193-224, for a whole batch)."""
    ctx = _ctx(masks)
    n_mod, n = masks.shape
    B = sample_idx.numel()
    if out is None:
        out = torch.empty(n_mod, B, device=masks.device)
    ctx.check(ctx.lib.mml_missing_mask_gather(ctx.handle, _p(masks, torch.float32), _p(sample_idx, torch.int64), _p(out, torch.float32), n_mod, n, B,
                                              _p(bad_flag, torch.int32), _stream(masks)), "missing_mask_gather")
    return out


def u8_lut(src: torch.Tensor, lut: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 ``lut[src]`` for a uint8 tensor and a 256-entry fp32 table (data/avmnist.py:188-191 for uint8 pixels, see data.luma_lut)."""
    ctx = _ctx(src)
    if lut.numel() != 256:
        raise MMLError("u8_lut: the table must have 256 entries")
    if out is None:
        out = torch.empty(src.shape, device=src.device)
    if out.numel() != src.numel():
        raise MMLError("u8_lut: out must have as many elements as src")
    ctx.check(ctx.lib.mml_stage_u8_lut_f32(ctx.handle, _p(src, torch.uint8), _p(lut, torch.float32), _p(out, torch.float32), src.numel(),
                                           _stream(src)), "stage_u8_lut")
    return out


# ---- stem ----------------------------------------------------------------------------------------------------
def stem_fprop(x, mask, w, y, stats) -> None:
    """stats: ``bn_stats_buffer(64)`` accumulator (zeroed by the caller) or None."""
    ctx = _ctx(x)
    B, H, W = x.shape
    ctx.check(ctx.lib.mml_stem_fprop(ctx.handle, _p(x, torch.float32), _p(mask, torch.float32), _p(w, torch.float32), _p(y, BF16),
                                     _p(stats, torch.float64), B, H, W, _stream(x)), "stem_fprop")


def stem_wgrad_workspace(x) -> int:
    ctx = _ctx(x)
    B, H, W = x.shape
    return int(ctx.lib.mml_stem_wgrad_workspace(ctx.handle, B, H, W))


def stem_wgrad(x, mask, dy, dw, workspace) -> None:
    ctx = _ctx(x)
    B, H, W = x.shape
    ctx.check(ctx.lib.mml_stem_wgrad(ctx.handle, _p(x, torch.float32), _p(mask, torch.float32), _p(dy, BF16), _p(dw, torch.float32),
                                     _p(workspace, torch.float32), workspace.numel() * 4, B, H, W, _stream(x)), "stem_wgrad")


def stem_wgrad_bn(x, mask, g, w, bn: "BNBuffers", bstat, dgamma, dbeta, dw, workspace) -> None:
    """Stem weight gradient with the stem BatchNorm's backward pass 2 folded in: ``g`` = what ``stem_bn_pool_bwd(apply=False)`` left in dx."""
    ctx = _ctx(x)
    B, H, W = x.shape
    ctx.check(ctx.lib.mml_stem_wgrad_bn(ctx.handle, _p(x, torch.float32), _p(mask, torch.float32), _p(g, BF16), _p(w, torch.float32),
                                        _p(bstat, torch.float64), _p(bn.mean), _p(bn.invstd), _p(bn.gamma), _p(dgamma), _p(dbeta), _p(dw, torch.float32),
                                        _p(workspace, torch.float32), workspace.numel() * 4, B, H, W, _stream(x)), "stem_wgrad_bn")


# ---- conv ----------------------------------------------------------------------------------------------------
def conv_fprop(g: ConvGeom, x, w_krsc, y, stats=None) -> None:
    """stats: ``bn_stats_buffer(K)`` accumulator of (sum, sum of squares) of y (zeroed by the caller) or None."""
    ctx = _ctx(x)
    ctx.check(ctx.lib.mml_conv_fprop(ctx.handle, C.byref(g), _p(x, BF16), _p(w_krsc, BF16), _p(y, BF16), _p(stats, torch.float64),
                                     _stream(x)), "conv_fprop")


def conv_dgrad(g: ConvGeom, dy, w_krsc, dx) -> None:
    ctx = _ctx(dy)
    ctx.check(ctx.lib.mml_conv_dgrad(ctx.handle, C.byref(g), _p(dy, BF16), _p(w_krsc, BF16), _p(dx, BF16), _stream(dy)), "conv_dgrad")


def conv_wgrad_workspace(g: ConvGeom, device) -> int:
    """Bytes of scratch mml_conv_wgrad needs for this geometry (per-CTA / per-split partial sums, reduced in a fixed order)."""
    ctx = Context.get(torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device())
    n = int(ctx.lib.mml_conv_wgrad_workspace(ctx.handle, C.byref(g)))
    if n < 0:
        raise MMLError("conv_wgrad: unsupported geometry")
    return n


class WgradScratch:
    """Scratch for the split partial sums of ``conv_wgrad``: ONE per stream that issues weight-gradient launches (launches on a stream
    are ordered, so they can share it; two streams must not).  Grows on demand -- during the eager warm-up steps, i.e. before a
    CUDA-graph capture bakes the address in."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.buf: Optional[torch.Tensor] = None

    def ensure(self, nbytes: int) -> Optional[torch.Tensor]:
        if nbytes <= 0:
            return self.buf
        if self.buf is None or self.buf.numel() * 4 < nbytes:
            if torch.cuda.is_current_stream_capturing():
                raise MMLError("conv_wgrad scratch would have to grow inside a CUDA-graph capture (run an eager step first)")
            self.buf = torch.empty((nbytes + 3) // 4, device=self.device)
        return self.buf


def conv_wgrad(g: ConvGeom, x, dy, dw_krsc, workspace=None) -> None:
    """dw = sum dy * x (OVERWRITES dw; bit-reproducible).  workspace: a ``WgradScratch`` or an fp32 tensor of at least
    ``conv_wgrad_workspace(g)`` bytes, private to the current stream (may be None when the geometry needs none)."""
    ctx = _ctx(x)
    if isinstance(workspace, WgradScratch):
        key = (g.N, g.H, g.W, g.C, g.K, g.R, g.S, g.stride, g.pad)
        need = _WS_CACHE.get(key)
        if need is None:
            need = _WS_CACHE[key] = conv_wgrad_workspace(g, x.device)
        workspace = workspace.ensure(need)
    ctx.check(ctx.lib.mml_conv_wgrad(ctx.handle, C.byref(g), _p(x, BF16), _p(dy, BF16), _p(dw_krsc, torch.float32), _p(workspace, torch.float32),
                                     workspace.numel() * 4 if workspace is not None else 0, _stream(x)), "conv_wgrad")


_WS_CACHE: dict = {}


# ---- batch norm / activations / pooling ---------------------------------------------------------------------------
class BNBuffers:
    """Device pointers of one training-mode BatchNorm: fp64 stats, affine parameters, running and saved statistics."""

    __slots__ = ("stats", "gamma", "beta", "rmean", "rvar", "mean", "invstd")

    def __init__(self, stats, gamma, beta, rmean, rvar, mean, invstd):
        self.stats, self.gamma, self.beta, self.rmean, self.rvar, self.mean, self.invstd = stats, gamma, beta, rmean, rvar, mean, invstd


def bn_train_fwd(x, bn: "BNBuffers", res, rbn, y, rows, Cn, relu, momentum=0.1, eps=1e-5) -> None:
    """y = relu?(bn(x) [+ res | + rbn(res)]) with batch statistics taken from bn.stats (fp64 sums from the conv epilogue)."""
    ctx = _ctx(x)
    z = C.c_void_p(0)
    r = (_p(rbn.stats, torch.float64), _p(rbn.gamma), _p(rbn.beta), _p(rbn.rmean), _p(rbn.rvar), _p(rbn.mean), _p(rbn.invstd)) if rbn is not None else (z,) * 7
    ctx.check(ctx.lib.mml_bn_train_fwd(ctx.handle, _p(x, BF16), _p(bn.stats, torch.float64), _p(bn.gamma), _p(bn.beta), _p(bn.rmean), _p(bn.rvar),
                                       _p(bn.mean), _p(bn.invstd), _p(res), *r, _p(y, BF16), rows, Cn, int(relu), float(momentum), float(eps),
                                       _stream(x)), "bn_train_fwd")


def bn_eval_coeffs(Cn, gamma, beta, rmean, rvar, eps, scale, shift) -> None:
    ctx = _ctx(gamma)
    ctx.check(ctx.lib.mml_bn_eval_coeffs(ctx.handle, Cn, _p(gamma), _p(beta), _p(rmean), _p(rvar), eps, _p(scale), _p(shift), _stream(gamma)),
              "bn_eval_coeffs")


def bn_act_fwd(x, scale, shift, res, rscale, rshift, y, rows, Cn, relu) -> None:
    ctx = _ctx(x)
    ctx.check(ctx.lib.mml_bn_act_fwd(ctx.handle, _p(x, BF16), _p(scale), _p(shift), _p(res), _p(rscale), _p(rshift), _p(y, BF16), rows, Cn,
                                     int(relu), _stream(x)), "bn_act_fwd")


def bn_bwd_reduce(dy1, dy2, y, x, mean, invstd, bstat, g_out, rows, Cn, relu) -> None:
    """pass 1: g = (dy1 [+ dy2]) * (y > 0); bstat += (sum g, sum g*xhat); g_out (optional, may alias dy1) = g."""
    ctx = _ctx(dy1)
    ctx.check(ctx.lib.mml_bn_bwd_reduce(ctx.handle, _p(dy1, BF16), _p(dy2), _p(y), _p(x, BF16), _p(mean), _p(invstd), _p(bstat, torch.float64),
                                        _p(g_out), rows, Cn, int(relu), _stream(dy1)), "bn_bwd_reduce")


def bn_bwd_apply(g, x, mean, invstd, gamma, bstat, dgamma, dbeta, dx, rows, Cn) -> None:
    """pass 2: dx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)); dx may alias g; dgamma / dbeta written."""
    ctx = _ctx(g)
    ctx.check(ctx.lib.mml_bn_bwd_apply(ctx.handle, _p(g, BF16), _p(x, BF16), _p(mean), _p(invstd), _p(gamma), _p(bstat, torch.float64), _p(dgamma),
                                       _p(dbeta), _p(dx, BF16), rows, Cn, _stream(g)), "bn_bwd_apply")


def maxpool_fwd(x, y, argmax, N, H, W, Cn) -> None:
    ctx = _ctx(x)
    ctx.check(ctx.lib.mml_maxpool3x3s2_fwd(ctx.handle, _p(x, BF16), _p(y, BF16), _p(argmax, torch.uint8), N, H, W, Cn, _stream(x)), "maxpool_fwd")


def maxpool_bwd(dy, dy2, argmax, dx, N, H, W, Cn) -> None:
    ctx = _ctx(dy)
    ctx.check(ctx.lib.mml_maxpool3x3s2_bwd(ctx.handle, _p(dy, BF16), _p(dy2), _p(argmax, torch.uint8), _p(dx, BF16), N, H, W, Cn, _stream(dy)), "maxpool_bwd")


def stem_bn_pool_fwd(x, bn: "BNBuffers", scale, shift, y, argmax, N, H, W, Cn, train: bool, momentum=0.1, eps=1e-5) -> None:
    """y = maxpool3x3s2(relu(bn(x))); train: batch statistics from bn.stats, eval: precomputed scale / shift."""
    ctx = _ctx(x)
    z = C.c_void_p(0)
    if train:
        a = (_p(bn.stats, torch.float64), _p(bn.gamma), _p(bn.beta), _p(bn.rmean), _p(bn.rvar), _p(bn.mean), _p(bn.invstd), z, z)
    else:
        a = (z, _p(bn.gamma), _p(bn.beta), z, z, z, z, _p(scale), _p(shift))
    ctx.check(ctx.lib.mml_stem_bn_pool_fwd(ctx.handle, _p(x, BF16), *a, _p(y, BF16), _p(argmax, torch.uint8), N, H, W, Cn, float(momentum), float(eps),
                                           _stream(x)), "stem_bn_pool_fwd")


def stem_bn_pool_bwd(dy, dy2, argmax, x, bn: "BNBuffers", bstat, dgamma, dbeta, dx, N, H, W, Cn, apply: bool = True) -> None:
    """apply=False: stop after pass 1 (dx holds g, bstat the sums); ``stem_wgrad_bn`` then finishes the BatchNorm backward inside the wgrad."""
    ctx = _ctx(dy)
    ctx.check(ctx.lib.mml_stem_bn_pool_bwd(ctx.handle, _p(dy, BF16), _p(dy2), _p(argmax, torch.uint8), _p(x, BF16), _p(bn.mean), _p(bn.invstd),
                                           _p(bn.gamma), _p(bn.beta), _p(bstat, torch.float64), _p(dgamma), _p(dbeta), _p(dx, BF16), N, H, W, Cn,
                                           1 if apply else 0, _stream(dy)), "stem_bn_pool_bwd")


# ---- ConvBlock encoders (MNISTAudio / MNISTImage, models/avmnist.py:34-185) -----------------------------------------------
def conv3x3_c1_fprop(x, mask, w, y, stats, K: int) -> None:
    """Conv2d(1, K, 3, padding=1) of (x * mask); y NHWC bf16 [B,H,W,64] (channels >= K zero)."""
    B, H, W = x.shape
    c = _ctx(x)
    c.check(c.lib.mml_conv3x3_c1_fprop(c.handle, _p(x, torch.float32), _p(mask, torch.float32), _p(w, torch.float32), _p(y, torch.bfloat16),
                                       _p(stats, torch.float64), B, H, W, int(K), _stream(x)), "mml_conv3x3_c1_fprop")


def conv3x3_c1_wgrad_workspace(x, K: int) -> int:
    c = _ctx(x)
    return int(c.lib.mml_conv3x3_c1_wgrad_workspace(c.handle, x.shape[0], x.shape[1], int(K)))


def conv3x3_c1_wgrad(x, mask, dy, dw, workspace, K: int) -> None:
    B, H, W = x.shape
    c = _ctx(x)
    c.check(c.lib.mml_conv3x3_c1_wgrad(c.handle, _p(x, torch.float32), _p(mask, torch.float32), _p(dy, torch.bfloat16), _p(dw, torch.float32),
                                       _p(workspace, torch.float32), workspace.numel() * 4, B, H, W, int(K), _stream(x)), "mml_conv3x3_c1_wgrad")


def maxpool_k_fwd(x, y, y_flat, argmax, k: int) -> None:
    """nn.MaxPool2d(k): x NHWC bf16 -> y NHWC bf16 and / or y_flat fp32 [B, C*P*Q] (nn.Flatten order)."""
    B, H, W, Cn = x.shape
    c = _ctx(x)
    c.check(c.lib.mml_maxpool_k_fwd(c.handle, _p(x, torch.bfloat16), _p(y, torch.bfloat16), _p(y_flat, torch.float32), _p(argmax, torch.uint8),
                                    B, H, W, Cn, int(k), _stream(x)), "mml_maxpool_k_fwd")


def maxpool_k_bwd(dy, dy_flat, argmax, dx, k: int) -> None:
    B, H, W, Cn = dx.shape
    c = _ctx(dx)
    c.check(c.lib.mml_maxpool_k_bwd(c.handle, _p(dy, torch.bfloat16), _p(dy_flat, torch.float32), _p(argmax, torch.uint8), _p(dx, torch.bfloat16),
                                    B, H, W, Cn, int(k), _stream(dx)), "mml_maxpool_k_bwd")


def bn_conv_bias_fold(conv_bias, momentum: float, running_mean=None, scale=None, shift=None) -> None:
    c = _ctx(conv_bias)
    f = torch.float32
    c.check(c.lib.mml_bn_conv_bias_fold(c.handle, _p(conv_bias, f), conv_bias.numel(), float(momentum), _p(running_mean, f), _p(scale, f),
                                        _p(shift, f), _stream(conv_bias)), "mml_bn_conv_bias_fold")


def avgpool_fwd(x, y, N, HW, Cn) -> None:
    ctx = _ctx(x)
    ctx.check(ctx.lib.mml_avgpool_fwd(ctx.handle, _p(x, BF16), _p(y, torch.float32), N, HW, Cn, _stream(x)), "avgpool_fwd")


def avgpool_bwd(dy, dx, N, HW, Cn) -> None:
    ctx = _ctx(dy)
    ctx.check(ctx.lib.mml_avgpool_bwd(ctx.handle, _p(dy, torch.float32), _p(dx, BF16), N, HW, Cn, _stream(dy)), "avgpool_bwd")


# ---- head ----------------------------------------------------------------------------------------------------
def head_params(fcA_w, fcA_b, fcI_w, fcI_b, w0, b0, w3, b3, w5, b5) -> HeadParams:
    ts = (fcA_w, fcA_b, fcI_w, fcI_b, w0, b0, w3, b3, w5, b5)
    for t in ts:
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise MMLError("head parameters must be contiguous fp32")
    hp = HeadParams(*[t.data_ptr() for t in ts], fcA_w.shape[1], fcI_w.shape[1], fcA_w.shape[0], fcI_w.shape[0], w0.shape[0], w3.shape[0], w5.shape[0])
    if w0.shape[1] != hp.EA + hp.EI or w3.shape[1] != hp.H1 or w5.shape[1] != hp.H2:
        raise MMLError("head parameter shapes are inconsistent")
    return hp


def head_grads(*ts) -> HeadGrads:
    for t in ts:
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise MMLError("head gradients must be contiguous fp32")
    return HeadGrads(*[t.data_ptr() for t in ts])


def head_scratch_per_sample(hp: HeadParams) -> int:
    from ._lib import load_library

    return load_library().mml_head_scratch_per_sample(C.byref(hp))


def head_fwd(hp, pooledA, pooledI, labels, drop_mask, drop_scale, scratch, logits, loss_out, pred) -> None:
    ctx = _ctx(pooledA)
    B = pooledA.shape[0]
    ctx.check(ctx.lib.mml_head_fwd(ctx.handle, C.byref(hp), _p(pooledA, torch.float32), _p(pooledI, torch.float32), _p(labels), _p(drop_mask),
                                   float(drop_scale), _p(scratch), _p(logits), _p(loss_out), _p(pred, torch.int32), B, _stream(pooledA)), "head_fwd")


def head_bwd(hp, hg, pooledA, pooledI, labels, drop_mask, drop_scale, scratch, loss_scale, dpooledA, dpooledI, phases: int = 3) -> None:
    ctx = _ctx(pooledA)
    B = pooledA.shape[0]
    ctx.check(ctx.lib.mml_head_bwd(ctx.handle, C.byref(hp), C.byref(hg), _p(pooledA), _p(pooledI), _p(labels, torch.int64), _p(drop_mask),
                                   float(drop_scale), _p(scratch), float(loss_scale), _p(dpooledA), _p(dpooledI), B, int(phases), _stream(pooledA)),
              "head_bwd")


def linear_fwd(x, w, bias, y) -> None:
    ctx = _ctx(x)
    ctx.check(ctx.lib.mml_linear_fwd(ctx.handle, _p(x, torch.float32), _p(w, torch.float32), _p(bias), _p(y, torch.float32), x.shape[0], x.shape[1],
                                     w.shape[0], _stream(x)), "linear_fwd")


_M64 = (1 << 64) - 1


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def engine_seed(client_id: int = 0, salt: int = 0) -> int:
    """Seed of one engine's dropout stream: ``torch.initial_seed()`` (so torch.manual_seed / the run's seed matter) mixed with the
    data-parallel rank (each rank drops different units of its shard) and a client id (FedAvg clients draw different masks).  The
    kernel mixes in the optimizer step counter, ``salt`` separates forward()-only calls in train mode from each other."""
    import os

    rank = int(os.environ.get("RANK", "0"))
    x = _splitmix64(torch.initial_seed() & _M64)
    x = _splitmix64(x ^ (rank + 1) * 0xD6E8FEB86659FD93 & _M64)
    x = _splitmix64(x ^ (int(client_id) + 1) * 0xA0761D6478BD642F & _M64)
    if salt:
        x = _splitmix64(x ^ (int(salt) * 0xE7037ED1A0B428DB & _M64))
    return x


def dropout_mask(mask, p, seed, step_counter) -> None:
    ctx = _ctx(mask)
    ctx.check(ctx.lib.mml_dropout_mask(ctx.handle, _p(mask, torch.uint8), mask.numel(), float(p), int(seed), _p(step_counter), _stream(mask)),
              "dropout_mask")


# ---- optimizer / aggregation ----------------------------------------------------------------------------------------
def adam_step(p, g, m, v, p_bf16, hyper, step, advance_step: bool = True) -> None:
    ctx = _ctx(p)
    ctx.check(ctx.lib.mml_adam_step(ctx.handle, _p(p, torch.float32), _p(g, torch.float32), _p(m, torch.float32), _p(v, torch.float32), _p(p_bf16),
                                    p.numel(), _p(hyper, torch.float32), _p(step, torch.int64), int(advance_step), _stream(p)), "adam_step")


def cast_bf16_f32(src, dst) -> None:
    ctx = _ctx(src)
    ctx.check(ctx.lib.mml_cast_bf16_f32(ctx.handle, _p(src, BF16), _p(dst, torch.float32), src.numel(), _stream(src)), "cast_bf16_f32")


def cast_f32_bf16(src, dst) -> None:
    ctx = _ctx(src)
    ctx.check(ctx.lib.mml_cast_f32_bf16(ctx.handle, _p(src, torch.float32), _p(dst, BF16), src.numel(), _stream(src)), "cast_f32_bf16")


def fedavg(client_ptrs, weights, K, out) -> None:
    """client_ptrs: device int64 tensor [K] of device pointers to fp32 buffers of out.numel() elements."""
    ctx = _ctx(out)
    ctx.check(ctx.lib.mml_fedavg(ctx.handle, _p(client_ptrs, torch.int64), _p(weights, torch.float32), K, _p(out, torch.float32), out.numel(),
                                 _stream(out)), "fedavg")


def scale_inplace(x, weights, idx) -> None:
    ctx = _ctx(x)
    ctx.check(ctx.lib.mml_scale_inplace(ctx.handle, _p(x, torch.float32), _p(weights, torch.float32), idx, x.numel(), _stream(x)), "scale_inplace")


# ---- data-parallel gradient all-reduce (library-owned NCCL communicator) --------------------------------------------------
COMM_ID_BYTES = 128


def comm_unique_id(device_index: int) -> bytes:
    c = Context.get(device_index)
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    c.check(c.lib.mml_comm_unique_id(c.handle, buf), "mml_comm_unique_id")
    return bytes(buf)


def comm_init(device_index: int, comm_id: bytes, rank: int, world: int, max_ctas: int = 0) -> None:
    if len(comm_id) != COMM_ID_BYTES:
        raise MMLError("communicator id must be 128 bytes")
    c = Context.get(device_index)
    buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(comm_id)
    c.check(c.lib.mml_comm_init(c.handle, buf, int(rank), int(world), int(max_ctas)), "mml_comm_init")


def comm_world(device_index: int) -> int:
    c = Context.get(device_index)
    return int(c.lib.mml_comm_world(c.handle))


def allreduce_bucket(buf) -> None:
    """In-place fp32 sum all-reduce of ``buf`` over the context's communicator, on the current stream."""
    ctx = _ctx(buf)
    ctx.check(ctx.lib.mml_allreduce_bucket(ctx.handle, _p(buf, torch.float32), buf.numel(), _stream(buf)), "mml_allreduce_bucket")


def debug_set(key: int, value: int) -> None:
    from ._lib import load_library

    load_library().mml_debug_set(int(key), int(value))


def _ctx_sm_count(device_index: int) -> int:
    return Context.get(device_index).sm_count


def set_sm_budget(device_index: int, sms: int) -> None:
    """SMs the persistent conv kernels may occupy for the launches that follow (0 = all)."""
    c = Context.get(device_index)
    c.check(c.lib.mml_ctx_set_sm_budget(c.handle, int(sms)), "mml_ctx_set_sm_budget")


def set_pdl(device_index: int, enable: bool) -> None:
    """Programmatic dependent launch for the launches that follow (see mml_ctx_set_pdl)."""
    c = Context.get(device_index)
    c.check(c.lib.mml_ctx_set_pdl(c.handle, int(bool(enable))), "mml_ctx_set_pdl")


def launch_count(device_index: int = 0) -> int:
    return Context.get(device_index).launches


# ---- MMIMDb gated late fusion (config 3) ----------------------------------------------------------------------------
BN1D_INPUT, BN1D_GATED, BN1D_MAXOUT, BN1D_MAX2 = 0, 1, 2, 3


def bn1d_fwd_desc(mode: int, B: int, Cn: int, gamma, beta, rmean, rvar, *, x=None, mask=None, h1=None, h2=None, gate=None, pre=None,
                  keep=None, keep_scale: float = 1.0, xhat=None, invstd=None, y_bf16=None, y_f32=None, momentum=0.1, eps=1e-5,
                  mix_a: float = 1.0, mix_b: float = 1.0) -> BN1dDesc:
    """Descriptor for ``bn1d_fwd`` (built once per plan; ``train`` / ``keep`` are set per launch)."""
    d = BN1dDesc()
    d.mode, d.B, d.C, d.train = mode, B, Cn, 1
    d.x, d.mask, d.ldx = _p(x, torch.float32), _p(mask, torch.float32), (x.stride(0) if x is not None else 0)
    d.h1, d.h2, d.gate = _p(h1, torch.float32), _p(h2, torch.float32), _p(gate, torch.float32)
    d.pre, d.keep, d.keep_scale = _p(pre, torch.bfloat16), _p(keep, torch.uint8), keep_scale
    d.momentum, d.eps, d.mix_a, d.mix_b = momentum, eps, mix_a, mix_b
    d.gamma, d.beta, d.running_mean, d.running_var = (_p(t, torch.float32) for t in (gamma, beta, rmean, rvar))
    d.xhat, d.invstd = _p(xhat, torch.float32), _p(invstd, torch.float32)
    d.y_bf16, d.ldy, d.y_f32 = _p(y_bf16, torch.bfloat16), (y_bf16.stride(0) if y_bf16 is not None else 0), _p(y_f32, torch.float32)
    d._keep_alive = (x, mask, h1, h2, gate, pre, keep, gamma, beta, rmean, rvar, xhat, invstd, y_bf16, y_f32)
    d._device = gamma
    return d


def bn1d_fwd(d: BN1dDesc, train: bool, use_keep: bool = True) -> None:
    ref = d._device
    c = _ctx(ref)
    keep = d.keep
    d.train = 1 if train else 0
    if not (train and use_keep):
        d.keep = None
    try:
        c.check(c.lib.mml_bn1d_fwd(c.handle, C.byref(d), _stream(ref)), "mml_bn1d_fwd")
    finally:
        d.keep = keep


def bn1d_bwd_desc(mode: int, B: int, Cn: int, dy, xhat, gamma, invstd, dgamma, dbeta, *, pre=None, keep=None, keep_scale: float = 1.0,
                  dpre=None, dz=None) -> BN1dBwdDesc:
    d = BN1dBwdDesc()
    d.mode, d.B, d.C = mode, B, Cn
    d.dy, d.lddy = _p(dy, torch.bfloat16), dy.stride(0)
    d.xhat, d.gamma, d.invstd = _p(xhat, torch.float32), _p(gamma, torch.float32), _p(invstd, torch.float32)
    d.dgamma, d.dbeta = _p(dgamma, torch.float32), _p(dbeta, torch.float32)
    d.pre, d.keep, d.keep_scale, d.dpre = _p(pre, torch.bfloat16), _p(keep, torch.uint8), keep_scale, _p(dpre, torch.bfloat16)
    d.dz = _p(dz, torch.float32)
    d._keep_alive = (dy, xhat, gamma, invstd, dgamma, dbeta, pre, keep, dpre, dz)
    d._device = xhat
    return d


def bn1d_bwd(d: BN1dBwdDesc, use_keep: bool = True) -> None:
    ref = d._device
    c = _ctx(ref)
    keep = d.keep
    if not use_keep:
        d.keep = None
    try:
        c.check(c.lib.mml_bn1d_bwd(c.handle, C.byref(d), _stream(ref)), "mml_bn1d_bwd")
    finally:
        d.keep = keep


def gmu_fwd(h1pre, h2pre, wz, h1, h2, gate) -> None:
    B, H = h1.shape
    c = _ctx(h1)
    c.check(c.lib.mml_gmu_fwd(c.handle, _p(h1pre, torch.bfloat16), _p(h2pre, torch.bfloat16), _p(wz, torch.float32), _p(h1, torch.float32),
                              _p(h2, torch.float32), _p(gate, torch.float32), B, H, _stream(h1)), "mml_gmu_fwd")


def gmu_bwd(dz, h1, h2, gate, wz, dwz, dh1pre, dh2pre) -> None:
    B, H = h1.shape
    c = _ctx(h1)
    c.check(c.lib.mml_gmu_bwd(c.handle, _p(dz, torch.float32), _p(h1, torch.float32), _p(h2, torch.float32), _p(gate, torch.float32),
                              _p(wz, torch.float32), _p(dwz, torch.float32), _p(dh1pre, torch.bfloat16), _p(dh2pre, torch.bfloat16), B, H,
                              _stream(h1)), "mml_gmu_bwd")


def bce_head_scratch_floats(B: int) -> int:
    from ._lib import load_library
    return int(load_library().mml_bce_head_scratch_floats(int(B)))


def bce_head_fwd(xn, w, bias, labels, logits, loss, dlogits, pred, scratch, threshold: float, grad_scale: float = 1.0) -> None:
    B, H = xn.shape
    NC = w.shape[0]
    c = _ctx(xn)
    c.check(c.lib.mml_bce_head_fwd(c.handle, _p(xn, torch.float32), _p(w, torch.float32), _p(bias, torch.float32), _p(labels, torch.float32),
                                   _p(logits, torch.float32), _p(loss, torch.float32), _p(dlogits, torch.float32), _p(pred, torch.uint8),
                                   _p(scratch, torch.float32), float(threshold), float(grad_scale), B, H, NC, _stream(xn)), "mml_bce_head_fwd")


def bce_head_bwd(dlogits, xn, w, dw, db, dxn) -> None:
    B, H = xn.shape
    NC = w.shape[0]
    c = _ctx(xn)
    c.check(c.lib.mml_bce_head_bwd(c.handle, _p(dlogits, torch.float32), _p(xn, torch.float32), _p(w, torch.float32), _p(dw, torch.float32),
                                   _p(db, torch.float32), _p(dxn, torch.bfloat16), B, H, NC, _stream(xn)), "mml_bce_head_bwd")


def pool_fwd(pre_a, pre_b, bias_a, bias_b, keep_a, keep_b, keep_scale: float, h_a, h_b, comb=None) -> None:
    B, H = h_a.shape
    c = _ctx(h_a)
    c.check(c.lib.mml_pool_fwd(c.handle, _p(pre_a, torch.bfloat16), _p(pre_b, torch.bfloat16), _p(bias_a, torch.float32), _p(bias_b, torch.float32),
                               _p(keep_a, torch.uint8), _p(keep_b, torch.uint8), float(keep_scale), _p(h_a, torch.float32), _p(h_b, torch.float32),
                               _p(comb, torch.bfloat16), B, H, _stream(h_a)), "mml_pool_fwd")


def pool_bwd(dz, h_a, h_b, keep_a, keep_b, keep_scale: float, kind: int, mix_a: float, mix_b: float, dpre_a, dpre_b, dbias_a, dbias_b,
             gate=None, dcomb=None) -> None:
    B, H = h_a.shape
    c = _ctx(h_a)
    c.check(c.lib.mml_pool_bwd(c.handle, _p(dz, torch.float32), _p(h_a, torch.float32), _p(h_b, torch.float32), _p(keep_a, torch.uint8),
                               _p(keep_b, torch.uint8), float(keep_scale), int(kind), float(mix_a), float(mix_b), _p(gate, torch.float32),
                               _p(dcomb, torch.bfloat16), _p(dpre_a, torch.bfloat16),
                               _p(dpre_b, torch.bfloat16), _p(dbias_a, torch.float32), _p(dbias_b, torch.float32), B, H, _stream(h_a)), "mml_pool_bwd")


# ---- MonomodalEncoder tail (encoder fc -> classifier -> CE) ----------------------------------------------------------
def mono_head_fwd(pooled, fc_w, fc_b, cls_w, cls_b, labels, emb, logits, dlogits, row_loss, loss, pred, loss_scale: float = 1.0) -> None:
    B, Fd = pooled.shape
    E, NC = fc_w.shape[0], cls_w.shape[0]
    c = _ctx(pooled)
    f32 = torch.float32
    c.check(c.lib.mml_mono_head_fwd(c.handle, _p(pooled, f32), _p(fc_w, f32), _p(fc_b, f32), _p(cls_w, f32), _p(cls_b, f32), _p(labels, torch.int64),
                                    _p(emb, f32), _p(logits, f32), _p(dlogits, f32), _p(row_loss, f32), _p(loss, f32), _p(pred, torch.int32),
                                    float(loss_scale), B, Fd, E, NC, _stream(pooled)), "mml_mono_head_fwd")


def mono_head_bwd(pooled, emb, dlogits, fc_w, cls_w, d_fc_w, d_fc_b, d_cls_w, d_cls_b, demb, dpooled) -> None:
    B, Fd = pooled.shape
    E, NC = fc_w.shape[0], cls_w.shape[0]
    c = _ctx(pooled)
    f32 = torch.float32
    c.check(c.lib.mml_mono_head_bwd(c.handle, _p(pooled, f32), _p(emb, f32), _p(dlogits, f32), _p(fc_w, f32), _p(cls_w, f32), _p(d_fc_w, f32),
                                    _p(d_fc_b, f32), _p(d_cls_w, f32), _p(d_cls_b, f32), _p(demb, f32), _p(dpooled, f32), B, Fd, E, NC,
                                    _stream(pooled)), "mml_mono_head_bwd")


def att_fwd(hid, b0, w2, b2, t, gate) -> None:
    B, Hd = t.shape
    c = _ctx(t)
    c.check(c.lib.mml_att_fwd(c.handle, _p(hid, torch.bfloat16), _p(b0, torch.float32), _p(w2, torch.float32), _p(b2, torch.float32),
                              _p(t, torch.float32), _p(gate, torch.float32), B, Hd, w2.shape[0], _stream(t)), "mml_att_fwd")


def att_bwd(dz, h_a, h_b, gate, t, w2, dw2, db2, db0, dhid) -> None:
    B, H = h_a.shape
    Hd = t.shape[1]
    c = _ctx(t)
    c.check(c.lib.mml_att_bwd(c.handle, _p(dz, torch.float32), _p(h_a, torch.float32), _p(h_b, torch.float32), _p(gate, torch.float32),
                              _p(t, torch.float32), _p(w2, torch.float32), _p(dw2, torch.float32), _p(db2, torch.float32), _p(db0, torch.float32),
                              _p(dhid, torch.bfloat16), B, H, Hd, w2.shape[0], _stream(t)), "mml_att_bwd")


# ---- MOSI / UttFusion (config 4) -----------------------------------------------------------------------------------------
def lstm_fwd(x, w_ih, w_hh, b_ih, b_hh, gates, cs, hs, h_last) -> None:
    B, T, IN = x.shape
    c = _ctx(x)
    f = torch.float32
    c.check(c.lib.mml_lstm_fwd(c.handle, _p(x, f), _p(w_ih, f), _p(w_hh, f), _p(b_ih, f), _p(b_hh, f), _p(gates, f), _p(cs, f), _p(hs, f),
                               _p(h_last, f), B, T, IN, w_hh.shape[1], _stream(x)), "mml_lstm_fwd")


def lstm_bwd(x, w_hh, gates, cs, hs, dh_last, dw_ih, dw_hh, db_ih, db_hh) -> None:
    B, T, IN = x.shape
    c = _ctx(x)
    f = torch.float32
    c.check(c.lib.mml_lstm_bwd(c.handle, _p(x, f), _p(w_hh, f), _p(gates, f), _p(cs, f), _p(hs, f), _p(dh_last, f), _p(dw_ih, f), _p(dw_hh, f),
                               _p(db_ih, f), _p(db_hh, f), B, T, IN, w_hh.shape[1], _stream(x)), "mml_lstm_bwd")


def relumax_fwd(conv, bias, keep, keep_scale: float, y, arg, y_off: int) -> None:
    B, P, Cn = conv.shape
    c = _ctx(y)
    c.check(c.lib.mml_relumax_fwd(c.handle, _p(conv, torch.bfloat16), _p(bias, torch.float32), _p(keep, torch.uint8), float(keep_scale),
                                  _p(y, torch.float32), _p(arg, torch.int32), B, P, Cn, y.shape[1], int(y_off), _stream(y)), "mml_relumax_fwd")


def relumax_bwd(dy, arg, keep, keep_scale: float, dconv, dbias, y_off: int) -> None:
    B, P, Cn = dconv.shape
    c = _ctx(dy)
    c.check(c.lib.mml_relumax_bwd(c.handle, _p(dy, torch.float32), _p(arg, torch.int32), _p(keep, torch.uint8), float(keep_scale),
                                  _p(dconv, torch.bfloat16), _p(dbias, torch.float32), B, P, Cn, dy.shape[1], int(y_off), _stream(dy)), "mml_relumax_bwd")


def dense_fwd(x, ldx: int, w, bias, keep, keep_scale: float, relu: bool, y, ldy: int, B: int) -> None:
    N, K = w.shape
    c = _ctx(w)
    f = torch.float32
    c.check(c.lib.mml_dense_fwd(c.handle, C.c_void_p(x.data_ptr()), int(ldx), _p(w, f), _p(bias, f), _p(keep, torch.uint8), float(keep_scale),
                                int(relu), C.c_void_p(y.data_ptr()), int(ldy), B, K, N, _stream(w)), "mml_dense_fwd")


def dense_bwd(dy, y, ldy: int, keep, keep_scale: float, relu: bool, x, ldx: int, w, dx, lddx: int, dw, db, B: int, lddy: int = 0) -> None:
    N, K = w.shape
    c = _ctx(w)
    f = torch.float32
    c.check(c.lib.mml_dense_bwd(c.handle, C.c_void_p(dy.data_ptr()), int(lddy) or N, C.c_void_p(y.data_ptr()), int(ldy), _p(keep, torch.uint8), float(keep_scale), int(relu),
                                C.c_void_p(x.data_ptr()), int(ldx), _p(w, f), C.c_void_p(dx.data_ptr()) if dx is not None else C.c_void_p(0), int(lddx),
                                _p(dw, f), _p(db, f), B, K, N, _stream(w)), "mml_dense_bwd")


def clip_grad_scale(g, clip: float, base_scale: float, hyper, groups: int, partial, norm_out=None) -> None:
    c = _ctx(g)
    c.check(c.lib.mml_clip_grad_scale(c.handle, _p(g, torch.float32), g.numel(), float(clip), float(base_scale), _p(hyper, torch.float32), int(groups),
                                      _p(partial, torch.float64), _p(norm_out, torch.float32), _stream(g)), "mml_clip_grad_scale")


def softmax_ce(logits, labels, dlogits, row_loss, loss, pred, loss_scale: float = 1.0) -> None:
    B, NC = logits.shape
    c = _ctx(logits)
    f = torch.float32
    c.check(c.lib.mml_softmax_ce(c.handle, _p(logits, f), _p(labels, torch.int64), _p(dlogits, f), _p(row_loss, f), _p(loss, f), _p(pred, torch.int32),
                                 float(loss_scale), B, NC, _stream(logits)), "mml_softmax_ce")
