"""ctypes binding of libmml_b200.so (the C ABI declared in include/mml_b200.h).

There is NO fallback: if the shared library is missing or the device is not an sm_100 part, importing the ops raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmml_b200.so")


class MMLError(RuntimeError):
    pass


class ConvGeom(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("N", "H", "W", "C", "K", "R", "S", "stride", "pad")]


class HeadParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("fcA_w", "fcA_b", "fcI_w", "fcI_b", "w0", "b0", "w3", "b3", "w5", "b5")] + [
        (n, C.c_int32) for n in ("FA", "FI", "EA", "EI", "H1", "H2", "NC")
    ]


class HeadGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("fcA_w", "fcA_b", "fcI_w", "fcI_b", "w0", "b0", "w3", "b3", "w5", "b5")]


class BN1dDesc(C.Structure):
    _fields_ = [("mode", C.c_int32), ("B", C.c_int32), ("C", C.c_int32), ("train", C.c_int32),
                ("x", C.c_void_p), ("mask", C.c_void_p), ("ldx", C.c_int64),
                ("h1", C.c_void_p), ("h2", C.c_void_p), ("gate", C.c_void_p),
                ("pre", C.c_void_p), ("keep", C.c_void_p), ("keep_scale", C.c_float), ("momentum", C.c_float), ("eps", C.c_float),
                ("mix_a", C.c_float), ("mix_b", C.c_float), ("reserved", C.c_float),
                ("gamma", C.c_void_p), ("beta", C.c_void_p), ("running_mean", C.c_void_p), ("running_var", C.c_void_p),
                ("xhat", C.c_void_p), ("invstd", C.c_void_p), ("y_bf16", C.c_void_p), ("ldy", C.c_int64), ("y_f32", C.c_void_p)]


class BN1dBwdDesc(C.Structure):
    _fields_ = [("mode", C.c_int32), ("B", C.c_int32), ("C", C.c_int32), ("reserved", C.c_int32),
                ("dy", C.c_void_p), ("lddy", C.c_int64), ("xhat", C.c_void_p), ("gamma", C.c_void_p), ("invstd", C.c_void_p),
                ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
                ("pre", C.c_void_p), ("keep", C.c_void_p), ("keep_scale", C.c_float), ("reserved2", C.c_float), ("dpre", C.c_void_p),
                ("dz", C.c_void_p)]


P, I32, I64, F32, U64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64

# name -> (restype, argtypes); every symbol declared in include/mml_b200.h
SIGNATURES = {
    "mml_version": (I32, []),
    "mml_bn_stat_slots": (I32, [I32]),
    "mml_ctx_create": (I32, [I32, C.POINTER(P)]),
    "mml_ctx_destroy": (None, [P]),
    "mml_last_error": (C.c_char_p, [P]),
    "mml_ctx_sm_count": (I32, [P]),
    "mml_ctx_launch_count": (I64, [P]),
    "mml_ctx_set_sm_budget": (I32, [P, I32]),
    "mml_ctx_set_pdl": (I32, [P, I32]),
    "mml_debug_set": (I32, [I32, I32]),
    "mml_mask_apply_f32": (I32, [P, P, P, P, P, I64, I64, P]),
    "mml_missing_mask_draw": (I32, [P, P, P, I32, I64, I64, I64, U64, C.c_uint32, P]),
    "mml_missing_mask_gather": (I32, [P, P, P, P, I32, I64, I64, P, P]),
    "mml_stage_u8_lut_f32": (I32, [P, P, P, P, I64, P]),
    "mml_stem_fprop": (I32, [P, P, P, P, P, P, I32, I32, I32, P]),
    "mml_stem_wgrad": (I32, [P, P, P, P, P, P, I64, I32, I32, I32, P]),
    "mml_stem_wgrad_workspace": (I64, [P, I32, I32, I32]),
    "mml_stem_wgrad_bn": (I32, [P] * 13 + [I64, I32, I32, I32, P]),
    "mml_conv_fprop": (I32, [P, C.POINTER(ConvGeom), P, P, P, P, P]),
    "mml_conv_dgrad": (I32, [P, C.POINTER(ConvGeom), P, P, P, P]),
    "mml_conv_wgrad": (I32, [P, C.POINTER(ConvGeom), P, P, P, P, I64, P]),
    "mml_conv_wgrad_workspace": (I64, [P, C.POINTER(ConvGeom)]),
    "mml_bn_train_fwd": (I32, [P] + [P] * 17 + [I64, I32, I32, F32, F32, P]),
    "mml_bn_eval_coeffs": (I32, [P, I32, P, P, P, P, F32, P, P, P]),
    "mml_bn_act_fwd": (I32, [P, P, P, P, P, P, P, P, I64, I32, I32, P]),
    "mml_bn_bwd_reduce": (I32, [P] * 9 + [I64, I32, I32, P]),
    "mml_bn_bwd_apply": (I32, [P] * 10 + [I64, I32, P]),
    "mml_maxpool3x3s2_fwd": (I32, [P, P, P, P, I32, I32, I32, I32, P]),
    "mml_maxpool3x3s2_bwd": (I32, [P, P, P, P, P, I32, I32, I32, I32, P]),
    "mml_stem_bn_pool_fwd": (I32, [P] * 13 + [I32, I32, I32, I32, F32, F32, P]),
    "mml_stem_bn_pool_bwd": (I32, [P] * 13 + [I32, I32, I32, I32, I32, P]),
    "mml_conv3x3_c1_fprop": (I32, [P, P, P, P, P, P, I32, I32, I32, I32, P]),
    "mml_conv3x3_c1_wgrad": (I32, [P, P, P, P, P, P, I64, I32, I32, I32, I32, P]),
    "mml_conv3x3_c1_wgrad_workspace": (I64, [P, I32, I32, I32]),
    "mml_maxpool_k_fwd": (I32, [P, P, P, P, P, I32, I32, I32, I32, I32, P]),
    "mml_maxpool_k_bwd": (I32, [P, P, P, P, P, I32, I32, I32, I32, I32, P]),
    "mml_bn_conv_bias_fold": (I32, [P, P, I32, F32, P, P, P, P]),
    "mml_avgpool_fwd": (I32, [P, P, P, I32, I32, I32, P]),
    "mml_avgpool_bwd": (I32, [P, P, P, I32, I32, I32, P]),
    "mml_head_scratch_per_sample": (I32, [C.POINTER(HeadParams)]),
    "mml_head_fwd": (I32, [P, C.POINTER(HeadParams), P, P, P, P, F32, P, P, P, P, I32, P]),
    "mml_head_bwd": (I32, [P, C.POINTER(HeadParams), C.POINTER(HeadGrads), P, P, P, P, F32, P, F32, P, P, I32, I32, P]),
    "mml_softmax_ce": (I32, [P, P, P, P, P, P, P, F32, I32, I32, P]),
    "mml_mono_head_fwd": (I32, [P] * 13 + [F32, I32, I32, I32, I32, P]),
    "mml_mono_head_bwd": (I32, [P] * 12 + [I32, I32, I32, I32, P]),
    "mml_linear_fwd": (I32, [P, P, P, P, P, I32, I32, I32, P]),
    "mml_dropout_mask": (I32, [P, P, I64, F32, U64, P, P]),
    "mml_bn1d_fwd": (I32, [P, C.POINTER(BN1dDesc), P]),
    "mml_bn1d_bwd": (I32, [P, C.POINTER(BN1dBwdDesc), P]),
    "mml_gmu_fwd": (I32, [P, P, P, P, P, P, P, I32, I32, P]),
    "mml_gmu_bwd": (I32, [P, P, P, P, P, P, P, P, P, I32, I32, P]),
    "mml_pool_fwd": (I32, [P, P, P, P, P, P, P, F32, P, P, P, I32, I32, P]),
    "mml_pool_bwd": (I32, [P, P, P, P, P, P, F32, I32, F32, F32, P, P, P, P, P, P, I32, I32, P]),
    "mml_att_fwd": (I32, [P, P, P, P, P, P, P, I32, I32, I32, P]),
    "mml_att_bwd": (I32, [P, P, P, P, P, P, P, P, P, P, P, I32, I32, I32, I32, P]),
    "mml_bce_head_scratch_floats": (I64, [I32]),
    "mml_bce_head_fwd": (I32, [P, P, P, P, P, P, P, P, P, P, F32, F32, I32, I32, I32, P]),
    "mml_bce_head_bwd": (I32, [P, P, P, P, P, P, P, I32, I32, I32, P]),
    "mml_lstm_fwd": (I32, [P] * 10 + [I32, I32, I32, I32, P]),
    "mml_lstm_bwd": (I32, [P] * 11 + [I32, I32, I32, I32, P]),
    "mml_relumax_fwd": (I32, [P, P, P, P, F32, P, P, I32, I32, I32, I32, I32, P]),
    "mml_relumax_bwd": (I32, [P, P, P, P, F32, P, P, I32, I32, I32, I32, I32, P]),
    "mml_dense_fwd": (I32, [P, P, I32, P, P, P, F32, I32, P, I32, I32, I32, I32, P]),
    "mml_dense_bwd": (I32, [P, P, I32, P, I32, P, F32, I32, P, I32, P, P, I32, P, P, I32, I32, I32, P]),
    "mml_clip_grad_scale": (I32, [P, P, I64, F32, F32, P, I32, P, P, P]),
    "mml_adam_step": (I32, [P, P, P, P, P, P, I64, P, P, I32, P]),
    "mml_cast_f32_bf16": (I32, [P, P, P, I64, P]),
    "mml_cast_bf16_f32": (I32, [P, P, P, I64, P]),
    "mml_comm_unique_id": (I32, [P, P]),
    "mml_comm_init": (I32, [P, P, I32, I32, I32]),
    "mml_comm_world": (I32, [P]),
    "mml_allreduce_bucket": (I32, [P, P, I64, P]),
    "mml_comm_destroy": (I32, [P]),
    "mml_fedavg": (I32, [P, P, P, I32, P, I64, P]),
    "mml_scale_inplace": (I32, [P, P, P, I32, I64, P]),
}

_lib: Optional[C.CDLL] = None
_lock = threading.Lock()


def load_library() -> C.CDLL:
    """dlopen libmml_b200.so and bind every symbol; raises MMLError if the library was not built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise MMLError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C task-specific-pretraining-multimodal_b200/csrc). There is no CPU / PyTorch fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError => header and library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


class Context:
    """One mml_ctx per (process, device)."""

    _by_device = {}

    def __init__(self, device_index: int):
        self.lib = load_library()
        h = P()
        rc = self.lib.mml_ctx_create(int(device_index), C.byref(h))
        if rc != 0:
            raise MMLError(f"mml_ctx_create({device_index}) failed ({rc}): {self.lib.mml_last_error(None).decode()}")
        self.handle = h
        self.device_index = int(device_index)
        self.sm_count = self.lib.mml_ctx_sm_count(h)

    @classmethod
    def get(cls, device_index: int) -> "Context":
        ctx = cls._by_device.get(device_index)
        if ctx is None:
            ctx = cls._by_device[device_index] = cls(device_index)
        return ctx

    def check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise MMLError(f"{what} failed ({rc}): {self.lib.mml_last_error(self.handle).decode()}")

    @property
    def launches(self) -> int:
        return int(self.lib.mml_ctx_launch_count(self.handle))
