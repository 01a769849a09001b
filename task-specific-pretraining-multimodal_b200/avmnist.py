"""AVMNIST late-fusion model -- drop-in for ``MML_Suite/models/avmnist.py:188-410`` on a B200.

Same constructor (``AVMNIST(audio_encoder, image_encoder, hidden_dim, *, dropout=0.0, fusion_fn="concat")``), same
sub-module names (``audio_encoder``, ``image_encoder``, ``net.0/3/5``) hence the same 346-entry ``state_dict()``, same
``forward(A=None, I=None, *, is_embd_A=False, is_embd_I=False)`` and the ``MultimodalModelProtocol`` methods
(``train_step``, ``validation_step``, ``get_embeddings``, ``get_encoder``, ``flatten_parameters``; protocols.py:13-40).

What differs is where the arithmetic happens: ``train_step`` runs ONE fused schedule (mask -> both encoders -> concat
head -> CE -> backward -> [allreduce] -> Adam) in libmml_b200.so, replayed as a CUDA graph, instead of autograd over
~400 ATen/cuDNN launches.  The missing-modality mask may be applied on the device: besides the reference's batch
contract (already masked tensors under ``Modality.AUDIO`` / ``Modality.IMAGE``) a batch may carry
``"<mod>_original"`` + ``"<mod>_missing_index"`` and the multiply of data/base_dataset.py:71 is then done by the stem
kernel -- bit-identical result.  Unsupported requests (other optimizers, fusion functions, embeddings as inputs)
raise; nothing silently falls back to PyTorch.
"""
from __future__ import annotations

import weakref
from typing import Any, Dict, Optional

import numpy as np
import torch
import torch.nn as nn

NUM_CLASSES = 10  # AVMNISTDataset.NUM_CLASSES (data/avmnist.py)


def _find(batch: Dict[Any, Any], name: str):
    """Batch keys are ``modalities.Modality`` members in the reference (str() == lower-case name) or plain strings."""
    if name in batch:
        return batch[name]
    for k, v in batch.items():
        if not isinstance(k, str) and str(k).lower().split(".")[-1] == name:
            return v
    return None


def _copy_in(dst: torch.Tensor, src: torch.Tensor) -> None:
    """Stage ``src`` into the plan's static input buffer.  Host sources are an H2D memcpy.  Device sources (a prefetched batch)
    are copied by a KERNEL: a device-to-device ``cudaMemcpyAsync`` is served by a copy engine and queues behind the
    prefetcher's in-flight H2D transfer of the next batch, which put the whole copy back on the critical path (measured:
    4.3 ms/step instead of 3.45)."""
    if src.is_cuda and src.dtype == dst.dtype and src.shape == dst.shape:
        torch.mul(src, 1, out=dst)
    else:
        dst.copy_(src, non_blocking=True)


class AVMNIST(nn.Module):
    def __init__(self, audio_encoder: nn.Module, image_encoder: nn.Module, hidden_dim: int, *, dropout: float = 0.0,
                 fusion_fn: str = "concat") -> None:
        super().__init__()
        if fusion_fn.lower() != "concat":
            raise ValueError(f"Unknown fusion function: {fusion_fn}")
        self.audio_encoder = audio_encoder
        self.image_encoder = image_encoder
        self.embd_size_A = audio_encoder.get_embedding_size()
        self.embd_size_I = image_encoder.get_embedding_size()
        fc_fusion = nn.Linear(self.embd_size_A + self.embd_size_I, hidden_dim)
        fc_intermediate = nn.Linear(hidden_dim, hidden_dim // 2)
        fc_out = nn.Linear(hidden_dim // 2, NUM_CLASSES)
        # indices 0 / 3 / 5 carry parameters, exactly like the reference's Sequential
        self.net = nn.Sequential(fc_fusion, nn.ReLU(), nn.Dropout(dropout) if dropout > 0 else nn.Identity(), fc_intermediate, nn.ReLU(), fc_out)
        self.dropout_p = float(dropout)
        self._engine = None
        self.world_size = 1
        self._dp = None

    # ---- mode switching -----------------------------------------------------------------------------------------
    def train(self, mode: bool = True):
        super().train(mode)
        self._uniform_mode = bool(mode)  # every sub-module now agrees with ``self.training``
        return self

    def _set_mode(self, training: bool) -> None:
        """``self.train()`` / ``self.eval()`` of the reference's step methods without walking ~190 sub-modules on every
        step (0.3 ms of host time per call, i.e. ~8 % of a B200 step) when the mode is already set."""
        if getattr(self, "_uniform_mode", None) is not training or self.training is not training:
            self.train(training)

    # ---- engine plumbing ---------------------------------------------------------------------------------------
    def _get_engine(self, device: torch.device):
        from .engine import LateFusionEngine

        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("mml_b200.AVMNIST runs on a B200 GPU only: there is no CPU / PyTorch fallback path")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        eng = self._engine
        if eng is None or eng.device != device:
            from .convblock import _ConvBlockEncoder, make_engine

            kinds = {isinstance(e, _ConvBlockEncoder) for e in (self.audio_encoder, self.image_encoder)}
            if kinds == {True}:  # AVMNIST(MNISTAudio, MNISTImage): configs/avmnist/centralised/train_avmnist.yaml
                eng = self._engine = make_engine(self, device, self.dropout_p)
            elif kinds == {False}:
                eng = self._engine = LateFusionEngine(self, device, self.dropout_p)
                self.audio_encoder._mml_owner = (weakref.ref(eng), "audio_encoder.")
                self.image_encoder._mml_owner = (weakref.ref(eng), "image_encoder.")
            else:
                raise NotImplementedError("mixing a ResNet encoder with a ConvBlock encoder is not a configuration of the reference")
            if self._dp is not None:
                self._dp.attach(eng)
        eng.fs.ensure_fresh()
        return eng

    def flatten_parameters(self) -> None:
        """MultimodalModelProtocol.flatten_parameters: (re)home all parameters in the flat device buffers."""
        p = next(self.parameters())
        if p.is_cuda:
            self._get_engine(p.device)

    def enable_data_parallel(self, dp) -> None:
        """dp: mml_b200.dist.DataParallel (one process per GPU, NCCL)."""
        self._dp = dp
        self.world_size = dp.world_size
        if self._engine is not None:
            dp.attach(self._engine)

    def get_encoder(self, modality) -> nn.Module:
        name = str(modality).lower().split(".")[-1]
        if name == "audio":
            return self.audio_encoder
        if name == "image":
            return self.image_encoder
        raise ValueError(f"Unknown modality: {modality}")

    # ---- staging ----------------------------------------------------------------------------------------------------
    @staticmethod
    def _as_bhw(t: torch.Tensor) -> torch.Tensor:
        if t.dim() == 4:
            if t.shape[1] != 1:
                raise ValueError("expected a 1-channel tensor")
            t = t[:, 0]
        if t.dim() != 3:
            raise ValueError(f"expected [B,H,W] or [B,1,H,W], got {tuple(t.shape)}")
        return t

    def _stage(self, eng, A, I, mask_a=None, mask_i=None, labels=None):
        for name, t in (("audio", A), ("image", I)):
            if t.dtype == torch.uint8:  # a plain cast would feed 0..255 instead of the reference's colormap -> "L" -> [0, 1] chain
                raise TypeError(f"{name}: uint8 pixels must go through the luminance table first (DevicePrefetcher(luts=...) / "
                                "datasets.AVMNIST.fused_loader, or datasets.AVMNIST.batches(image_form='f32'))")
        A, I = self._as_bhw(A), self._as_bhw(I)
        B = A.shape[0]
        if I.shape[0] != B:
            raise ValueError("audio and image batch sizes differ")
        plan = eng.plan_for(B, A.shape[1], A.shape[2], I.shape[1], I.shape[2])
        _copy_in(plan.audio.x, A)
        _copy_in(plan.image.x, I)
        for enc_plan, m in ((plan.audio, mask_a), (plan.image, mask_i)):
            if m is None:
                enc_plan.mask.fill_(1.0)
            else:
                _copy_in(enc_plan.mask, torch.as_tensor(m).reshape(B))
        if labels is not None:
            from . import ops

            ops.check_class_labels(labels, NUM_CLASSES)
            _copy_in(plan.labels, torch.as_tensor(labels).reshape(B))
        from .data import note_inputs_consumed
        note_inputs_consumed(eng.device)  # a prefetcher may overwrite the batch's device buffers from here on
        return plan

    # ---- forward ---------------------------------------------------------------------------------------------------
    def forward(self, A: Optional[torch.Tensor] = None, I: Optional[torch.Tensor] = None, *, is_embd_A: bool = False,
                is_embd_I: bool = False) -> torch.Tensor:
        """Logits [B,10] fp32.  train(): batch-statistics BN (running stats updated) + dropout; eval(): running stats.

        No autograd graph is recorded: training goes through ``train_step`` (the fused path).
        """
        assert not all((A is None, I is None)), "At least one of A, I must be provided"
        assert not all([is_embd_A, is_embd_I]), "Cannot have all embeddings as True"
        if is_embd_A or is_embd_I or A is None or I is None:
            raise NotImplementedError(
                "mml_b200.AVMNIST.forward needs both raw modalities; pre-computed embeddings / a None modality are outside the "
                "late-fusion hot path (the reference's own None branch builds a CPU tensor of embedding width, avmnist.py:261-262)")
        eng = self._get_engine(A.device if A.is_cuda else next(self.parameters()).device)
        plan = self._stage(eng, A, I)
        if self.training:
            plan.run_forward_train_mode()
        else:
            plan.run_eval(with_loss=False)
        return plan.logits.clone()

    # ---- MultimodalModelProtocol -------------------------------------------------------------------------------------
    def _unpack(self, batch: Dict[Any, Any]):
        """(audio, image, audio_mask, image_mask, labels, pattern_name) from a collated batch.

        Reference contract (data/avmnist.py:248-277): already-masked tensors under Modality.AUDIO / Modality.IMAGE.
        Extension: ``<mod>_original`` + ``<mod>_missing_index`` => the x * mask of base_dataset.py:71 runs on the device.
        """
        A, I = _find(batch, "audio"), _find(batch, "image")
        mask_a = mask_i = None
        if "audio_original" in batch and "audio_missing_index" in batch:
            A, mask_a = batch["audio_original"], batch["audio_missing_index"]
        if "image_original" in batch and "image_missing_index" in batch:
            I, mask_i = batch["image_original"], batch["image_missing_index"]
        if A is None or I is None:
            raise KeyError("batch needs audio and image tensors (Modality.AUDIO / Modality.IMAGE or *_original + *_missing_index)")
        return A, I, mask_a, mask_i, batch["labels"], batch.get("pattern_name")

    def train_step(self, batch: Dict[Any, Any], optimizer: torch.optim.Optimizer, loss_functions, device, metric_recorder=None,
                   **kwargs) -> Dict[str, Any]:
        """One fused training step; returns {"loss": float} like avmnist.py:269-310."""
        eng = self._get_engine(device)
        self._check_loss(loss_functions)
        A, I, mask_a, mask_i, labels, miss_type = self._unpack(batch)
        self._set_mode(True)
        fs = eng.fs
        fs.adopt_optimizer(optimizer)
        fs.sync_hyper(optimizer, 1.0 / self.world_size)
        plan = self._stage(eng, A, I, mask_a, mask_i, labels)
        given = kwargs.get("dropout_mask")
        if given is not None:
            plan.drop_mask.copy_(torch.as_tensor(given).reshape(plan.drop_mask.shape).to(torch.uint8), non_blocking=True)
        plan.want_pred = metric_recorder is not None
        plan.train_step(given_dropout=given is not None)
        fs._host_step += 1
        plan.loss_ready.synchronize()  # loss / predictions are final after the forward half; the backward half keeps running (engine.publish)
        loss = float(plan.h_loss[0])
        if metric_recorder is not None:
            predictions = plan.h_pred.numpy().astype(np.int64)
            targets = labels.detach().cpu().numpy() if torch.is_tensor(labels) else np.asarray(labels)
            metric_recorder.update_group_all("classification", predictions=predictions, targets=targets, m_types=np.array(miss_type))
        return {"loss": loss}

    def validation_step(self, batch: Dict[Any, Any], loss_functions, device, metric_recorder=None, return_test_info: bool = False,
                        **kwargs) -> Dict[str, Any]:
        """Eval-mode forward + CE + argmax (avmnist.py:312-360)."""
        eng = self._get_engine(device)
        self._check_loss(loss_functions)
        A, I, mask_a, mask_i, labels, miss_type = self._unpack(batch)
        self._set_mode(False)
        plan = self._stage(eng, A, I, mask_a, mask_i, labels)
        plan.run_eval(with_loss=True)
        plan.h_loss.copy_(plan.loss, non_blocking=True)
        plan.h_pred.copy_(plan.pred, non_blocking=True)
        torch.cuda.current_stream(eng.device).synchronize()
        loss = float(plan.h_loss[0])
        predictions = plan.h_pred.numpy().astype(np.int64)
        targets = labels.detach().cpu().numpy() if torch.is_tensor(labels) else np.asarray(labels)
        mt = np.array(miss_type)
        if metric_recorder is not None:
            metric_recorder.update_group_all(group_name="classification", predictions=predictions, targets=targets, m_types=mt)
        if return_test_info:
            return {"loss": loss, "predictions": predictions, "labels": targets, "miss_types": mt}
        return {"loss": loss}

    def get_embeddings(self, dataloader, device) -> Dict[Any, Any]:
        """Per-modality encoder embeddings of the fully-available samples (avmnist.py:362-401)."""
        from collections import defaultdict

        out = defaultdict(list)
        self._set_mode(False)
        self._get_engine(device)
        for batch in dataloader:
            A, I = _find(batch, "audio"), _find(batch, "image")
            miss = np.array(batch["pattern_name"])
            keep = torch.from_numpy(miss == "ai")
            A, I = A[keep].to(device).float(), I[keep].to(device).float()
            if A.shape[0] == 0:
                continue
            ka = next((k for k in batch if str(k).lower().endswith("audio")), "audio")
            ki = next((k for k in batch if str(k).lower().endswith("image")), "image")
            out[ka].append(self.audio_encoder(A).cpu().numpy())
            out[ki].append(self.image_encoder(I).cpu().numpy())
            out["label"] += batch["labels"]
        return out

    @staticmethod
    def _check_loss(loss_functions) -> None:
        """The fused head implements what the reference's YAML resolves to: one CrossEntropyLoss() term, weight 1.0
        (experiment_utils/loss.py:48,98-113; ``loss_args`` is never read, :91).  Anything else must fail loudly.  ``ignore_index`` is not
        implemented by the kernels: labels are range-checked instead (ops.check_class_labels), so no row can match it."""
        if loss_functions is None:
            return
        items = list(loss_functions.items()) if hasattr(loss_functions, "items") else None
        if not items or len(items) != 1:
            raise NotImplementedError("mml_b200 fused step supports a LossFunctionGroup with exactly one cross_entropy term")
        term = items[0][1]
        fn, weight = getattr(term, "loss_fn", term), float(getattr(term, "weight", 1.0))
        ok = isinstance(fn, nn.CrossEntropyLoss) and fn.reduction == "mean" and fn.label_smoothing == 0.0 and fn.weight is None \
            and weight == 1.0
        if not ok:
            raise NotImplementedError("mml_b200 fused step implements CrossEntropyLoss() with default arguments and weight 1.0 only")
