"""MMIMDb gated late-fusion model -- drop-in for ``MML_Suite/models/mmimdb.py`` (+ ``models/gates/gated_bimodal.py``,
``models/maxout.py``) on a B200.

Same classes and constructors as the YAML tags build them (configs/mmimdb/centralised/mmimdb_baseline.yaml:10-31):
``MMIMDbModalityEncoder(input_dim, output_dim)``, ``GatedBiModalNetwork(input_one_dim, input_two_dim, output_one_dim,
output_two_dim, use_bias=False)``, ``MaxOut``, ``MLPGenreClassifier(input_size, output_size, hidden_size)`` and
``MMIMDb(image_encoder, text_encoder, gated_bimodal_network=..., classifier=..., binary_threshold=0.5)``; same sub-module
names, hence the same 38-entry ``state_dict()`` (``multimodal_pooling={...}`` builds ``MultimodalPooling`` -- max / avg / sum /
attention / gated -- instead of the GMU, mmimdb_pooling.yaml); same ``forward(I, T)`` / ``train_step`` / ``validation_step`` /
``get_embeddings`` / ``get_encoder`` / ``logits_transform``.  The torch modules inside are parameter CONTAINERS only (they
give the reference's initialisation and names); the arithmetic of a step is one fused schedule in libmml_b200.so
(``gated_engine.py``).  Unsupported requests (pooling fusion, biased GMU, embeddings as inputs, other optimizers or
losses) raise -- nothing falls back to PyTorch.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Any, Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from .avmnist import _copy_in, _find

FULL_PATTERN = "it"  # MMIMDbDataset.get_full_modality() (data/mmimdb.py)


class MaxOut(nn.Module):
    """maxout.py:6-41 -- element-wise max over ``num_units`` Linear layers (container; fused into the step)."""

    def __init__(self, input_dim: int, output_dim: int, num_units: int = 2, use_bias: bool = True) -> None:
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(input_dim, output_dim, bias=use_bias) for _ in range(num_units)])

    def forward(self, x):
        raise NotImplementedError("mml_b200.MaxOut is evaluated inside the fused MMIMDb step only")


class GatedBiModalNetwork(nn.Module):
    """gated_bimodal.py:6-60 -- GMU with one scalar gate per sample (container; fused into the step)."""

    def __init__(self, input_one_dim: int, input_two_dim: int, output_one_dim: int, output_two_dim: int, *, use_bias: bool = False):
        super().__init__()
        if use_bias:
            raise NotImplementedError("mml_b200 GatedBiModalNetwork implements the reference default use_bias=False")
        self.fc_one = nn.Linear(input_one_dim, output_one_dim, bias=False)
        self.fc_two = nn.Linear(input_two_dim, output_two_dim, bias=False)
        self.hidden_sigmoid = nn.Linear(output_one_dim + output_two_dim, 1, bias=False)
        self.activation = nn.Tanh()
        self.gate_activation = nn.Sigmoid()

    def forward(self, modality_one, modality_two):
        raise NotImplementedError("mml_b200.GatedBiModalNetwork is evaluated inside the fused MMIMDb step only")


class MultimodalPooling(nn.Module):
    """pooling.py:6-126 -- tanh(proj) of both embeddings, dropout, then an element-wise max / average / sum, or a per-sample
    mix whose weights come from a small attention / gate network on [a | b] (container; fused into the step)."""

    def __init__(self, input_dim_a: int, input_dim_b: int, output_dim: int, pooling_type: str = "gated", hidden_dim: Optional[int] = None,
                 dropout: float = 0.0):
        super().__init__()
        self.pooling_type = pooling_type.lower()
        if self.pooling_type not in ("max", "avg", "average", "sum", "attention", "gated"):
            raise ValueError(f"Unknown pooling type: {pooling_type}")
        self.input_dim_a, self.input_dim_b, self.output_dim = input_dim_a, input_dim_b, output_dim
        self.hidden_dim = hidden_dim or max(input_dim_a, input_dim_b)
        self.dropout = dropout
        self.proj_a = nn.Linear(input_dim_a, output_dim)
        self.proj_b = nn.Linear(input_dim_b, output_dim)
        self.dropout_layer = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        self.activation = nn.Tanh()
        if self.pooling_type == "attention":  # pooling.py:55-63
            self.attention_layer = nn.Sequential(nn.Linear(output_dim * 2, self.hidden_dim), nn.Tanh(), nn.Linear(self.hidden_dim, 2), nn.Softmax(dim=1))
        elif self.pooling_type == "gated":    # pooling.py:65-72
            self.gate_layer = nn.Sequential(nn.Linear(output_dim * 2, self.hidden_dim), nn.Tanh(), nn.Linear(self.hidden_dim, 1), nn.Sigmoid())

    def forward(self, x_a, x_b):
        raise NotImplementedError("mml_b200.MultimodalPooling is evaluated inside the fused MMIMDb step only")


class MLPGenreClassifier(nn.Module):
    """mmimdb.py:20-60."""

    def __init__(self, input_size: int, output_size: int, hidden_size: int) -> None:
        super().__init__()
        self.input_size, self.output_size, self.hidden_size = input_size, output_size, hidden_size
        self.net = nn.Sequential(
            nn.BatchNorm1d(input_size), MaxOut(input_size, hidden_size, use_bias=False), nn.Dropout(p=0.5),
            nn.BatchNorm1d(hidden_size), MaxOut(hidden_size, hidden_size, use_bias=False), nn.Dropout(p=0.5),
            nn.BatchNorm1d(hidden_size), nn.Linear(hidden_size, output_size))

    def forward(self, tensor):
        raise NotImplementedError("mml_b200.MLPGenreClassifier is evaluated inside the fused MMIMDb step only")


class MMIMDbModalityEncoder(nn.Module):
    """mmimdb.py:63-92 -- BatchNorm1d(input_dim) -> Linear(input_dim, output_dim)."""

    def __init__(self, input_dim: int, output_dim: int) -> None:
        super().__init__()
        self.net = nn.Sequential(nn.BatchNorm1d(input_dim), nn.Linear(input_dim, output_dim))
        self._mml_owner = None  # (weakref to the parent MMIMDb, "image" | "text")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        owner = self._mml_owner[0]() if self._mml_owner is not None else None
        if owner is None:
            raise NotImplementedError("mml_b200.MMIMDbModalityEncoder runs as part of an MMIMDb model or inside mml_b200.mono.MonomodalEncoder "
                                      "(monomodal pre-training)")
        return owner.encode(self._mml_owner[1], x)


class MMIMDb(nn.Module):
    def __init__(self, image_encoder: MMIMDbModalityEncoder, text_encoder: MMIMDbModalityEncoder,
                 gated_bimodal_network: Optional[GatedBiModalNetwork] = None, multimodal_pooling: Optional[Dict[str, Any]] = None,
                 classifier: MLPGenreClassifier = None, binary_threshold: float = 0.5) -> None:
        super().__init__()
        self.image_model = image_encoder
        self.text_model = text_encoder
        if multimodal_pooling is not None:  # mmimdb.py:132-146
            self.fusion_module = MultimodalPooling(
                input_dim_a=image_encoder.net[-1].out_features, input_dim_b=text_encoder.net[-1].out_features, output_dim=classifier.input_size,
                pooling_type=multimodal_pooling.get("pooling_type", "gated"), hidden_dim=multimodal_pooling.get("hidden_dim", None),
                dropout=multimodal_pooling.get("dropout", 0.0))
            self.fusion_type = "pooling"
        elif gated_bimodal_network is not None:
            self.fusion_module = gated_bimodal_network
            self.fusion_type = "gated"
        else:
            raise ValueError("Either gated_bimodal_network or multimodal_pooling must be provided")
        self.mm_mlp = classifier
        self.binary_threshold = binary_threshold
        self.monitor = None
        self._engine = None
        self._dp = None
        self.world_size = 1
        import weakref
        image_encoder._mml_owner = (weakref.ref(self), "image")
        text_encoder._mml_owner = (weakref.ref(self), "text")

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

    # ---- engine plumbing --------------------------------------------------------------------------------------------
    def _get_engine(self, device):
        from .gated_engine import GatedFusionEngine

        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("mml_b200.MMIMDb runs on a B200 GPU only: there is no CPU / PyTorch fallback path")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        eng = self._engine
        if eng is None or eng.device != device:
            eng = self._engine = GatedFusionEngine(self, device)
            if self._dp is not None:
                self._dp.attach(eng)
        eng.fs.ensure_fresh()
        return eng

    def flatten_parameters(self) -> None:
        p = next(self.parameters())
        if p.is_cuda:
            self._get_engine(p.device)

    def enable_data_parallel(self, dp) -> None:
        self._dp = dp
        self.world_size = dp.world_size
        if self._engine is not None:
            dp.attach(self._engine)

    def get_encoder(self, modality) -> MMIMDbModalityEncoder:
        name = str(modality).lower().split(".")[-1]
        if name == "image":
            return self.image_model
        if name == "text":
            return self.text_model
        raise ValueError(f"Invalid modality: {modality}. Must be one of image, text")

    def logits_transform(self, logits: torch.Tensor) -> np.ndarray:
        predictions = torch.sigmoid(logits).detach().cpu().numpy()
        return (predictions > self.binary_threshold).astype(int)

    # ---- staging ----------------------------------------------------------------------------------------------------
    def _stage(self, eng, I, T, mask_i=None, mask_t=None, labels=None):
        if I.dim() != 2 or T.dim() != 2 or I.shape[0] != T.shape[0]:
            raise ValueError(f"expected I [B,{'D'}] and T [B,D], got {tuple(I.shape)} / {tuple(T.shape)}")
        B = I.shape[0]
        plan = eng.plan_for(B)
        if I.shape[1] != plan.DI or T.shape[1] != plan.DT:
            raise ValueError(f"feature widths {I.shape[1]}/{T.shape[1]} do not match the encoders ({plan.DI}/{plan.DT})")
        plan.threshold = float(self.binary_threshold)
        _copy_in(plan.xI, I)
        _copy_in(plan.xT, T)
        for dst, m in ((plan.mI, mask_i), (plan.mT, mask_t)):
            if m is None:
                dst.fill_(1.0)
            else:
                _copy_in(dst, torch.as_tensor(m).reshape(B))
        if labels is not None:
            _copy_in(plan.labels, labels.reshape(B, plan.NC))
        from .data import note_inputs_consumed
        note_inputs_consumed(eng.device)  # a prefetcher may overwrite the batch's device buffers from here on
        return plan

    def forward(self, I: torch.Tensor, T: torch.Tensor, *, is_embd_I: bool = False, is_embd_T: bool = False) -> torch.Tensor:
        """Logits [B, classes] fp32 (mmimdb.py:166-200).  No autograd graph: training goes through ``train_step``."""
        assert not all((I is None, T is None)), "At least one modality must be provided"
        assert not all((is_embd_I, is_embd_T)), "Cannot both be embeddings"
        if is_embd_I or is_embd_T or I is None or T is None:
            raise NotImplementedError("mml_b200.MMIMDb.forward needs both raw modalities (pre-computed embeddings are outside the hot path)")
        eng = self._get_engine(I.device if I.is_cuda else next(self.parameters()).device)
        plan = self._stage(eng, I, T)
        if self.training:
            plan.run_forward_train_mode()
        else:
            plan.run_eval(with_loss=False)
        return plan.logits.clone()

    def encode(self, which: str, x: torch.Tensor) -> torch.Tensor:
        """One encoder's embedding [B, E] (eval: running statistics; train: batch statistics, running stats updated)."""
        eng = self._get_engine(x.device if x.is_cuda else next(self.parameters()).device)
        plan = eng.plan_for(x.shape[0])
        return plan.encode(which, x.float(), self.training)

    # ---- MultimodalModelProtocol ------------------------------------------------------------------------------------
    def _unpack(self, batch: Dict[Any, Any]):
        """(image, text, image_mask, text_mask, labels, pattern_name).  Reference contract (data/mmimdb.py:170-198): already
        masked tensors under Modality.IMAGE / Modality.TEXT; extension: ``<mod>_original`` + ``<mod>_missing_index`` => the
        x * mask of base_dataset.py:71 is fused into the first kernel."""
        I, T = _find(batch, "image"), _find(batch, "text")
        mask_i = mask_t = None
        if "image_original" in batch and "image_missing_index" in batch:
            I, mask_i = batch["image_original"], batch["image_missing_index"]
        if "text_original" in batch and "text_missing_index" in batch:
            T, mask_t = batch["text_original"], batch["text_missing_index"]
        if I is None or T is None:
            raise KeyError("batch needs image and text tensors (Modality.IMAGE / Modality.TEXT or *_original + *_missing_index)")
        return I, T, mask_i, mask_t, batch["label"], batch.get("pattern_name")

    def train_step(self, batch: Dict[Any, Any], optimizer: torch.optim.Optimizer, loss_functions, device, metric_recorder=None,
                   epoch: Optional[int] = None, **kwargs) -> Dict[str, Any]:
        """One fused training step; returns {"loss": float} like mmimdb.py:202-245."""
        eng = self._get_engine(device)
        self._check_loss(loss_functions)
        I, T, mask_i, mask_t, labels, miss_type = self._unpack(batch)
        self._set_mode(True)
        fs = eng.fs
        fs.adopt_optimizer(optimizer)
        fs.sync_hyper(optimizer, 1.0 / self.world_size)
        plan = self._stage(eng, I, T, mask_i, mask_t, labels)
        given = kwargs.get("dropout_masks")
        if given is not None:
            plan.keep1.copy_(torch.as_tensor(given[0]).reshape(plan.keep1.shape).to(torch.uint8), non_blocking=True)
            plan.keep2.copy_(torch.as_tensor(given[1]).reshape(plan.keep2.shape).to(torch.uint8), non_blocking=True)
            pool = kwargs.get("pool_masks")
            if plan.pooling and plan.pool_p > 0:
                if pool is None:
                    raise ValueError("dropout_masks given for a pooling model with dropout: pool_masks=(mask_a, mask_b) is needed too")
                plan.keepA.copy_(torch.as_tensor(pool[0]).reshape(plan.keepA.shape).to(torch.uint8), non_blocking=True)
                plan.keepB.copy_(torch.as_tensor(pool[1]).reshape(plan.keepB.shape).to(torch.uint8), non_blocking=True)
        plan.train_step(given_dropout=given is not None)
        fs._host_step += 1
        return self._finish(eng, plan, labels, miss_type, metric_recorder, False)

    def validation_step(self, batch: Dict[Any, Any], loss_functions, device, metric_recorder=None, return_test_info: bool = False,
                        epoch: Optional[int] = None, **kwargs) -> Dict[str, Any]:
        eng = self._get_engine(device)
        self._check_loss(loss_functions)
        I, T, mask_i, mask_t, labels, miss_type = self._unpack(batch)
        self._set_mode(False)
        plan = self._stage(eng, I, T, mask_i, mask_t, labels)
        plan.run_eval(with_loss=True)
        return self._finish(eng, plan, labels, miss_type, metric_recorder, return_test_info)

    def _finish(self, eng, plan, labels, miss_type, metric_recorder, return_test_info):
        plan.h_loss.copy_(plan.loss, non_blocking=True)
        if metric_recorder is not None or return_test_info:
            plan.h_pred.copy_(plan.pred, non_blocking=True)
        torch.cuda.current_stream(eng.device).synchronize()
        loss = float(plan.h_loss[0])
        if metric_recorder is None and not return_test_info:
            return {"loss": loss}
        predictions = plan.h_pred.numpy().astype(int)
        targets = labels.detach().cpu().numpy() if torch.is_tensor(labels) else np.asarray(labels)
        mt = np.array(miss_type)
        if metric_recorder is not None:
            metric_recorder.update_group_all("classification", predictions=predictions, targets=targets, m_types=mt)
        if return_test_info:
            return {"loss": loss, "predictions": predictions, "labels": targets, "miss_types": mt}
        return {"loss": loss}

    def get_embeddings(self, dataloader, device) -> Dict[Any, Any]:
        """mmimdb.py:296-338: per-modality embeddings of the fully-available samples."""
        embeddings = defaultdict(list)
        self._set_mode(False)
        self._get_engine(device)
        for batch in dataloader:
            I, T = _find(batch, "image"), _find(batch, "text")
            keep = torch.from_numpy(np.array(batch["pattern_name"]) == FULL_PATTERN)
            I, T = I[keep].to(device).float(), T[keep].to(device).float()
            if I.shape[0] == 0:
                continue
            ki = next((k for k in batch if str(k).lower().endswith("image")), "image")
            kt = next((k for k in batch if str(k).lower().endswith("text")), "text")
            embeddings[ki].append(self.image_model(I).cpu().numpy())
            embeddings[kt].append(self.text_model(T).cpu().numpy())
            embeddings["label"] += batch["label"]
        return embeddings

    @staticmethod
    def _check_loss(loss_functions) -> None:
        """The fused tail implements what the YAML resolves to: one BCEWithLogitsLoss() term, weight 1.0 (loss.py:52)."""
        if loss_functions is None:
            return
        items = list(loss_functions.items()) if hasattr(loss_functions, "items") else None
        if not items or len(items) != 1:
            raise NotImplementedError("mml_b200 fused MMIMDb step supports a LossFunctionGroup with exactly one bce_with_logits term")
        term = items[0][1]
        fn, weight = getattr(term, "loss_fn", term), float(getattr(term, "weight", 1.0))
        ok = isinstance(fn, nn.BCEWithLogitsLoss) and fn.reduction == "mean" and fn.weight is None and fn.pos_weight is None and weight == 1.0
        if not ok:
            raise NotImplementedError("mml_b200 fused MMIMDb step implements BCEWithLogitsLoss() with default arguments and weight 1.0 only")

    def __str__(self) -> str:
        return f"{self.image_model}\n{self.text_model}\n{self.fusion_module}\n{self.mm_mlp}"
