"""Host side of the missing-modality data path: pattern probabilities, per-sample masks and the device prefetcher.

Reference (MML_Suite): ``MissingPatternConfig.generate_patterns`` (config/data_config.py:58-106) turns per-modality missing
rates into ``pattern -> {modality: P(present)}``; ``MultimodalBaseDataset._initialise_missing_masks``
(data/base_dataset.py:46-59) draws one 0/1 mask per (sample, modality) from those probabilities at dataset construction;
``get_samples`` (:61-74) multiplies ``original * mask`` on the CPU inside the DataLoader worker.  Here the multiply is fused
into the first kernel of each encoder (``<mod>_original`` + ``<mod>_missing_index`` batches, see avmnist.py / mmimdb.py), so
the host only has to produce the masks -- and to get the batch onto the device without stalling the step, which is what
``DevicePrefetcher`` does (the reference's loop does a blocking ``.to(device)`` inside ``train_step``).

``create_missing_mask`` itself lives in the external ``modalities`` package that is not part of the reference tree; the
draw below is its documented behaviour (independent Bernoulli(P(present)) per sample and modality).
"""
from __future__ import annotations

from collections import OrderedDict, deque
from itertools import chain, combinations
from typing import Any, Dict, Iterable, Iterator, Mapping, Optional, Sequence, Tuple

import torch


def pattern_name(modalities: Iterable[str]) -> str:
    """Pattern key of a modality subset: first letters, sorted (data_config.py:50-53, 75)."""
    return "".join(sorted(str(m).lower().split(".")[-1][0] for m in modalities))


def generate_patterns(modalities: "Mapping[str, Tuple[float, Optional[Sequence[str]]]]",
                      selected_patterns: Optional[Sequence[str]] = None) -> "OrderedDict[str, Dict[str, float]]":
    """pattern -> {modality: P(present)}.  ``modalities``: name -> (missing_rate, apply_to or None).

    Every non-empty subset S of the modalities is a pattern; a member of S is present with probability 1 (or
    ``round(1 - rate, 4)`` if the pattern is listed in that modality's ``apply_to``), a non-member with probability 0;
    the full pattern always uses ``round(1 - rate, 4)`` for every modality; ``selected_patterns`` filters the result.
    """
    names = [str(m).lower().split(".")[-1] for m in modalities]
    spec = {n: v for n, v in zip(names, modalities.values())}
    subsets = sorted(chain.from_iterable(combinations(names, r) for r in range(1, len(names) + 1)), key=lambda c: (len(c), c))
    out: "OrderedDict[str, Dict[str, float]]" = OrderedDict()
    for subset in subsets:
        key = pattern_name(subset)
        probs = {}
        for n in names:
            rate, apply_to = spec[n]
            if n not in subset:
                probs[n] = 0.0
            else:
                probs[n] = round(1.0 - rate, 4) if (apply_to is not None and key in apply_to) else 1.0
        out[key] = probs
    out[pattern_name(names)] = {n: round(1.0 - spec[n][0], 4) for n in names}
    if selected_patterns:
        keep = {"".join(sorted(p)) for p in selected_patterns}
        out = OrderedDict((k, v) for k, v in out.items() if k in keep)
    return out


def draw_missing_masks(patterns: "Mapping[str, Mapping[str, float]]", num_samples: int,
                       generator: Optional[torch.Generator] = None) -> Dict[str, Dict[str, torch.Tensor]]:
    """pattern -> {modality: fp32 [num_samples] of 0/1}, drawn once (the reference draws at dataset construction and
    indexes by sample, base_dataset.py:46-59, data/avmnist.py:193-224)."""
    masks: Dict[str, Dict[str, torch.Tensor]] = {}
    for pat, probs in patterns.items():
        masks[pat] = {m: torch.bernoulli(torch.full((num_samples,), float(p)), generator=generator) for m, p in probs.items()}
    return masks


def attach_masks(batch: Dict[Any, Any], masks: Mapping[str, torch.Tensor], modalities: Sequence[str]) -> Dict[Any, Any]:
    """Turn a batch of ORIGINAL tensors into the ``<mod>_original`` + ``<mod>_missing_index`` form the fused step consumes."""
    out = dict(batch)
    for m in modalities:
        if m in out:
            out[f"{m}_original"] = out.pop(m)
        out[f"{m}_missing_index"] = masks[m]
    return out


class DevicePrefetcher:
    """Iterate a loader of batch dicts with the host->device copies of batch n+1 running under step n.

    Tensors go to fixed per-slot device buffers (pinned host memory makes the copies truly asynchronous: use
    ``DataLoader(pin_memory=True)``); non-tensor entries pass through.  The consumer's stream waits on the slot's copy event,
    and a slot is only overwritten after the consumer's stream has been waited on, so there is no allocator traffic and no race.

    The copies run on ONE dedicated high-priority stream per device (created once, shared by every prefetcher of that device).
    Round 1 rotated over candidate streams and kept the fastest from three samples each; under data parallelism that calibration
    picked a stream that serialised with the step about half of the time (prefetched 5.6 ms vs blocking 4.0 ms at N = 4).  The
    aliasing it tried to dodge -- CUDA multiplexing streams onto 8 hardware queues -- is removed at the source instead:
    ``CUDA_DEVICE_MAX_CONNECTIONS=32`` is set before CUDA initialises (mml_b200/__init__.py), and a high-priority stream keeps the
    small H2D transfers from queueing behind the step's kernels.
    """

    _streams: Dict[Any, "torch.cuda.Stream"] = {}

    def __init__(self, loader: Iterable[Dict[Any, Any]], device, depth: int = 1):
        self.loader, self.device, self.depth = loader, torch.device(device), max(1, int(depth))
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher stages batches onto a CUDA device")
        key = self.device.index if self.device.index is not None else torch.cuda.current_device()
        if key not in DevicePrefetcher._streams:
            DevicePrefetcher._streams[key] = torch.cuda.Stream(device=self.device, priority=-1)
        self.stream = DevicePrefetcher._streams[key]
        self.slots = [dict() for _ in range(self.depth + 1)]
        self.h2d_bytes = 0

    def _stage(self, batch: Dict[Any, Any], slot: Dict[Any, torch.Tensor]):
        main = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(main)  # the slot's previous consumer is done before it is overwritten
        out = {}
        with torch.cuda.stream(self.stream):
            for k, v in batch.items():
                if torch.is_tensor(v):
                    buf = slot.get(k)
                    if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                        buf = slot[k] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                    buf.copy_(v, non_blocking=True)
                    self.h2d_bytes += v.numel() * v.element_size()
                    out[k] = buf
                else:
                    out[k] = v
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return out, ev

    def __iter__(self) -> Iterator[Dict[Any, Any]]:
        it = iter(self.loader)
        queue: deque = deque()
        n = 0
        for _ in range(self.depth):
            try:
                queue.append(self._stage(next(it), self.slots[n % len(self.slots)]))
                n += 1
            except StopIteration:
                break
        while queue:
            out, ev = queue.popleft()
            try:
                queue.append(self._stage(next(it), self.slots[n % len(self.slots)]))
                n += 1
            except StopIteration:
                pass
            torch.cuda.current_stream(self.device).wait_event(ev)
            yield out
