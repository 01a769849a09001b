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

Device side (csrc/staging.cu): ``DeviceMaskTable`` draws the whole table on the GPU (counter-based Philox4x32-10: the bits depend
on (seed, pattern, modality, sample) only, so every data-parallel rank and the CPU oracle agree) and gathers a batch's masks by
sample index; ``luma_lut`` + ``DevicePrefetcher(luts=...)`` turn the reference's per-item ``uint8 -> gist_earth -> "L" -> float32``
image conversion (data/avmnist.py:188-191) into one byte per pixel over PCIe and a 256-entry table lookup on the device.
"""
from __future__ import annotations

from collections import OrderedDict, deque
from itertools import chain, combinations
from typing import Any, Dict, Iterable, Iterator, Mapping, Optional, Sequence, Tuple

import numpy as np
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


def _philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Philox4x32-10 (Salmon et al., SC'11) on uint64 numpy arrays holding 32-bit words; returns the four output words."""
    m0, m1, lo = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    sh = np.uint64(32)
    for r in range(10):
        ka, kb = np.uint64((k0 + r * 0x9E3779B9) & 0xFFFFFFFF), np.uint64((k1 + r * 0xBB67AE85) & 0xFFFFFFFF)
        a, b = m0 * c0, m1 * c2  # 32 x 32 -> 64 bit products
        c0, c1, c2, c3 = (b >> sh) ^ c1 ^ ka, b & lo, (a >> sh) ^ c3 ^ kb, a & lo
    return c0, c1, c2, c3


def philox_missing_masks(patterns: "Mapping[str, Mapping[str, float]]", num_samples: int, seed: int,
                         first_sample: int = 0) -> Dict[str, Dict[str, torch.Tensor]]:
    """The table ``DeviceMaskTable(patterns, num_samples, seed, device)`` draws on the GPU (``mml_missing_mask_draw``), computed on the host:
    pattern k, modality m, sample i is present iff the 24-bit uniform from word ``i % 4`` of the Philox block with counter
    ``(lo32(i // 4), hi32(i // 4), m, k)`` and key ``(lo32(seed), hi32(seed))`` is below P(present).  A pure function of its arguments, so
    the host datasets, every data-parallel rank and the device table agree bit for bit (``first_sample``: a shard of the sample range)."""
    i = np.arange(first_sample, first_sample + int(num_samples), dtype=np.uint64)
    q, word = i >> np.uint64(2), (i & np.uint64(3)).astype(np.int64)
    c0, c1 = q & np.uint64(0xFFFFFFFF), q >> np.uint64(32)
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    out: Dict[str, Dict[str, torch.Tensor]] = {}
    for k, (pat, probs) in enumerate(patterns.items()):
        out[pat] = {}
        for m, (mod, p) in enumerate(probs.items()):
            words = np.stack(_philox4x32_10(c0, c1, np.full_like(q, m), np.full_like(q, k), k0, k1))  # [4, n]
            bits = words[word, np.arange(i.size)]
            u = (bits >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24)
            out[pat][mod] = torch.from_numpy((u < np.float32(p)).astype(np.float32))
    return out


def attach_masks(batch: Dict[Any, Any], masks: Mapping[str, torch.Tensor], modalities: Sequence[str]) -> Dict[Any, Any]:
    """Turn a batch of ORIGINAL tensors into the ``<mod>_original`` + ``<mod>_missing_index`` form the fused step consumes."""
    out = dict(batch)
    for m in modalities:
        if m in out:
            out[f"{m}_original"] = out.pop(m)
        out[f"{m}_missing_index"] = masks[m]
    return out


class DeviceMaskTable:
    """pattern -> fp32 [n_modalities, num_samples] masks resident on the device, drawn there (``mml_missing_mask_draw``).

    ``masks[pattern]`` plays the role of ``MultimodalBaseDataset.masks[pattern]`` (base_dataset.py:46-59); ``batch(pattern, idx)``
    is the per-item lookup of ``__getitem__`` for a whole batch: {modality: fp32 [B]} views of one gathered [n_modalities, B] buffer.
    The draw is a pure function of (seed, pattern index, modality index, sample index)."""

    def __init__(self, patterns: "Mapping[str, Mapping[str, float]]", num_samples: int, seed: int, device):
        from . import ops

        self.device = torch.device(device)
        self.num_samples, self.seed = int(num_samples), int(seed)
        self.patterns = list(patterns)
        self.modalities = {pat: list(probs) for pat, probs in patterns.items()}
        self.masks: Dict[str, torch.Tensor] = {}
        self.bad = torch.zeros(1, dtype=torch.int32, device=self.device)
        for k, (pat, probs) in enumerate(patterns.items()):
            p = torch.tensor([float(v) for v in probs.values()], dtype=torch.float32).to(self.device)
            self.masks[pat] = ops.missing_mask_draw(p, self.num_samples, self.seed, stream_id=k)

    def as_dict(self, pattern: str) -> Dict[str, torch.Tensor]:
        return {m: self.masks[pattern][j] for j, m in enumerate(self.modalities[pattern])}

    def batch(self, pattern: str, sample_idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        from . import ops

        idx = sample_idx.to(self.device, dtype=torch.int64, non_blocking=True).contiguous()
        g = ops.missing_mask_gather(self.masks[pattern], idx, out, self.bad)
        return {m: g[j] for j, m in enumerate(self.modalities[pattern])}

    def check_indices(self) -> None:
        """Raises if any ``batch`` call so far carried a sample index outside [0, num_samples) (one host sync; call it per epoch)."""
        if int(self.bad.item()):
            raise IndexError("DeviceMaskTable.batch: a sample index was outside [0, num_samples)")


def luma_lut(cmap_table, scale: str = "mul") -> torch.Tensor:
    """fp32 [256] table of ``AVMNIST._load_image`` (data/avmnist.py:188-191) for uint8 pixels.

    ``cmap_table``: the colormap's 256 colours as floats in [0, 1], shape [256, 3 or 4] -- for the reference:
    ``matplotlib.cm.gist_earth(np.arange(256))`` (matplotlib indexes the table directly for integer pixels).  Each entry goes through
    the same arithmetic as the reference: ``np.uint8(colour * 255)`` (truncation), PIL's "L" conversion
    ``(19595 R + 38470 G + 7471 B + 0x8000) >> 16`` (alpha ignored), then fp32 scaling -- ``scale="mul"``: torchvision's
    ``ToDtype(float32, scale=True)`` = ``x * (1/255)``; ``scale="div"``: ``x.float() / 255.0`` (train_monomodal.py:55-62)."""
    t = np.asarray(cmap_table, dtype=np.float64)
    if t.ndim != 2 or t.shape[0] != 256 or t.shape[1] not in (3, 4):
        raise ValueError(f"expected a [256, 3|4] colour table, got {t.shape}")
    rgb = np.uint8(t[:, :3] * 255).astype(np.uint32)
    lum = ((rgb[:, 0] * 19595 + rgb[:, 1] * 38470 + rgb[:, 2] * 7471 + 0x8000) >> 16).astype(np.float32)
    if scale == "mul":
        lut = lum * np.float32(1.0 / 255.0)
    elif scale == "div":
        lut = lum / np.float32(255.0)
    else:
        raise ValueError("scale must be 'mul' (torchvision ToDtype) or 'div' (x / 255.0)")
    return torch.from_numpy(lut.astype(np.float32))


# The consumer of a prefetched batch (every fused ``train_step`` / ``validation_step``) copies it into its static input buffers first thing;
# right after those copies it records an event here.  The prefetcher waits for THAT event before it overwrites the slot, instead of for
# everything queued on the consumer's stream (= the whole previous step, collectives included).
_inputs_consumed: Dict[int, "torch.cuda.Event"] = {}


def note_inputs_consumed(device) -> None:
    """Called by the step methods once the batch has been copied into the plan's static buffers (stream-ordered on the current stream)."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))
    _inputs_consumed[idx] = ev


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

    def __init__(self, loader: Iterable[Dict[Any, Any]], device, depth: int = 1, luts: Optional[Mapping[Any, torch.Tensor]] = None):
        """``luts``: {batch key: fp32 [256] table}: a uint8 tensor under that key crosses PCIe as bytes and is expanded to fp32 on the
        device by ``mml_stage_u8_lut_f32`` on the copy stream (``luma_lut``: the AVMNIST image conversion)."""
        self.loader, self.device, self.depth = loader, torch.device(device), max(1, int(depth))
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher stages batches onto a CUDA device")
        self.luts = {k: v.to(self.device, dtype=torch.float32).contiguous() for k, v in (luts or {}).items()}
        import os as _os
        self.sync = _os.environ.get("MML_PREFETCH_SYNC", "event")  # "stream": wait for the consumer's whole stream (round-1 behaviour)
        key = self.device.index if self.device.index is not None else torch.cuda.current_device()
        if key not in DevicePrefetcher._streams:
            DevicePrefetcher._streams[key] = torch.cuda.Stream(device=self.device, priority=-1)
        self.stream = DevicePrefetcher._streams[key]
        self.slots = [dict() for _ in range(self.depth + 1)]
        self.h2d_bytes = 0

    def _stage(self, batch: Dict[Any, Any], slot: Dict[Any, torch.Tensor]):
        main = torch.cuda.current_stream(self.device)
        # the slot's previous consumer must be done with it before it is overwritten: its "inputs consumed" event when the step method
        # recorded one (consumed here, so a consumer that records nothing falls back to waiting for its whole stream)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        ev = _inputs_consumed.pop(idx, None) if self.sync == "event" else None
        if ev is not None:
            self.stream.wait_event(ev)
        else:
            self.stream.wait_stream(main)
        out = {}
        with torch.cuda.stream(self.stream):
            for k, v in batch.items():
                if torch.is_tensor(v):
                    buf = slot.get(k)
                    if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                        buf = slot[k] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                    buf.copy_(v, non_blocking=True)
                    self.h2d_bytes += v.numel() * v.element_size()
                    if k in self.luts and v.dtype == torch.uint8:
                        from . import ops

                        f = slot.get((k, "f32"))
                        if f is None or f.shape != v.shape:
                            f = slot[(k, "f32")] = torch.empty(v.shape, dtype=torch.float32, device=self.device)
                        ops.u8_lut(buf, self.luts[k], f)
                        buf = f
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
