"""The reference's dataset classes (AVMNIST, MMIMDb, MOSI / MOSEI), laid out for fused steps that consume 100 k samples/s.

AVMNIST first.  Reference: ``MML_Suite/data/avmnist.py:21-277`` (class ``AVMNIST``) on top of ``MultimodalBaseDataset``
(``data/base_dataset.py:16-154``) and ``PatternSpecificDataset`` (``data/pattern.py:6-19``): a CSV of
``audio`` / ``image`` / ``label`` columns whose cells are paths of ``torch.save``-d items; ``__getitem__`` loads both files
(lru-cached), runs the image through ``uint8 -> cm.gist_earth -> PIL "L" -> PILToTensor -> ToDtype(float32, scale=True)``,
looks the sample's masks up and multiplies (``get_samples``, base_dataset.py:61-74).  One item per Python call on a DataLoader
worker tops out four orders of magnitude below what the fused step eats (2.4 ms per 256 samples).

Here every file is read ONCE at construction into three contiguous host arrays -- audio fp32
``[N, H, W]``, image uint8 ``[N, h, w]`` (one byte per pixel: the whole image chain is a 256-entry table, ``data.luma_lut``),
labels int64 ``[N]`` -- plus the mask table ``[pattern][modality][len(self)]``.  Two views of that storage:

* ``__getitem__`` / ``collate_fn`` / ``get_pattern_batches``: the reference's item and batch dictionaries, bit for bit (same keys,
  shapes, dtypes, ordering of the validation / test splits, ``random.choice`` of the training pattern).  Pinned against the
  unmodified reference class by ``oracle/make_golden_loader.py`` (fixture ``tests/golden/avmnist_loader.npz``).
* ``batches()``: whole batches by vectorised row gathers into rotating pinned staging buffers, in the form the fused step takes --
  ``<mod>_original`` + ``<mod>_missing_index`` (the multiply runs in the first kernel of each encoder) with the image still uint8 --
  and ``fused_loader()`` = ``DevicePrefetcher(batches(), luts={"image_original": lut})``: 1 byte per image pixel over PCIe, table
  lookup on the copy stream (``mml_stage_u8_lut_f32``).

``MOSI`` / ``MOSEI`` (``MML_Suite/data/mosi.py:17-301``, one pickle of padded ``[N, T, F]`` arrays per split) get the same two views.  ``MMIMDb``
(``MML_Suite/data/mmimdb.py:14-207``, one HDF5 file per split) too: its items are pinned against the reference class with a dict-backed
stand-in for ``h5py`` (not installed in this image; the one ``h5py.File`` call is the only unexercised line).

Mask *sampling* follows ``data.draw_missing_masks`` (independent Bernoulli(P(present)); the reference's generator lives in the
un-vendored ``modalities`` package) or is handed in (``masks=``: e.g. ``DeviceMaskTable(...).masks`` copied back, or
the table of a reference run).
"""
from __future__ import annotations

import random
from itertools import combinations
from pathlib import Path
from typing import Any, Dict, Iterator, List, Mapping, Optional, Sequence, Union

import numpy as np
import torch
from torch.utils.data import Dataset

from .data import draw_missing_masks, luma_lut, philox_missing_masks

NAMES = ("audio", "image")


def _modality_keys(names: Sequence[str] = ("audio", "image")) -> Dict[str, Any]:
    """``modalities.Modality`` members when the package is importable (the reference's batch keys), else the lower-case names
    (``str(Modality.AUDIO) == "audio"``, so the ``<mod>_original`` / ``<mod>_missing_index`` keys are the same either way)."""
    try:
        from modalities import Modality  # type: ignore

        return {n: Modality.from_str(n) for n in tuple(names) + ("multimodal",)}
    except Exception:
        return {n: n for n in tuple(names) + ("multimodal",)}


def _name(modality: Any) -> str:
    return str(modality).lower().split(".")[-1]


def _colour_table(cmap: Any) -> np.ndarray:
    """[256, 3|4] float colours of ``cmap``: a table, or a matplotlib colormap (called on ``arange(256)``: integer input indexes the
    colormap's table directly, which is what ``cm.gist_earth(uint8 image)`` does, data/avmnist.py:189)."""
    if cmap is None:
        try:
            from matplotlib import cm  # type: ignore

            cmap = cm.gist_earth
        except Exception as e:  # pragma: no cover - depends on the image
            raise ImportError("AVMNIST needs the gist_earth colour table: install matplotlib or pass cmap=<[256, 3|4] table or colormap>") from e
    if callable(cmap):
        cmap = cmap(np.arange(256, dtype=np.uint8))
    t = np.asarray(cmap, dtype=np.float64)
    if t.ndim != 2 or t.shape[0] != 256 or t.shape[1] not in (3, 4):
        raise ValueError(f"expected a [256, 3|4] colour table, got {t.shape}")
    return t


def _maybe_pin(t: torch.Tensor, pin: bool) -> torch.Tensor:
    return t.pin_memory() if pin else t


class _PatternDataset(Dataset):
    """What the reference keeps in ``MultimodalBaseDataset`` (data/base_dataset.py:16-154) -- splits, pattern names, the mask table, the
    index -> (pattern, sample) rule -- plus the staging machinery of the batch-granular path.  Subclasses set ``MODS`` (modality names
    in ``AVAILABLE_MODALITIES`` order), ``DEFAULT_PATTERNS`` and store their arrays."""

    VALID_SPLITS: List[str] = ["train", "valid", "test"]
    MODS: tuple = ()
    DEFAULT_PATTERNS: Dict[str, Dict[str, float]] = {}
    AVAILABLE_MODALITIES: Dict[str, Any] = {}

    @classmethod
    def get_full_modality(cls) -> str:
        return "".join(sorted(k[0] for k in cls.MODS))

    @classmethod
    def get_all_possible_patterns(cls) -> List[str]:
        mods = list(cls.MODS)
        return sorted("".join(m[0] for m in sorted(c)) for r in range(1, len(mods) + 1) for c in combinations(mods, r))

    def validate_patterns(self, patterns: Sequence[str]) -> List[str]:
        bad = set(patterns) - set(self.get_all_possible_patterns())
        if bad:
            raise ValueError(f"Invalid patterns: {bad}\nValid patterns are: {self.get_all_possible_patterns()}")
        return list(patterns)

    def _configure(self, split, target_modality, missing_patterns, selected_patterns, _id=1) -> None:
        self.split = str(split).lower()
        assert split in self.VALID_SPLITS, f"Invalid split provided, must be one of {self.VALID_SPLITS}"
        assert isinstance(_id, int), "ID must be an integer."
        self._id = _id
        self.keys = _modality_keys(self.MODS)
        self.AVAILABLE_MODALITIES = {n: self.keys[n] for n in self.MODS}
        # pattern -> {modality name: P(present)}
        mp = missing_patterns or self.DEFAULT_PATTERNS
        self.missing_patterns = {pat: {_name(m): float(p) for m, p in probs.items()} for pat, probs in mp.items()}
        self.selected_patterns = self.validate_patterns(selected_patterns) if selected_patterns is not None else self.get_all_possible_patterns()
        for pat in self.selected_patterns:
            if pat not in self.missing_patterns:
                raise ValueError(f"selected pattern {pat!r} has no entry in missing_patterns {list(self.missing_patterns)}")
        self.current_pattern = None
        tm = _name(target_modality)
        assert tm in self.MODS + ("multimodal",), f"Invalid modality provided, must be one of {list(self.MODS) + ['multimodal']}"
        self.target_modality = self.keys[tm]
        self._target = tm

    def _loads(self, m: str) -> bool:
        return self._target in ("multimodal", m)

    def _finish(self, num_samples: int, masks, generator, pin, mask_seed: Optional[int] = None) -> None:
        """Mask table (given, or drawn -- from ``generator`` with torch.bernoulli, or from ``mask_seed`` with the counter-based Philox draw
        that ``data.DeviceMaskTable`` performs on the GPU: same table on the host, on the device and on every rank) once ``num_samples``
        is known."""
        self.num_samples = int(num_samples)
        self.pattern_indices = {pattern: list(range(self.num_samples)) for pattern in self.selected_patterns}
        self._pin = torch.cuda.is_available() if pin is None else bool(pin)
        if masks is not None:
            self.masks = {pat: {_name(m): torch.as_tensor(v, dtype=torch.float32).reshape(-1) for m, v in tab.items()} for pat, tab in masks.items()}
            for pat in self.missing_patterns:
                for m in self.MODS:
                    if pat not in self.masks or m not in self.masks[pat] or self.masks[pat][m].numel() < self.num_samples:
                        raise ValueError(f"masks[{pat!r}][{m!r}] must hold at least {self.num_samples} entries")
        elif mask_seed is not None:
            self.masks = philox_missing_masks(self.missing_patterns, len(self), int(mask_seed))
        else:
            # one draw per (pattern, modality, dataset index) at construction, length len(self) like base_dataset.py:46-59
            self.masks = draw_missing_masks(self.missing_patterns, len(self), generator)
        self.generator = generator
        # [pattern][modality][sample] as one tensor for the vectorised batch path
        self._pat_index = {pat: i for i, pat in enumerate(self.missing_patterns)}
        self._mask_table = torch.stack([torch.stack([self.masks[pat][m][: self.num_samples] for m in self.MODS]) for pat in self.missing_patterns])

    def __len__(self) -> int:
        return self.num_samples if self.split == "train" else self.num_samples * len(self.selected_patterns)

    def _get_pattern_and_sample_idx(self, idx: int):
        if self.split == "train" or self.split == "trn":
            return random.choice(self.selected_patterns), idx  # base_dataset.py:87-89: Python's global ``random``
        return self.selected_patterns[idx // self.num_samples], idx % self.num_samples

    def _item_head(self, idx: int):
        pattern, i = self._get_pattern_and_sample_idx(int(idx))
        if not 0 <= i < self.num_samples:
            raise IndexError(idx)
        self.current_pattern = pattern
        return pattern, i

    @staticmethod
    def _masked(sample: Dict[Any, Any], key: Any, m: str, original: torch.Tensor) -> None:
        """``get_samples`` (base_dataset.py:61-74): original, original * mask and the complementary ``_reverse`` tensor."""
        mask = sample[f"{m}_missing_index"]
        sample[f"{m}_original"] = original
        sample[key] = original * mask
        sample[f"{m}_reverse"] = original * -1 * (mask - 1)

    def get_split(self) -> str:
        return self.split

    def get_selected_patterns(self) -> List[str]:
        return self.selected_patterns

    def get_missing_patterns(self):
        return self.missing_patterns

    # ---- batch-granular machinery ------------------------------------------------------------------------------------------
    def _epoch(self, batch_size: int, shuffle: Optional[bool], drop_last: bool, pattern: Optional[str], generator, rotate: int,
               rank: int = 0, world: int = 1):
        """(staging set, rows int64 [B], pattern names, masks fp32 [B, n_modalities]) per batch.

        Order: the training split visits every sample once (shuffled unless ``shuffle=False``) with an independent uniformly drawn
        pattern per sample (the vectorised form of ``random.choice``, base_dataset.py:87-89; drawn from ``generator``, not from
        Python's ``random``); the other splits walk ``selected_patterns`` in order, all samples of one pattern after the other --
        dataset index order, base_dataset.py:90-93 -- or only ``pattern``.

        Data parallel (``world`` > 1): the epoch's order is laid out in GLOBAL batches of ``world * batch_size`` samples and rank
        ``r`` takes rows ``[r * batch_size, (r + 1) * batch_size)`` of each -- every rank must pass a generator in the same state (same
        seed), as with torch's ``DistributedSampler``.  Training needs ``drop_last=True`` then (a ragged last step would leave ranks with
        different step counts and hang the gradient all-reduce); evaluation splits the ragged tail into near-equal contiguous shards."""
        if batch_size < 1 or rotate < 1:
            raise ValueError("batch_size and rotate must be positive")
        if world < 1 or not 0 <= rank < world:
            raise ValueError(f"rank {rank} is not in [0, world = {world})")
        gen = generator if generator is not None else self.generator
        train = self.split == "train"
        if shuffle is None:
            shuffle = train
        N = self.num_samples
        if train:
            if pattern is not None:
                pats = torch.full((N,), self.selected_patterns.index(pattern), dtype=torch.long)
            else:
                pats = torch.randint(len(self.selected_patterns), (N,), generator=gen)
            rows = torch.randperm(N, generator=gen) if shuffle else torch.arange(N)
            pats = pats[rows] if shuffle else pats
        else:
            which = [self.selected_patterns.index(pattern)] if pattern is not None else range(len(self.selected_patterns))
            rows = torch.cat([torch.arange(N) for _ in which]) if len(which) else torch.empty(0, dtype=torch.long)
            pats = torch.cat([torch.full((N,), k, dtype=torch.long) for k in which]) if len(which) else rows
            if shuffle:
                perm = torch.randperm(rows.numel(), generator=gen)
                rows, pats = rows[perm], pats[perm]
        table_row = torch.tensor([self._pat_index[p] for p in self.selected_patterns], dtype=torch.long)
        bufs: List[Dict[str, torch.Tensor]] = [dict() for _ in range(rotate)]
        total = rows.numel()
        gb = batch_size * world  # one global batch
        if world > 1 and train and not drop_last and total % gb != 0:
            raise ValueError(f"data-parallel training over {total} samples in global batches of {gb} leaves a ragged last step: pass drop_last=True")
        stop = total - (total % gb) if drop_last else total
        n = 0
        for g0 in range(0, stop, gb):
            size = min(gb, stop - g0)
            if size == gb:
                lo, hi = g0 + rank * batch_size, g0 + (rank + 1) * batch_size
            else:  # ragged tail (evaluation, or one process): near-equal contiguous shards, the first ``size % world`` ranks take one more
                q, rem = divmod(size, world)
                lo = g0 + rank * q + min(rank, rem)
                hi = lo + q + (1 if rank < rem else 0)
                if hi == lo:
                    continue
            r, p = rows[lo:hi], pats[lo:hi]
            yield bufs[n % rotate], r, [self.selected_patterns[k] for k in p.tolist()], self._mask_table[table_row[p], :, r]
            n += 1

    def _staging(self, buf: Dict[str, torch.Tensor], key: str, shape, dtype) -> torch.Tensor:
        """View of ``shape`` on the pinned buffer ``buf[key]`` (allocated once at the largest leading dimension seen: the ragged last
        batch of an epoch reuses the full-size buffer)."""
        shape = tuple(shape)
        t = buf.get(key)
        if t is None or t.shape[1:] != torch.Size(shape[1:]) or t.dtype != dtype or t.shape[0] < shape[0]:
            t = buf[key] = _maybe_pin(torch.empty(shape, dtype=dtype), self._pin)
        return t[: shape[0]]

    def _gather(self, buf: Dict[str, torch.Tensor], key: str, src: Optional[torch.Tensor], rows: torch.Tensor) -> torch.Tensor:
        """rows of ``src`` (or ``rows`` itself when ``src`` is None) into the pinned staging tensor ``buf[key]``."""
        if src is None:
            t = self._staging(buf, key, rows.shape, rows.dtype)
            t.copy_(rows)
            return t
        t = self._staging(buf, key, (rows.numel(),) + tuple(src.shape[1:]), src.dtype)
        torch.index_select(src, 0, rows, out=t)
        return t

    def batches(self, batch_size: int, **kwargs) -> Iterator[Dict[Any, Any]]:  # pragma: no cover - abstract
        raise NotImplementedError

    def _luts(self, kwargs) -> Optional[Dict[Any, torch.Tensor]]:
        return None

    def background_batches(self, batch_size: int, ahead: int = 2, **kwargs) -> Iterator[Dict[Any, Any]]:
        """``batches()`` produced by a worker thread, ``ahead`` batches in front of the consumer (the row gathers are ``index_select``
        calls that release the GIL, so they overlap the step's host code instead of adding 1.4-3.3 ms per 256-sample AVMNIST batch to
        it).  ``rotate`` defaults to ``ahead + 3`` staging sets: ``ahead`` queued, one being filled, one with the consumer, one spare."""
        import queue
        import threading

        kwargs.setdefault("rotate", ahead + 3)
        if kwargs["rotate"] < ahead + 2:
            raise ValueError("rotate must be at least ahead + 2 (queued batches + the one being filled + the one in use)")
        q: "queue.Queue" = queue.Queue(maxsize=max(1, int(ahead)))
        stop = threading.Event()
        done = object()

        def put(item) -> bool:
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def work():
            try:
                for b in self.batches(batch_size, **kwargs):
                    if not put(b):
                        return
                put(done)
            except BaseException as e:  # surfaced in the consumer
                put(e)

        t = threading.Thread(target=work, name="mml-batches", daemon=True)
        t.start()
        try:
            while True:
                b = q.get()
                if b is done:
                    return
                if isinstance(b, BaseException):
                    raise b
                yield b
        finally:
            stop.set()
            t.join(timeout=5.0)

    def fused_loader(self, device, batch_size: int, depth: int = 1, ahead: int = 2, **kwargs):
        """``batches()`` (on a worker thread when ``ahead`` > 0) behind the device prefetcher; the result feeds the fused
        ``train_step / validation_step`` of the matching model directly.  AVMNIST: uint8 images cross PCIe as bytes and are expanded
        through the luminance table on the copy stream.  Staging sets: ``ahead`` + ``depth`` + 3 (queued + staged on the copy stream +
        filling / in use / spare)."""
        from .data import DevicePrefetcher

        kwargs.setdefault("rotate", ahead + depth + 3)
        luts = self._luts(kwargs)
        it = self.background_batches(batch_size, ahead=ahead, **kwargs) if ahead > 0 else self.batches(batch_size, **kwargs)
        return DevicePrefetcher(it, device, depth=depth, luts=luts)


class AVMNIST(_PatternDataset):
    """Same constructor and item / batch contract as ``data.avmnist.AVMNIST`` (data/avmnist.py:45-59), plus ``cmap`` (colour table),
    ``masks`` (explicit mask table), ``generator`` (mask draw / batch shuffling) and ``pin`` (staging buffers of the batch path in pinned
    memory; default: iff CUDA is available -- the dataset arrays themselves stay in ordinary host memory)."""

    NUM_CLASSES: int = 10
    MODS = NAMES
    AVAILABLE_MODALITIES: Dict[str, Any] = {"audio": "audio", "image": "image"}
    # the reference's default (data/avmnist.py:73-77)
    DEFAULT_PATTERNS = {"ai": {"audio": 1.0, "image": 1.0}, "a": {"audio": 1.0, "image": 0.0}, "i": {"audio": 0.0, "image": 1.0}}

    def __init__(self, data_fp: Union[str, Path], split: str, target_modality: Any = "multimodal", *,
                 missing_patterns: Optional[Mapping[str, Mapping[Any, float]]] = None, selected_patterns: Optional[Sequence[str]] = None,
                 audio_column: str = "audio", image_column: str = "image", labels_column: str = "label",
                 split_indices: Optional[Sequence[int]] = None, _id: int = 1, cmap: Any = None,
                 masks: Optional[Mapping[str, Mapping[Any, torch.Tensor]]] = None, generator: Optional[torch.Generator] = None,
                 pin: Optional[bool] = None, mask_seed: Optional[int] = None) -> None:
        import pandas as pd

        self._configure(split, target_modality, missing_patterns, selected_patterns, _id)
        self.data_fp = Path(data_fp)
        if not self.data_fp.exists():
            raise FileNotFoundError(f"Data file not found: {data_fp}")
        self.audio_column, self.image_column, self.labels_column = audio_column, image_column, labels_column
        self.data = pd.read_csv(self.data_fp)
        if split_indices is not None:
            self.data = self.data.iloc[list(split_indices)].reset_index(drop=True)
        missing_columns = [c for c in (audio_column, image_column, labels_column) if c not in self.data.columns]
        if missing_columns:
            raise ValueError(f"Missing required columns: {missing_columns}")
        labels = torch.from_numpy(np.array(self.data[labels_column].to_numpy(), dtype=np.int64))
        audio = self._read_audio(self.data[audio_column]) if self._loads("audio") else None
        image = self._read_images(self.data[image_column]) if self._loads("image") else None
        self._store(labels, audio, image, cmap, masks, generator, pin, mask_seed)

    @classmethod
    def from_arrays(cls, labels, audio=None, image_u8=None, split: str = "train", target_modality: Any = "multimodal", *,
                    missing_patterns=None, selected_patterns=None, _id: int = 1, cmap: Any = None, masks=None,
                    generator: Optional[torch.Generator] = None, pin: Optional[bool] = None, mask_seed: Optional[int] = None) -> "AVMNIST":
        """The same dataset over arrays that are already in memory (``audio`` fp32 [N, H, W], ``image_u8`` uint8 [N, h, w], ``labels``
        [N]) instead of a CSV of per-item files; ``None`` for a modality that the target does not load."""
        self = cls.__new__(cls)
        self._configure(split, target_modality, missing_patterns, selected_patterns, _id)
        self.data_fp = self.data = None
        labels = torch.as_tensor(labels, dtype=torch.long).reshape(-1).clone()
        a = i = None
        if self._loads("audio"):
            a = torch.as_tensor(audio, dtype=torch.float32).contiguous()
            if a.dim() != 3 or a.shape[0] != labels.numel():
                raise ValueError(f"audio must be [N, H, W] with N = {labels.numel()}, got {tuple(a.shape)}")
        if self._loads("image"):
            i = torch.as_tensor(image_u8)
            if i.dtype != torch.uint8 or i.dim() != 3 or i.shape[0] != labels.numel():
                raise TypeError(f"image_u8 must be uint8 [N, h, w] with N = {labels.numel()}, got {i.dtype} {tuple(i.shape)}")
            i = i.contiguous()
        self._store(labels, a, i, cmap, masks, generator, pin, mask_seed)
        return self

    def _store(self, labels, audio, image_u8, cmap, masks, generator, pin, mask_seed=None) -> None:
        self._finish(labels.numel(), masks, generator, pin, mask_seed)
        self.labels = labels
        self.audio = audio
        self.image_u8 = self.lut = None
        if image_u8 is not None:
            self.lut = luma_lut(_colour_table(cmap))  # fp32 [256]: the reference's image chain as a function of the pixel value
            self.image_u8 = image_u8

    # ---- file reading (once) ------------------------------------------------------------------------------------------
    @staticmethod
    def _read_audio(paths) -> torch.Tensor:
        """``torch.load(path, weights_only=True)`` per row (data/avmnist.py:174), stacked: spectrograms share one shape."""
        items = [torch.load(str(p), weights_only=True) for p in paths]
        if not items:
            return torch.empty(0, 0, 0)
        shape = items[0].shape
        for p, t in zip(paths, items):
            if t.shape != shape:
                raise ValueError(f"audio item {p} has shape {tuple(t.shape)}, expected {tuple(shape)} (batches are dense tensors)")
        return torch.stack(items).contiguous()

    @staticmethod
    def _read_images(paths) -> torch.Tensor:
        """``np.array(torch.load(path, weights_only=False))`` per row (data/avmnist.py:188); uint8 pixels stay uint8."""
        items = [np.array(torch.load(str(p), weights_only=False)) for p in paths]
        if not items:
            return torch.empty(0, 0, 0, dtype=torch.uint8)
        for p, a in zip(paths, items):
            if a.dtype != np.uint8 or a.ndim != 2:
                raise TypeError(f"image item {p}: expected a 2-D uint8 array (the shipped AVMNIST images), got {a.dtype} {a.shape}; "
                                "float images index the colormap differently and are not a table lookup")
            if a.shape != items[0].shape:
                raise ValueError(f"image item {p} has shape {a.shape}, expected {items[0].shape}")
        return torch.from_numpy(np.stack(items)).contiguous()

    # ---- reference item / batch contract ----------------------------------------------------------------------------------
    def image_float(self, rows) -> torch.Tensor:
        """fp32 [n, 1, h, w] images of ``rows`` = ``_load_image`` of the reference for each of them (table lookup on the host)."""
        return self.lut[self.image_u8[rows].long()].unsqueeze(-3)

    def __getitem__(self, idx: int) -> Dict[Any, Any]:
        pattern, i = self._item_head(idx)
        sample: Dict[Any, Any] = {"labels": self.labels[i].clone(), "pattern_name": pattern, "missing_mask": {}, "sample_idx": i}
        for m in NAMES:
            sample[f"{m}_missing_index"] = self.masks[pattern][m][i]
        for m in NAMES:
            if self._loads(m):
                self._masked(sample, self.keys[m], m, self.audio[i].clone() if m == "audio" else self.image_float(i))
        return sample

    def collate_fn(self, batch: List[Dict[Any, Any]]) -> Dict[Any, Any]:
        """data/avmnist.py:248-277: labels, pattern names, (empty) ``missing_masks`` and the MASKED tensors under the modality keys."""
        ka, ki = self.keys["audio"], self.keys["image"]
        collated: Dict[Any, Any] = {
            "labels": torch.stack([b["labels"] for b in batch]),
            "pattern_name": [b["pattern_name"] for b in batch],
            "missing_masks": {mod: torch.tensor([b["missing_mask"][mod] for b in batch]) for mod in (ka, ki) if mod in batch[0]["missing_mask"]},
        }
        if self._target == "multimodal":
            for mod in (ka, ki):
                if mod in batch[0]:
                    collated[mod] = torch.stack([b[mod] for b in batch])
        else:
            collated[self.target_modality] = torch.stack([b[self.target_modality] for b in batch])
        return collated

    def get_pattern_batches(self, batch_size: int, **dataloader_kwargs) -> Dict[str, Any]:
        """pattern -> DataLoader over that pattern's slice of a validation / test split (data/avmnist.py:226-246)."""
        from torch.utils.data import DataLoader

        if self.split == "train":
            raise ValueError("Pattern-specific batches only available for validation/test")
        return {pattern: DataLoader(PatternSpecificDataset(self, pattern), batch_size=batch_size, shuffle=False, collate_fn=self.collate_fn,
                                    **dataloader_kwargs) for pattern in self.selected_patterns}

    # ---- batch-granular path ------------------------------------------------------------------------------------------------
    def batches(self, batch_size: int, shuffle: Optional[bool] = None, drop_last: bool = False, pattern: Optional[str] = None,
                image_form: str = "u8", rotate: int = 4, generator: Optional[torch.Generator] = None, rank: int = 0,
                world: int = 1) -> Iterator[Dict[Any, Any]]:
        """Whole batches in the fused step's input form: ``labels`` int64 [B], ``pattern_name`` list, ``sample_idx`` int64 [B],
        ``audio_original`` fp32 [B, H, W], ``image_original`` uint8 [B, 1, h, w] (``image_form="u8"``: expand on the device with
        ``DevicePrefetcher(luts={"image_original": ds.lut})``) or fp32 (``"f32"``: table lookup on the host), ``<mod>_missing_index`` fp32 [B].
        Order: ``_PatternDataset._epoch``.  A yielded batch's tensors live in one of ``rotate`` pinned staging buffer sets and stay valid
        until ``rotate - 1`` further batches have been drawn (enough for a copy stream one batch ahead of the step).  ``rank`` / ``world``:
        this process's shard of every global batch under data parallelism (``_epoch``)."""
        if image_form not in ("u8", "f32"):
            raise ValueError("image_form must be 'u8' or 'f32'")
        for buf, r, names, m in self._epoch(batch_size, shuffle, drop_last, pattern, generator, rotate, rank, world):
            out: Dict[Any, Any] = {"pattern_name": names}
            out["labels"] = self._gather(buf, "labels", self.labels, r)
            out["sample_idx"] = self._gather(buf, "sample_idx", None, r)
            if self.audio is not None:
                out["audio_original"] = self._gather(buf, "audio", self.audio, r)
                out["audio_missing_index"] = self._gather(buf, "audio_mask", None, m[:, 0])
            if self.image_u8 is not None:
                img = self._gather(buf, "image", self.image_u8, r).unsqueeze(1)
                if image_form == "f32":
                    f = self._staging(buf, "image_f32", img.shape, torch.float32)
                    torch.index_select(self.lut, 0, img.reshape(-1).long(), out=f.view(-1))
                    img = f
                out["image_original"] = img
                out["image_missing_index"] = self._gather(buf, "image_mask", None, m[:, 1])
            yield out

    def _luts(self, kwargs):
        return {"image_original": self.lut} if self.image_u8 is not None and kwargs.get("image_form", "u8") == "u8" else None


class MultimodalSentimentDataset(_PatternDataset):
    """CMU-MOSI / CMU-MOSEI: same constructor and item contract as ``data.mosi.MultimodalSentimentDataset`` (data/mosi.py:17-202) -- one
    pickle ``{split: {"audio", "vision", "text", <labels_key>, "audio_lengths", "vision_lengths"}}`` of zero-padded ``[N, T, F]`` arrays.
    The reference's own ``collate_fn`` cannot run (``_collate_train_batch`` indexes ``b[""]``, data/mosi.py:231; its loaders use torch's
    default collation, ``DataConfig.use_collate_fn = False``), so none is mirrored: ``DataLoader(ds)`` with the default collation gives the
    reference's batches, ``batches()`` / ``fused_loader()`` give the fused step's form (``<mod>_original`` + ``<mod>_missing_index``)."""

    NUM_CLASSES: int = 3
    MODS = ("audio", "video", "text")
    AVAILABLE_MODALITIES: Dict[str, Any] = {"audio": "audio", "video": "video", "text": "text"}
    RAW_KEYS = {"audio": "audio", "video": "vision", "text": "text"}
    # the reference's default (data/mosi.py:62-70)
    DEFAULT_PATTERNS = {
        "atv": {"audio": 1.0, "text": 1.0, "video": 1.0}, "at": {"audio": 1.0, "text": 1.0, "video": 0.0},
        "av": {"audio": 1.0, "text": 0.0, "video": 1.0}, "tv": {"audio": 0.0, "text": 1.0, "video": 1.0},
        "a": {"audio": 1.0, "text": 0.0, "video": 0.0}, "t": {"audio": 0.0, "text": 1.0, "video": 0.0},
        "v": {"audio": 0.0, "text": 0.0, "video": 1.0},
    }

    def __init__(self, data_fp: Union[str, Path], split: str, target_modality: Any = "multimodal", *,
                 missing_patterns: Optional[Mapping[str, Mapping[Any, float]]] = None, selected_patterns: Optional[Sequence[str]] = None,
                 labels_key: str = "classification_labels", aligned: bool = False, length: Optional[int] = None,
                 num_classes: Optional[int] = None, batch_size: int = 1, masks=None, generator: Optional[torch.Generator] = None,
                 pin: Optional[bool] = None, mask_seed: Optional[int] = None) -> None:
        import pickle

        if num_classes is not None:
            self.NUM_CLASSES = num_classes
        self._configure(split, target_modality, missing_patterns, selected_patterns)
        self._batch_size = batch_size
        self.data_fp = Path(data_fp)
        self.aligned = aligned
        self.length = length if aligned else None
        self.labels_key = labels_key
        if not self.data_fp.exists():
            raise FileNotFoundError(f"Data file not found: {self.data_fp}")
        with open(self.data_fp, "rb") as f:
            raw_data = pickle.load(f)
        if self.split not in raw_data:
            raise KeyError(f"Split '{self.split}' not found in data")
        split_data = raw_data[self.split]
        if labels_key not in split_data:
            raise KeyError(f"Labels key '{labels_key}' not found in data")
        label = torch.tensor(split_data[labels_key], dtype=torch.float32 if "regression" in labels_key else torch.long)
        self.original_label_size = label.size(0)
        self._finish(len(label), masks, generator, pin, mask_seed)
        self.data: Dict[Any, torch.Tensor] = {"label": label}
        for m in self.MODS:  # the reference converts all three whatever the target modality (data/mosi.py:137-145)
            self.data[self.keys[m]] = torch.tensor(split_data[self.RAW_KEYS[m]]).float().contiguous()
        if not aligned:
            self.data["audio_lengths"] = torch.tensor(split_data["audio_lengths"]).float()
            self.data["video_lengths"] = torch.tensor(split_data["vision_lengths"]).float()

    def __getitem__(self, idx: int) -> Dict[Any, Any]:
        pattern, i = self._item_head(idx)
        sample: Dict[Any, Any] = {"label": self.data["label"][i], "pattern_name": pattern, "missing_index": {}, "sample_idx": i}
        for m in self.MODS:
            sample[f"{m}_missing_index"] = self.masks[pattern][m][i]
        if not self.aligned:
            sample["audio_length"] = self.data["audio_lengths"][i]
            sample["video_length"] = self.data["video_lengths"][i]
        for m in self.MODS:
            if self._loads(m):
                self._masked(sample, self.keys[m], m, self.data[self.keys[m]][i])
        return sample

    def batches(self, batch_size: int, shuffle: Optional[bool] = None, drop_last: bool = False, pattern: Optional[str] = None,
                rotate: int = 4, generator: Optional[torch.Generator] = None, rank: int = 0, world: int = 1) -> Iterator[Dict[Any, Any]]:
        """``label`` [B], ``pattern_name``, ``sample_idx``, ``<mod>_original`` fp32 [B, T, F] + ``<mod>_missing_index`` fp32 [B] for the
        loaded modalities and, for unaligned data, ``audio_length`` / ``video_length`` [B]; order and staging as in ``AVMNIST.batches``."""
        for buf, r, names, msk in self._epoch(batch_size, shuffle, drop_last, pattern, generator, rotate, rank, world):
            out: Dict[Any, Any] = {"pattern_name": names}
            out["label"] = self._gather(buf, "label", self.data["label"], r)
            out["sample_idx"] = self._gather(buf, "sample_idx", None, r)
            if not self.aligned:
                out["audio_length"] = self._gather(buf, "audio_length", self.data["audio_lengths"], r)
                out["video_length"] = self._gather(buf, "video_length", self.data["video_lengths"], r)
            for j, m in enumerate(self.MODS):
                if self._loads(m):
                    out[f"{m}_original"] = self._gather(buf, m, self.data[self.keys[m]], r)
                    out[f"{m}_missing_index"] = self._gather(buf, m + "_mask", None, msk[:, j])
            yield out

    @staticmethod
    def normalize_features(features: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
        """Zero mean / unit std along the time dimension (data/mosi.py:258-271)."""
        mean = torch.mean(features, dim=0, keepdim=True)
        std = torch.std(features, dim=0, keepdim=True).clamp(min=eps)
        return (features - mean) / std

    @staticmethod
    def get_num_classes(is_classification: bool = True) -> int:
        return 3 if is_classification else 1


class MOSI(MultimodalSentimentDataset):
    """CMU-MOSI (data/mosi.py:288-301)."""


class MOSEI(MultimodalSentimentDataset):
    """CMU-MOSEI (data/mosi.py:274-286)."""


def _open_h5(path):
    """``h5py.File(path, "r")`` (data/mmimdb.py:88); h5py is imported here so that the module loads without it."""
    try:
        import h5py  # type: ignore
    except Exception as e:
        raise ImportError("MMIMDb reads an HDF5 file: install h5py, or build the dataset with MMIMDb.from_arrays(...)") from e
    return h5py.File(Path(path), "r")


class MMIMDb(_PatternDataset):
    """MM-IMDb features: same constructor and item contract as ``data.mmimdb.MMIMDb`` (data/mmimdb.py:14-207) -- one HDF5 file per split
    with ``vgg_features`` [N, 4096], ``features`` [N, 300], multi-hot ``genres`` [N, 23] and ``imdb_ids``.  The reference indexes the open
    file per item; here the four datasets are read once into host arrays.  Batches feed ``mml_b200.mmimdb.MMIMDb.train_step``."""

    VALID_SPLITS: List[str] = ["train", "val", "test"]
    NUM_CLASSES: int = 23
    MODS = ("image", "text")
    AVAILABLE_MODALITIES: Dict[str, Any] = {"image": "image", "text": "text"}
    # the reference's default (data/mmimdb.py:73-77)
    DEFAULT_PATTERNS = {"it": {"image": 1.0, "text": 1.0}, "i": {"image": 1.0, "text": 0.0}, "t": {"image": 0.0, "text": 1.0}}

    def __init__(self, data_fp: Union[str, Path], split: str, target_modality: Any = "multimodal", *,
                 missing_patterns: Optional[Mapping[str, Mapping[Any, float]]] = None, selected_patterns: Optional[Sequence[str]] = None,
                 image_key: str = "vgg_features", text_key: str = "features", labels_key: str = "genres", imdb_ids_key: str = "imdb_ids",
                 split_indices: Optional[Sequence[int]] = None, _id: int = 1, masks=None, generator: Optional[torch.Generator] = None,
                 pin: Optional[bool] = None, mask_seed: Optional[int] = None) -> None:
        self._configure(split, target_modality, missing_patterns, selected_patterns, _id)
        data_fp = Path(data_fp)
        if not data_fp.exists():
            raise FileNotFoundError(f"Dataset file not found: {data_fp}")
        f = _open_h5(data_fp)
        try:
            keys = list(f.keys())
            assert imdb_ids_key in keys, f"IMDb IDs key {imdb_ids_key} not found in the dataset"
            assert image_key in keys, f"Image key {image_key} not found in the dataset"
            assert text_key in keys, f"Text key {text_key} not found in the dataset"
            assert labels_key in keys, f"Labels key {labels_key} not found in the dataset"
            ids = [x.decode("utf-8") if isinstance(x, bytes) else str(x) for x in np.asarray(f[imdb_ids_key][...]).tolist()]
            self._store(np.asarray(f[labels_key][...]), np.asarray(f[image_key][...]), np.asarray(f[text_key][...]), ids, masks, generator, pin, mask_seed)
        finally:
            close = getattr(f, "close", None)
            if close is not None:
                close()

    @classmethod
    def from_arrays(cls, labels, image, text, imdb_ids: Optional[Sequence[str]] = None, split: str = "train", target_modality: Any = "multimodal", *,
                    missing_patterns=None, selected_patterns=None, _id: int = 1, masks=None, generator: Optional[torch.Generator] = None,
                    pin: Optional[bool] = None, mask_seed: Optional[int] = None) -> "MMIMDb":
        """The same dataset over arrays in memory: ``labels`` [N, 23], ``image`` [N, 4096], ``text`` [N, 300] (any float / integer dtype)."""
        self = cls.__new__(cls)
        self._configure(split, target_modality, missing_patterns, selected_patterns, _id)
        self._store(labels, image, text, imdb_ids, masks, generator, pin, mask_seed)
        return self

    def _store(self, labels, image, text, ids, masks, generator, pin, mask_seed=None) -> None:
        lab = torch.as_tensor(np.asarray(labels)).float().contiguous()  # ``torch.as_tensor(...).float()`` per item in the reference (:125-155)
        self._finish(lab.shape[0], masks, generator, pin, mask_seed)
        img, txt = torch.as_tensor(np.asarray(image)).float().contiguous(), torch.as_tensor(np.asarray(text)).float().contiguous()
        if img.shape[0] != self.num_samples or txt.shape[0] != self.num_samples:
            raise ValueError(f"labels / image / text disagree on the number of samples: {lab.shape[0]} / {img.shape[0]} / {txt.shape[0]}")
        self.label = lab
        self.data = {"image": img, "text": txt}
        self.imdb_ids = list(ids) if ids is not None else [str(i) for i in range(self.num_samples)]

    def _load_id(self, idx: int) -> str:
        return self.imdb_ids[idx]

    def __getitem__(self, idx: int) -> Dict[Any, Any]:
        pattern, i = self._item_head(idx)
        sample: Dict[Any, Any] = {"label": self.label[i].clone(), "pattern_name": pattern, "missing_mask": {}, "sample_idx": i}
        for m in self.MODS:
            sample[f"{m}_missing_index"] = self.masks[pattern][m][i]
        for m in self.MODS:
            if self._loads(m):
                self._masked(sample, self.keys[m], m, self.data[m][i].clone())
        return sample

    def batches(self, batch_size: int, shuffle: Optional[bool] = None, drop_last: bool = False, pattern: Optional[str] = None,
                rotate: int = 4, generator: Optional[torch.Generator] = None, rank: int = 0, world: int = 1) -> Iterator[Dict[Any, Any]]:
        """``label`` fp32 [B, 23], ``pattern_name``, ``sample_idx``, ``<mod>_original`` fp32 [B, D] + ``<mod>_missing_index`` fp32 [B] for the
        loaded modalities; order and staging as in ``AVMNIST.batches``."""
        for buf, r, names, msk in self._epoch(batch_size, shuffle, drop_last, pattern, generator, rotate, rank, world):
            out: Dict[Any, Any] = {"pattern_name": names}
            out["label"] = self._gather(buf, "label", self.label, r)
            out["sample_idx"] = self._gather(buf, "sample_idx", None, r)
            for j, m in enumerate(self.MODS):
                if self._loads(m):
                    out[f"{m}_original"] = self._gather(buf, m, self.data[m], r)
                    out[f"{m}_missing_index"] = self._gather(buf, m + "_mask", None, msk[:, j])
            yield out


class PatternSpecificDataset(Dataset):
    """The samples of one pattern of a validation / test split (data/pattern.py:6-19)."""

    def __init__(self, parent_dataset: _PatternDataset, pattern: str):
        self.parent, self.pattern = parent_dataset, pattern
        self.sample_indices = parent_dataset.pattern_indices[pattern]

    def __len__(self) -> int:
        return len(self.sample_indices)

    def __getitem__(self, idx: int) -> Dict[Any, Any]:
        if not 0 <= idx < len(self.sample_indices):
            raise IndexError(idx)
        return self.parent[idx + self.parent.selected_patterns.index(self.pattern) * self.parent.num_samples]


AVMNISTDataset = AVMNIST
