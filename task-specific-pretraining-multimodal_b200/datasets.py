"""The AVMNIST dataset of the reference, laid out for a step that consumes 100 k samples/s.

Reference: ``MML_Suite/data/avmnist.py:21-277`` (class ``AVMNIST``) on top of ``MultimodalBaseDataset``
(``data/base_dataset.py:16-154``) and ``PatternSpecificDataset`` (``data/pattern.py:6-19``): a CSV of
``audio`` / ``image`` / ``label`` columns whose cells are paths of ``torch.save``-d items; ``__getitem__`` loads both files
(lru-cached), runs the image through ``uint8 -> cm.gist_earth -> PIL "L" -> PILToTensor -> ToDtype(float32, scale=True)``,
looks the sample's masks up and multiplies (``get_samples``, base_dataset.py:61-74).  One item per Python call on a DataLoader
worker tops out four orders of magnitude below what the fused step eats (2.4 ms per 256 samples).

Here every file is read ONCE at construction into three contiguous host arrays (pinned when CUDA is there) -- audio fp32
``[N, H, W]``, image uint8 ``[N, h, w]`` (one byte per pixel: the whole image chain is a 256-entry table, ``data.luma_lut``),
labels int64 ``[N]`` -- plus the mask table ``[pattern][modality][len(self)]``.  Two views of that storage:

* ``__getitem__`` / ``collate_fn`` / ``get_pattern_batches``: the reference's item and batch dictionaries, bit for bit (same keys,
  shapes, dtypes, ordering of the validation / test splits, ``random.choice`` of the training pattern).  Pinned against the
  unmodified reference class by ``oracle/make_golden_loader.py`` (fixture ``tests/golden/avmnist_loader.npz``).
* ``batches()``: whole batches by vectorised row gathers into rotating pinned staging buffers, in the form the fused step takes --
  ``<mod>_original`` + ``<mod>_missing_index`` (the multiply runs in the first kernel of each encoder) with the image still uint8 --
  and ``fused_loader()`` = ``DevicePrefetcher(batches(), luts={"image_original": lut})``: 1 byte per image pixel over PCIe, table
  lookup on the copy stream (``mml_stage_u8_lut_f32``).

Mask *sampling* follows ``data.draw_missing_masks`` (independent Bernoulli(P(present)); the reference's generator lives in the
un-vendored ``modalities`` package) or is handed in (``masks=``: e.g. ``DeviceMaskTable(...).masks`` copied back, or
the table of a reference run).
"""
from __future__ import annotations

import random
from itertools import combinations
from pathlib import Path
from typing import Any, Callable, Dict, Iterator, List, Mapping, Optional, Sequence, Union

import numpy as np
import torch
from torch.utils.data import Dataset

from .data import draw_missing_masks, luma_lut

NAMES = ("audio", "image")


def _modality_keys() -> Dict[str, Any]:
    """``modalities.Modality`` members when the package is importable (the reference's batch keys), else the lower-case names
    (``str(Modality.AUDIO) == "audio"``, so the ``<mod>_original`` / ``<mod>_missing_index`` keys are the same either way)."""
    try:
        from modalities import Modality  # type: ignore

        return {"audio": Modality.AUDIO, "image": Modality.IMAGE, "multimodal": Modality.MULTIMODAL}
    except Exception:
        return {"audio": "audio", "image": "image", "multimodal": "multimodal"}


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


class AVMNIST(Dataset):
    """Same constructor and item / batch contract as ``data.avmnist.AVMNIST`` (data/avmnist.py:45-59), plus ``cmap`` (colour table),
    ``masks`` (explicit mask table), ``generator`` (mask draw / batch shuffling) and ``pin`` (default: pinned iff CUDA is available)."""

    NUM_CLASSES: int = 10
    VALID_SPLITS: List[str] = ["train", "valid", "test"]
    AVAILABLE_MODALITIES: Dict[str, Any] = {"audio": "audio", "image": "image"}

    @staticmethod
    def get_full_modality() -> str:
        return "".join(sorted(k[0] for k in AVMNIST.AVAILABLE_MODALITIES))

    @classmethod
    def get_all_possible_patterns(cls) -> List[str]:
        mods = list(cls.AVAILABLE_MODALITIES.keys())
        return sorted("".join(m[0] for m in sorted(c)) for r in range(1, len(mods) + 1) for c in combinations(mods, r))

    def validate_patterns(self, patterns: Sequence[str]) -> List[str]:
        bad = set(patterns) - set(self.get_all_possible_patterns())
        if bad:
            raise ValueError(f"Invalid patterns: {bad}\nValid patterns are: {self.get_all_possible_patterns()}")
        return list(patterns)

    def __init__(self, data_fp: Union[str, Path], split: str, target_modality: Any = "multimodal", *,
                 missing_patterns: Optional[Mapping[str, Mapping[Any, float]]] = None, selected_patterns: Optional[Sequence[str]] = None,
                 audio_column: str = "audio", image_column: str = "image", labels_column: str = "label",
                 split_indices: Optional[Sequence[int]] = None, _id: int = 1, cmap: Any = None,
                 masks: Optional[Mapping[str, Mapping[Any, torch.Tensor]]] = None, generator: Optional[torch.Generator] = None,
                 pin: Optional[bool] = None) -> None:
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
        tm = self._target
        labels = torch.from_numpy(np.array(self.data[labels_column].to_numpy(), dtype=np.int64))
        audio = self._read_audio(self.data[audio_column]) if tm in ("audio", "multimodal") else None
        image = self._read_images(self.data[image_column]) if tm in ("image", "multimodal") else None
        self._store(labels, audio, image, cmap, masks, generator, pin)

    @classmethod
    def from_arrays(cls, labels, audio=None, image_u8=None, split: str = "train", target_modality: Any = "multimodal", *,
                    missing_patterns=None, selected_patterns=None, _id: int = 1, cmap: Any = None, masks=None,
                    generator: Optional[torch.Generator] = None, pin: Optional[bool] = None) -> "AVMNIST":
        """The same dataset over arrays that are already in memory (``audio`` fp32 [N, H, W], ``image_u8`` uint8 [N, h, w], ``labels``
        [N]) instead of a CSV of per-item files; ``None`` for a modality that the target does not load."""
        self = cls.__new__(cls)
        self._configure(split, target_modality, missing_patterns, selected_patterns, _id)
        self.data_fp = self.data = None
        labels = torch.as_tensor(labels, dtype=torch.long).reshape(-1).clone()
        tm = self._target
        a = i = None
        if tm in ("audio", "multimodal"):
            a = torch.as_tensor(audio, dtype=torch.float32).contiguous()
            if a.dim() != 3 or a.shape[0] != labels.numel():
                raise ValueError(f"audio must be [N, H, W] with N = {labels.numel()}, got {tuple(a.shape)}")
        if tm in ("image", "multimodal"):
            i = torch.as_tensor(image_u8)
            if i.dtype != torch.uint8 or i.dim() != 3 or i.shape[0] != labels.numel():
                raise TypeError(f"image_u8 must be uint8 [N, h, w] with N = {labels.numel()}, got {i.dtype} {tuple(i.shape)}")
            i = i.contiguous()
        self._store(labels, a, i, cmap, masks, generator, pin)
        return self

    def _configure(self, split, target_modality, missing_patterns, selected_patterns, _id) -> None:
        self.split = str(split).lower()
        assert split in self.VALID_SPLITS, f"Invalid split provided, must be one of {self.VALID_SPLITS}"
        assert isinstance(_id, int), "ID must be an integer."
        self._id = _id
        self.keys = _modality_keys()
        self.AVAILABLE_MODALITIES = {n: self.keys[n] for n in NAMES}
        # pattern -> {modality name: P(present)}; the default is the reference's (data/avmnist.py:73-77)
        mp = missing_patterns or {"ai": {"audio": 1.0, "image": 1.0}, "a": {"audio": 1.0, "image": 0.0}, "i": {"audio": 0.0, "image": 1.0}}
        self.missing_patterns = {pat: {_name(m): float(p) for m, p in probs.items()} for pat, probs in mp.items()}
        self.selected_patterns = self.validate_patterns(selected_patterns) if selected_patterns is not None else self.get_all_possible_patterns()
        for pat in self.selected_patterns:
            if pat not in self.missing_patterns:
                raise ValueError(f"selected pattern {pat!r} has no entry in missing_patterns {list(self.missing_patterns)}")
        self.current_pattern = None
        tm = _name(target_modality)
        assert tm in ("audio", "image", "multimodal"), "Invalid modality provided, must be one of [audio, image, multimodal]"
        self.target_modality = self.keys[tm]
        self._target = tm

    def _store(self, labels, audio, image_u8, cmap, masks, generator, pin) -> None:
        self.num_samples = int(labels.numel())
        self.pattern_indices = {pattern: list(range(self.num_samples)) for pattern in self.selected_patterns}
        pin = torch.cuda.is_available() if pin is None else bool(pin)
        self._pin = pin
        self.labels = _maybe_pin(labels, pin)
        self.audio = _maybe_pin(audio, pin) if audio is not None else None
        self.image_u8 = self.lut = None
        if image_u8 is not None:
            self.lut = luma_lut(_colour_table(cmap))  # fp32 [256]: the reference's image chain as a function of the pixel value
            self.image_u8 = _maybe_pin(image_u8, pin)
        if masks is not None:
            self.masks = {pat: {_name(m): torch.as_tensor(v, dtype=torch.float32).reshape(-1) for m, v in tab.items()} for pat, tab in masks.items()}
            for pat in self.missing_patterns:
                for m in NAMES:
                    if pat not in self.masks or m not in self.masks[pat] or self.masks[pat][m].numel() < self.num_samples:
                        raise ValueError(f"masks[{pat!r}][{m!r}] must hold at least {self.num_samples} entries")
        else:
            # one draw per (pattern, modality, dataset index) at construction, length len(self) like base_dataset.py:46-59
            self.masks = draw_missing_masks(self.missing_patterns, len(self), generator)
        self.generator = generator
        # [pattern][modality][sample] as one tensor for the vectorised batch path
        self._pat_index = {pat: i for i, pat in enumerate(self.missing_patterns)}
        self._mask_table = torch.stack([torch.stack([self.masks[pat][m][: self.num_samples] for m in NAMES]) for pat in self.missing_patterns])

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
    def __len__(self) -> int:
        return self.num_samples if self.split == "train" else self.num_samples * len(self.selected_patterns)

    def _get_pattern_and_sample_idx(self, idx: int):
        if self.split == "train" or self.split == "trn":
            return random.choice(self.selected_patterns), idx  # base_dataset.py:87-89: Python's global ``random``
        return self.selected_patterns[idx // self.num_samples], idx % self.num_samples

    def image_float(self, rows) -> torch.Tensor:
        """fp32 [n, 1, h, w] images of ``rows`` = ``_load_image`` of the reference for each of them (table lookup on the host)."""
        return self.lut[self.image_u8[rows].long()].unsqueeze(-3)

    def __getitem__(self, idx: int) -> Dict[Any, Any]:
        pattern, i = self._get_pattern_and_sample_idx(int(idx))
        if not 0 <= i < self.num_samples:
            raise IndexError(idx)
        self.current_pattern = pattern
        sample: Dict[Any, Any] = {"labels": self.labels[i].clone(), "pattern_name": pattern, "missing_mask": {}, "sample_idx": i}
        for m in NAMES:
            sample[f"{m}_missing_index"] = self.masks[pattern][m][i]
        for m in NAMES:
            if self._target in ("multimodal", m):
                original = self.audio[i].clone() if m == "audio" else self.image_float(i)
                mask = sample[f"{m}_missing_index"]
                sample[f"{m}_original"] = original
                sample[self.keys[m]] = original * mask
                sample[f"{m}_reverse"] = original * -1 * (mask - 1)
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

    def get_split(self) -> str:
        return self.split

    def get_selected_patterns(self) -> List[str]:
        return self.selected_patterns

    def get_missing_patterns(self):
        return self.missing_patterns

    # ---- batch-granular path ------------------------------------------------------------------------------------------------
    def batches(self, batch_size: int, shuffle: Optional[bool] = None, drop_last: bool = False, pattern: Optional[str] = None,
                image_form: str = "u8", rotate: int = 4, generator: Optional[torch.Generator] = None) -> Iterator[Dict[Any, Any]]:
        """Whole batches in the fused step's input form: ``labels`` int64 [B], ``pattern_name`` list, ``sample_idx`` int64 [B],
        ``audio_original`` fp32 [B, H, W], ``image_original`` uint8 [B, 1, h, w] (``image_form="u8"``: expand on the device with
        ``DevicePrefetcher(luts={"image_original": ds.lut})``) or fp32 (``"f32"``: table lookup on the host), ``<mod>_missing_index`` fp32 [B].

        Order: the training split visits every sample once (shuffled unless ``shuffle=False``) with an independent uniformly drawn
        pattern per sample (the vectorised form of ``random.choice``, base_dataset.py:87-89; drawn from ``generator``, not from
        Python's ``random``); the other splits walk ``selected_patterns`` in order, all samples of one pattern after the other --
        dataset index order, data/base_dataset.py:90-93 -- or only ``pattern``.  A yielded batch's tensors live in one of ``rotate``
        pinned staging buffer sets and stay valid until ``rotate - 1`` further batches have been drawn (enough for a copy stream
        one batch ahead of the step)."""
        if image_form not in ("u8", "f32"):
            raise ValueError("image_form must be 'u8' or 'f32'")
        if batch_size < 1 or rotate < 1:
            raise ValueError("batch_size and rotate must be positive")
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
        stop = total - (total % batch_size) if drop_last else total
        for n, lo in enumerate(range(0, stop, batch_size)):
            r, p = rows[lo:lo + batch_size], pats[lo:lo + batch_size]
            buf = bufs[n % rotate]
            out: Dict[Any, Any] = {"pattern_name": [self.selected_patterns[k] for k in p.tolist()]}
            out["labels"] = self._gather(buf, "labels", self.labels, r)
            out["sample_idx"] = self._gather(buf, "sample_idx", None, r)
            m = self._mask_table[table_row[p], :, r]  # [B, n_modalities]
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

    def background_batches(self, batch_size: int, ahead: int = 2, **kwargs) -> Iterator[Dict[Any, Any]]:
        """``batches()`` produced by a worker thread, ``ahead`` batches in front of the consumer (the row gathers are ``index_select``
        calls that release the GIL, so they overlap the step's host code instead of adding 1.4-3.3 ms per 256-sample batch to it).
        ``rotate`` defaults to ``ahead + 3`` staging sets: ``ahead`` queued, one being filled, one with the consumer, one spare."""
        import queue
        import threading

        kwargs.setdefault("rotate", ahead + 3)
        if kwargs["rotate"] < ahead + 2:
            raise ValueError("rotate must be at least ahead + 2 (queued batches + the one being filled + the one in use)")
        q: "queue.Queue" = queue.Queue(maxsize=max(1, int(ahead)))
        stop = threading.Event()
        done = object()

        def work():
            try:
                for b in self.batches(batch_size, **kwargs):
                    while not stop.is_set():
                        try:
                            q.put(b, timeout=0.1)
                            break
                        except queue.Full:
                            continue
                    if stop.is_set():
                        return
                item: Any = done
            except BaseException as e:  # surfaced in the consumer
                item = e
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return
                except queue.Full:
                    continue

        t = threading.Thread(target=work, name="avmnist-batches", daemon=True)
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
        """``batches()`` (on a worker thread when ``ahead`` > 0) behind the device prefetcher: uint8 images cross PCIe as bytes and are
        expanded through the luminance table on the copy stream; the result feeds ``mml_b200.avmnist.AVMNIST.train_step /
        validation_step`` directly.  Staging sets: ``ahead`` + ``depth`` + 3 (queued + staged on the copy stream + filling / in use / spare)."""
        from .data import DevicePrefetcher

        kwargs.setdefault("rotate", ahead + depth + 3)
        luts = {"image_original": self.lut} if self.image_u8 is not None and kwargs.get("image_form", "u8") == "u8" else None
        it = self.background_batches(batch_size, ahead=ahead, **kwargs) if ahead > 0 else self.batches(batch_size, **kwargs)
        return DevicePrefetcher(it, device, depth=depth, luts=luts)


class PatternSpecificDataset(Dataset):
    """The samples of one pattern of a validation / test split (data/pattern.py:6-19)."""

    def __init__(self, parent_dataset: AVMNIST, pattern: str):
        self.parent, self.pattern = parent_dataset, pattern
        self.sample_indices = parent_dataset.pattern_indices[pattern]

    def __len__(self) -> int:
        return len(self.sample_indices)

    def __getitem__(self, idx: int) -> Dict[Any, Any]:
        if not 0 <= idx < len(self.sample_indices):
            raise IndexError(idx)
        return self.parent[idx + self.parent.selected_patterns.index(self.pattern) * self.parent.num_samples]


AVMNISTDataset = AVMNIST
