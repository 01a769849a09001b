"""Generate tests/golden/avmnist_loader.npz and mosi_loader.npz from the UNMODIFIED reference dataset classes.  Run in the build container only.

    python oracle/make_golden_loader.py

TEST INFRASTRUCTURE ONLY.  Writes a tiny AVMNIST-shaped dataset (CSV + torch.save-d items) into a scratch directory, runs the
reference's ``data.avmnist.AVMNIST`` (MML_Suite/data/avmnist.py:21-277, imported through oracle/ref_import.py) over it and stores the
raw inputs together with everything the class yields: items of a validation split (all three patterns), items of a training split
(``random.choice`` of the pattern under ``random.seed``), a monomodal test split, ``collate_fn`` batches, ``get_pattern_batches``
and a ``split_indices`` subset.  ``tests/test_datasets_cpu.py`` rebuilds the files from the stored inputs and requires
``mml_b200.datasets.AVMNIST`` to reproduce every stored output bit for bit.

Two things the reference takes from packages that are not in this image are pinned down explicitly and stored with the fixture:
the colormap (``matplotlib.cm.gist_earth`` -> a 256-colour table; matplotlib indexes the table directly for uint8 input) and the mask
draw (``modalities.create_missing_mask`` -> a seeded Bernoulli matrix; mask SAMPLING parity is unpinned by the reference).
"""
from __future__ import annotations

import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference  # noqa: E402

GOLDEN_DIR = os.environ.get("MML_GOLDEN_DIR") or os.path.join(os.path.dirname(HERE), "tests", "golden")  # the override is for the regeneration test
N, AUDIO_HW, IMAGE_HW = 5, (6, 9), (12, 10)


def raw_inputs():
    g = torch.Generator().manual_seed(2718)
    audio = torch.rand(N, *AUDIO_HW, generator=g) * torch.tensor([1e7, 1.0, 3.6e-9, 250.0, 1e-3]).reshape(N, 1, 1)  # the shipped spectrograms span 3.6e-9 .. 1.2e7
    audio[1, 0, 0], audio[2, 1, 1] = -0.0, float("inf")
    image = torch.randint(0, 256, (N, *IMAGE_HW), generator=g, dtype=torch.uint8)
    image[0, 0, :4] = torch.tensor([0, 1, 254, 255], dtype=torch.uint8)
    labels = np.array([3, 0, 9, 1, 7], dtype=np.int64)
    rng = np.random.default_rng(11)
    x = np.linspace(0.0, 1.0, 256)
    table = np.stack([np.clip(np.interp(x, [0, .3, .7, 1], [0, .2, .9, 1]) + rng.normal(0, .02, 256), 0, 1),
                      np.clip(np.interp(x, [0, .5, 1], [0, .8, .95]) + rng.normal(0, .02, 256), 0, 1),
                      np.clip(np.interp(x, [0, .2, 1], [.4, .3, 1]) + rng.normal(0, .02, 256), 0, 1), np.ones(256)], axis=1)
    return audio, image.numpy(), labels, table


def write_files(root: str, audio: torch.Tensor, image: np.ndarray, labels: np.ndarray) -> str:
    import pandas as pd

    rows = []
    for i in range(len(labels)):
        pa, pi = os.path.join(root, f"a{i}.pt"), os.path.join(root, f"i{i}.pt")
        torch.save(audio[i].clone(), pa)
        torch.save(image[i].copy(), pi)
        rows.append({"audio": pa, "image": pi, "label": int(labels[i])})
    csv = os.path.join(root, "data.csv")
    pd.DataFrame(rows).to_csv(csv, index=False)
    return csv


def items_to_arrays(prefix: str, items, M, out: dict) -> None:
    """Stack what the reference's ``__getitem__`` returned; absent modality entries are simply not stored."""
    out[f"{prefix}_labels"] = np.array([int(it["labels"]) for it in items], dtype=np.int64)
    out[f"{prefix}_sample_idx"] = np.array([int(it["sample_idx"]) for it in items], dtype=np.int64)
    out[f"{prefix}_pattern"] = np.array([it["pattern_name"] for it in items])
    out[f"{prefix}_keys"] = np.array([str(k) for k in items[0].keys()])
    for mod, m in (("audio", M.AUDIO), ("image", M.IMAGE)):
        out[f"{prefix}_{mod}_missing_index"] = np.array([float(it[f"{mod}_missing_index"]) for it in items], dtype=np.float32)
        if m in items[0]:
            out[f"{prefix}_{mod}"] = torch.stack([it[m] for it in items]).view(torch.int32).numpy()  # bit patterns (-0.0, inf * 0 = nan)
            out[f"{prefix}_{mod}_original"] = torch.stack([it[f"{mod}_original"] for it in items]).view(torch.int32).numpy()
            out[f"{prefix}_{mod}_reverse"] = torch.stack([it[f"{mod}_reverse"] for it in items]).view(torch.int32).numpy()


def main() -> None:
    ns = import_reference()
    M = ns.Modality
    import data.avmnist as RD
    import data.base_dataset as RB

    audio, image, labels, table = raw_inputs()
    RD.cm.gist_earth = lambda a: table[np.asarray(a)]  # Colormap.__call__ on integer input: table lookup
    drawn = []

    def create_missing_mask(n_modalities, batch_size, missing_rates):
        g = torch.Generator().manual_seed(1000 + len(drawn))
        keep = 1.0 - torch.tensor(list(missing_rates), dtype=torch.float32)
        m = torch.bernoulli(keep.expand(batch_size, n_modalities).contiguous(), generator=g)
        drawn.append(m)
        return m

    RB.create_missing_mask = create_missing_mask
    out = {"audio": audio.view(torch.int32).numpy(), "image": image, "labels": labels, "table": table}

    def masks_of(ds, prefix):
        for pat, tab in ds.masks.items():
            for m, v in tab.items():
                out[f"{prefix}_masks_{pat}_{m}"] = v.numpy().astype(np.float32)

    with tempfile.TemporaryDirectory() as root:
        csv = write_files(root, audio, image, labels)

        # (1) validation split, default patterns: idx -> (pattern idx // n, sample idx % n)
        ds = RD.AVMNIST(csv, "valid")
        assert len(ds) == 3 * N and ds.selected_patterns == ["a", "ai", "i"]
        masks_of(ds, "valid")
        items = [ds[i] for i in range(len(ds))]
        items_to_arrays("valid", items, M, out)
        for k, lo in enumerate((0, 7)):
            c = ds.collate_fn(items[lo:lo + 7])
            assert c["missing_masks"] == {}
            out[f"valid_collate{k}_labels"] = c["labels"].numpy()
            out[f"valid_collate{k}_pattern"] = np.array(c["pattern_name"])
            out[f"valid_collate{k}_audio"] = c[M.AUDIO].view(torch.int32).numpy()
            out[f"valid_collate{k}_image"] = c[M.IMAGE].view(torch.int32).numpy()
            out[f"valid_collate{k}_keys"] = np.array([str(x) for x in c.keys()])
        for pat, loader in ds.get_pattern_batches(2).items():
            bs = list(loader)
            out[f"valid_pb_{pat}_sizes"] = np.array([len(b["labels"]) for b in bs])
            out[f"valid_pb_{pat}_labels"] = torch.cat([b["labels"] for b in bs]).numpy()
            out[f"valid_pb_{pat}_audio"] = torch.cat([b[M.AUDIO] for b in bs]).view(torch.int32).numpy()
            out[f"valid_pb_{pat}_image"] = torch.cat([b[M.IMAGE] for b in bs]).view(torch.int32).numpy()
            out[f"valid_pb_{pat}_pattern"] = np.array(sum((b["pattern_name"] for b in bs), []))

        # (2) training split, audio present with P = 0.6, two selected patterns, pattern per item from Python's random
        mp = {"ai": {M.AUDIO: 0.6, M.IMAGE: 1.0}, "i": {M.AUDIO: 0.0, M.IMAGE: 1.0}}
        ds = RD.AVMNIST(csv, "train", missing_patterns=mp, selected_patterns=["ai", "i"])
        assert len(ds) == N
        masks_of(ds, "train")
        random.seed(3)
        items_to_arrays("train", [ds[i] for i in (4, 0, 2, 2, 1, 3)], M, out)

        # (3) monomodal test split: only the target modality is loaded, len = n * patterns
        ds = RD.AVMNIST(csv, "test", M.AUDIO, selected_patterns=["ai"])
        masks_of(ds, "testa")
        items = [ds[i] for i in range(len(ds))]
        assert M.IMAGE not in items[0] and "image_original" not in items[0]
        items_to_arrays("testa", items, M, out)
        c = ds.collate_fn(items)
        out["testa_collate_audio"] = c[M.AUDIO].view(torch.int32).numpy()
        out["testa_collate_keys"] = np.array([str(x) for x in c.keys()])

        # (4) split_indices subset
        ds = RD.AVMNIST(csv, "valid", selected_patterns=["i"], split_indices=[4, 1, 2])
        masks_of(ds, "sub")
        items_to_arrays("sub", [ds[i] for i in range(len(ds))], M, out)

    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, "avmnist_loader.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays,", len(drawn), "mask draws")
    mosi_case(ns, RB)
    mmimdb_case(ns, RB)


MOSI_N, MOSI_T, MOSI_F = 4, 6, {"audio": 3, "vision": 4, "text": 8}


def mosi_raw():
    """{split: {...}} in the pickle layout data/mosi.py:118-158 reads: zero-padded [N, T, F] float arrays, labels, lengths."""
    rng = np.random.default_rng(99)
    raw = {}
    for split, n in (("train", MOSI_N), ("valid", MOSI_N - 1), ("test", 2)):
        d = {}
        lengths = rng.integers(2, MOSI_T + 1, size=n)
        for k, f in MOSI_F.items():
            x = rng.normal(size=(n, MOSI_T, f)).astype(np.float32)
            for i, ln in enumerate(lengths):
                x[i, ln:] = 0.0
            d[k] = x
        d["classification_labels"] = rng.integers(0, 3, size=n)
        d["regression_labels"] = rng.normal(size=n).astype(np.float32)
        d["audio_lengths"], d["vision_lengths"] = lengths.copy(), lengths.copy()
        raw[split] = d
    return raw


def mosi_items(prefix: str, items, M, out: dict) -> None:
    out[f"{prefix}_label"] = torch.stack([it["label"] for it in items]).numpy()
    out[f"{prefix}_sample_idx"] = np.array([int(it["sample_idx"]) for it in items], dtype=np.int64)
    out[f"{prefix}_pattern"] = np.array([it["pattern_name"] for it in items])
    out[f"{prefix}_keys"] = np.array([str(k) for k in items[0].keys()])
    if "audio_length" in items[0]:
        out[f"{prefix}_audio_length"] = np.array([float(it["audio_length"]) for it in items], dtype=np.float32)
        out[f"{prefix}_video_length"] = np.array([float(it["video_length"]) for it in items], dtype=np.float32)
    for mod, m in (("audio", M.AUDIO), ("video", M.VIDEO), ("text", M.TEXT)):
        out[f"{prefix}_{mod}_missing_index"] = np.array([float(it[f"{mod}_missing_index"]) for it in items], dtype=np.float32)
        if m in items[0]:
            for suffix, key in (("", m), ("_original", f"{mod}_original"), ("_reverse", f"{mod}_reverse")):
                out[f"{prefix}_{mod}{suffix}"] = torch.stack([it[key] for it in items]).contiguous().view(torch.int32).numpy()


def mosi_case(ns, RB) -> None:
    """tests/golden/mosi_loader.npz from the unmodified ``data.mosi.MOSI`` (MML_Suite/data/mosi.py:17-301)."""
    import pickle

    import data.mosi as RM

    M = ns.Modality
    raw = mosi_raw()
    drawn = []

    def create_missing_mask(n_modalities, batch_size, missing_rates):
        g = torch.Generator().manual_seed(2000 + len(drawn))
        keep = 1.0 - torch.tensor(list(missing_rates), dtype=torch.float32)
        m = torch.bernoulli(keep.expand(batch_size, n_modalities).contiguous(), generator=g)
        drawn.append(m)
        return m

    RB.create_missing_mask = create_missing_mask
    out = {}
    for split, d in raw.items():
        for k, v in d.items():
            out[f"raw_{split}_{k}"] = np.asarray(v)

    def masks_of(ds, prefix):
        for pat, tab in ds.masks.items():
            for m, v in tab.items():
                out[f"{prefix}_masks_{pat}_{m}"] = v.numpy().astype(np.float32)

    with tempfile.TemporaryDirectory() as root:
        fp = os.path.join(root, "mosi.pkl")
        with open(fp, "wb") as f:
            pickle.dump(raw, f)
        # (1) validation split, all seven patterns, unaligned (lengths in the items)
        ds = RM.MOSI(fp, "valid")
        assert len(ds) == 7 * (MOSI_N - 1) and ds.selected_patterns == ["a", "at", "atv", "av", "t", "tv", "v"]
        masks_of(ds, "valid")
        mosi_items("valid", [ds[i] for i in range(len(ds))], M, out)
        # (2) training split: audio present with P = 0.8 in the full pattern (the MOSI YAMLs' audio rate 0.2), pattern per item from ``random``
        mp = {"atv": {M.AUDIO: 0.8, M.TEXT: 1.0, M.VIDEO: 1.0}, "t": {M.AUDIO: 0.0, M.TEXT: 1.0, M.VIDEO: 0.0}}
        ds = RM.MOSI(fp, "train", missing_patterns=mp, selected_patterns=["atv", "t"], aligned=True, length=MOSI_T)
        masks_of(ds, "train")
        random.seed(5)
        mosi_items("train", [ds[i] for i in (3, 0, 1, 1, 2)], M, out)
        # (3) monomodal test split with regression labels
        ds = RM.MOSI(fp, "test", M.TEXT, selected_patterns=["atv"], labels_key="regression_labels")
        masks_of(ds, "testt")
        mosi_items("testt", [ds[i] for i in range(len(ds))], M, out)
    path = os.path.join(GOLDEN_DIR, "mosi_loader.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays,", len(drawn), "mask draws")




class FakeH5(dict):
    """Dict-backed stand-in for ``h5py.File`` (h5py is not installed here): ``keys()``, ``[key]`` -> array with ``[idx]`` / ``[...]`` / ``len``."""

    def close(self):
        pass


def mmimdb_raw(n: int = 5):
    rng = np.random.default_rng(123)
    return {"imdb_ids": np.array([f"tt{1000 + i:07d}".encode() for i in range(n)]), "vgg_features": np.abs(rng.normal(size=(n, 12))).astype(np.float32),
            "features": rng.normal(0, 0.3, size=(n, 7)).astype(np.float64), "genres": (rng.random((n, 23)) < 0.2).astype(np.int64)}


def mmimdb_case(ns, RB) -> None:
    """tests/golden/mmimdb_loader.npz from the unmodified ``data.mmimdb.MMIMDb`` (MML_Suite/data/mmimdb.py:14-207) over a dict-backed HDF5."""
    import data.mmimdb as RMM

    M = ns.Modality
    raw = mmimdb_raw()
    drawn = []

    def create_missing_mask(n_modalities, batch_size, missing_rates):
        g = torch.Generator().manual_seed(3000 + len(drawn))
        keep = 1.0 - torch.tensor(list(missing_rates), dtype=torch.float32)
        m = torch.bernoulli(keep.expand(batch_size, n_modalities).contiguous(), generator=g)
        drawn.append(m)
        return m

    RB.create_missing_mask = create_missing_mask
    RMM.h5.File = lambda path, mode="r": FakeH5(raw)
    out = {f"raw_{k}": v for k, v in raw.items()}

    def store(prefix, ds, idxs):
        for pat, tab in ds.masks.items():
            for m, v in tab.items():
                out[f"{prefix}_masks_{pat}_{m}"] = v.numpy().astype(np.float32)
        items = [ds[i] for i in idxs]
        out[f"{prefix}_label"] = torch.stack([it["label"] for it in items]).numpy()
        out[f"{prefix}_sample_idx"] = np.array([int(it["sample_idx"]) for it in items], dtype=np.int64)
        out[f"{prefix}_pattern"] = np.array([it["pattern_name"] for it in items])
        out[f"{prefix}_keys"] = np.array([str(k) for k in items[0].keys()])
        out[f"{prefix}_ids"] = np.array([ds._load_id(int(it["sample_idx"])) for it in items])
        for mod, m in (("image", M.IMAGE), ("text", M.TEXT)):
            out[f"{prefix}_{mod}_missing_index"] = np.array([float(it[f"{mod}_missing_index"]) for it in items], dtype=np.float32)
            if m in items[0]:
                for suffix, key in (("", m), ("_original", f"{mod}_original"), ("_reverse", f"{mod}_reverse")):
                    out[f"{prefix}_{mod}{suffix}"] = torch.stack([it[key] for it in items]).contiguous().view(torch.int32).numpy()

    with tempfile.TemporaryDirectory() as root:
        fp = os.path.join(root, "val.h5")
        open(fp, "wb").close()  # the class only checks that the path exists before h5py opens it
        ds = RMM.MMIMDb(fp, "val")
        assert len(ds) == 15 and ds.selected_patterns == ["i", "it", "t"]
        store("val", ds, range(len(ds)))
        mp = {"it": {M.IMAGE: 0.3, M.TEXT: 0.7}, "t": {M.IMAGE: 0.0, M.TEXT: 1.0}}  # missing_exp/baseline_30_70.yaml: text 0.3, image 0.7 missing
        ds = RMM.MMIMDb(fp, "train", missing_patterns=mp, selected_patterns=["it", "t"])
        random.seed(8)
        store("train", ds, (1, 4, 4, 0, 2))
        ds = RMM.MMIMDb(fp, "test", M.TEXT, selected_patterns=["it"])
        store("testt", ds, range(len(ds)))
    path = os.path.join(GOLDEN_DIR, "mmimdb_loader.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays,", len(drawn), "mask draws")


if __name__ == "__main__":
    main()
