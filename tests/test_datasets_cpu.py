"""``mml_b200.datasets.AVMNIST`` against the reference's own dataset class ("the real AVMNIST loader").

``tests/golden/avmnist_loader.npz`` was written by ``oracle/make_golden_loader.py`` from the UNMODIFIED ``data.avmnist.AVMNIST``
(MML_Suite/data/avmnist.py:21-277) over a tiny CSV + ``torch.save`` dataset; the raw inputs are stored with it, the files are rebuilt
here, and every item / batch the reference produced must come back bit for bit (the comparison is on int32 views: -0.0, inf * 0 = nan).
The batch-granular path (``batches()``) has no reference counterpart; it is checked against the item path.
"""
import os
import random

import numpy as np
import pytest
import torch

from mml_b200.datasets import AVMNIST, PatternSpecificDataset

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "avmnist_loader.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def csv(gold, tmp_path_factory):
    import pandas as pd

    root = tmp_path_factory.mktemp("avmnist_files")
    audio = torch.from_numpy(gold["audio"]).view(torch.float32)
    rows = []
    for i, lab in enumerate(gold["labels"]):
        pa, pi = str(root / f"a{i}.pt"), str(root / f"i{i}.pt")
        torch.save(audio[i].clone(), pa)
        torch.save(gold["image"][i].copy(), pi)
        rows.append({"audio": pa, "image": pi, "label": int(lab)})
    path = str(root / "data.csv")
    pd.DataFrame(rows).to_csv(path, index=False)
    return path


def _masks(gold, prefix):
    out = {}
    for k in gold.files:
        if k.startswith(prefix + "_masks_"):
            pat, mod = k[len(prefix) + 7:].split("_")
            out.setdefault(pat, {})[mod] = torch.from_numpy(gold[k])
    return out


def _bits(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int32).numpy()


def _check_items(gold, prefix, items, ds):
    """Modality entries are keyed by ``modalities.Modality`` members when that package is importable, else by name: ``ds.keys``."""
    assert [str(k) for k in items[0].keys()] == list(gold[f"{prefix}_keys"])
    assert [int(it["labels"]) for it in items] == list(gold[f"{prefix}_labels"])
    assert [it["sample_idx"] for it in items] == list(gold[f"{prefix}_sample_idx"])
    assert [it["pattern_name"] for it in items] == list(gold[f"{prefix}_pattern"])
    for it in items:
        assert it["labels"].dtype == torch.long and it["labels"].dim() == 0 and it["missing_mask"] == {}
    for mod in ("audio", "image"):
        assert [float(it[f"{mod}_missing_index"]) for it in items] == list(gold[f"{prefix}_{mod}_missing_index"])
        if f"{prefix}_{mod}" not in gold.files:
            assert ds.keys[mod] not in items[0] and f"{mod}_original" not in items[0]
            continue
        for suffix, key in (("", ds.keys[mod]), ("_original", f"{mod}_original"), ("_reverse", f"{mod}_reverse")):
            got = torch.stack([it[key] for it in items])
            assert got.dtype == torch.float32
            want = gold[f"{prefix}_{mod}{suffix}"]
            assert got.shape == want.shape, (mod, suffix, got.shape, want.shape)
            assert np.array_equal(_bits(got), want), (prefix, mod, suffix)


def test_validation_items_collate_and_pattern_batches(gold, csv):
    ds = AVMNIST(csv, "valid", cmap=gold["table"], masks=_masks(gold, "valid"))
    assert len(ds) == 15 and ds.selected_patterns == ["a", "ai", "i"] and ds.get_full_modality() == "ai" and ds.num_samples == 5
    items = [ds[i] for i in range(len(ds))]
    _check_items(gold, "valid", items, ds)
    with pytest.raises(IndexError):
        ds[15]
    for k, lo in enumerate((0, 7)):
        c = ds.collate_fn(items[lo:lo + 7])
        assert [str(x) for x in c.keys()] == list(gold[f"valid_collate{k}_keys"]) and c["missing_masks"] == {}
        assert np.array_equal(c["labels"].numpy(), gold[f"valid_collate{k}_labels"])
        assert c["pattern_name"] == list(gold[f"valid_collate{k}_pattern"])
        assert np.array_equal(_bits(c[ds.keys["audio"]]), gold[f"valid_collate{k}_audio"])
        assert np.array_equal(_bits(c[ds.keys["image"]]), gold[f"valid_collate{k}_image"])
    loaders = ds.get_pattern_batches(2)
    assert list(loaders) == ["a", "ai", "i"]
    for pat, loader in loaders.items():
        assert isinstance(loader.dataset, PatternSpecificDataset) and len(loader.dataset) == 5
        bs = list(loader)
        assert [len(b["labels"]) for b in bs] == list(gold[f"valid_pb_{pat}_sizes"])
        assert np.array_equal(torch.cat([b["labels"] for b in bs]).numpy(), gold[f"valid_pb_{pat}_labels"])
        assert np.array_equal(_bits(torch.cat([b[ds.keys["audio"]] for b in bs])), gold[f"valid_pb_{pat}_audio"])
        assert np.array_equal(_bits(torch.cat([b[ds.keys["image"]] for b in bs])), gold[f"valid_pb_{pat}_image"])
        assert sum((b["pattern_name"] for b in bs), []) == list(gold[f"valid_pb_{pat}_pattern"])


def test_training_items_pick_the_pattern_like_the_reference(gold, csv):
    mp = {"ai": {"audio": 0.6, "image": 1.0}, "i": {"audio": 0.0, "image": 1.0}}
    ds = AVMNIST(csv, "train", missing_patterns=mp, selected_patterns=["ai", "i"], cmap=gold["table"], masks=_masks(gold, "train"))
    assert len(ds) == 5
    random.seed(3)  # base_dataset.py:87-89 draws from Python's global generator
    _check_items(gold, "train", [ds[i] for i in (4, 0, 2, 2, 1, 3)], ds)
    with pytest.raises(ValueError):
        ds.get_pattern_batches(2)


def test_monomodal_split_and_split_indices(gold, csv):
    ds = AVMNIST(csv, "test", "audio", selected_patterns=["ai"], masks=_masks(gold, "testa"))  # no colour table needed without images
    assert len(ds) == 5 and ds.image_u8 is None
    items = [ds[i] for i in range(len(ds))]
    _check_items(gold, "testa", items, ds)
    c = ds.collate_fn(items)
    assert [str(x) for x in c.keys()] == list(gold["testa_collate_keys"])
    assert np.array_equal(_bits(c[ds.keys["audio"]]), gold["testa_collate_audio"])
    sub = AVMNIST(csv, "valid", selected_patterns=["i"], split_indices=[4, 1, 2], cmap=gold["table"], masks=_masks(gold, "sub"))
    assert len(sub) == 3
    _check_items(gold, "sub", [sub[i] for i in range(3)], sub)


def test_constructor_errors(gold, csv, tmp_path):
    with pytest.raises(FileNotFoundError):
        AVMNIST(str(tmp_path / "nope.csv"), "train", cmap=gold["table"])
    with pytest.raises(AssertionError):
        AVMNIST(csv, "validation", cmap=gold["table"])
    with pytest.raises(ValueError, match="Invalid patterns"):
        AVMNIST(csv, "valid", selected_patterns=["x"], cmap=gold["table"])
    with pytest.raises(ValueError, match="Missing required columns"):
        AVMNIST(csv, "valid", labels_column="digit", cmap=gold["table"])
    with pytest.raises(ValueError, match="colour table"):
        AVMNIST(csv, "valid", cmap=np.zeros((16, 3)))
    bad = tmp_path / "f.pt"
    torch.save(np.zeros((3, 3), dtype=np.float32), str(bad))
    import pandas as pd

    df = pd.read_csv(csv)
    df.loc[1, "image"] = str(bad)
    p = str(tmp_path / "bad.csv")
    df.to_csv(p, index=False)
    with pytest.raises(TypeError, match="uint8"):
        AVMNIST(p, "valid", cmap=gold["table"])


@pytest.mark.parametrize("image_form", ["u8", "f32"])
def test_batches_equal_the_item_path(gold, csv, image_form):
    """Evaluation order = dataset index order; tensors = the items' ORIGINAL tensors + masks (the multiply runs on the device)."""
    ds = AVMNIST(csv, "valid", cmap=gold["table"], masks=_masks(gold, "valid"), pin=False)
    items = [ds[i] for i in range(len(ds))]
    got = []
    for b in ds.batches(4, image_form=image_form, rotate=1):
        got.append({k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.items()})
    assert [len(b["labels"]) for b in got] == [4, 4, 4, 3]
    assert sum((b["pattern_name"] for b in got), []) == [it["pattern_name"] for it in items]
    cat = {k: torch.cat([b[k] for b in got]) for k in got[0] if k != "pattern_name"}
    assert cat["labels"].tolist() == [int(it["labels"]) for it in items] and cat["labels"].dtype == torch.long
    assert cat["sample_idx"].tolist() == [it["sample_idx"] for it in items]
    assert np.array_equal(_bits(cat["audio_original"]), _bits(torch.stack([it["audio_original"] for it in items])))
    for mod in ("audio", "image"):
        assert cat[f"{mod}_missing_index"].dtype == torch.float32
        assert cat[f"{mod}_missing_index"].tolist() == [float(it[f"{mod}_missing_index"]) for it in items]
    want_img = torch.stack([it["image_original"] for it in items])
    if image_form == "u8":
        assert cat["image_original"].dtype == torch.uint8 and cat["image_original"].shape == want_img.shape
        assert torch.equal(ds.lut[cat["image_original"].long()], want_img)  # what mml_stage_u8_lut_f32 computes on the device
    else:
        assert torch.equal(cat["image_original"], want_img)
    # x * mask of the reference == ORIGINAL * missing_index of the batch path
    masked = cat["audio_original"] * cat["audio_missing_index"].reshape(-1, 1, 1)
    assert np.array_equal(_bits(masked), _bits(torch.stack([it[ds.keys["audio"]] for it in items])))
    one = list(ds.batches(8, pattern="ai", drop_last=True))
    assert len(one) == 0 and [b["pattern_name"] for b in ds.batches(5, pattern="ai")] == [["ai"] * 5]


def test_training_batches_visit_every_sample_once(gold, csv):
    mp = {"ai": {"audio": 0.6, "image": 1.0}, "i": {"audio": 0.0, "image": 1.0}}
    g = torch.Generator().manual_seed(5)
    ds = AVMNIST(csv, "train", missing_patterns=mp, selected_patterns=["ai", "i"], cmap=gold["table"], generator=g, pin=False)
    assert set(ds.masks) == {"ai", "i"} and all(v.shape == (5,) for t in ds.masks.values() for v in t.values())
    assert not ds.masks["i"]["audio"].any() and ds.masks["ai"]["image"].all()
    seen, n_batches = [], 0
    for b in ds.batches(2, drop_last=False):
        n_batches += 1
        for j, (i, pat) in enumerate(zip(b["sample_idx"].tolist(), b["pattern_name"])):
            seen.append(i)
            assert b["audio_missing_index"][j] == ds.masks[pat]["audio"][i] and b["image_missing_index"][j] == ds.masks[pat]["image"][i]
            assert torch.equal(b["audio_original"][j], ds.audio[i]) and torch.equal(b["image_original"][j, 0], ds.image_u8[i])
            assert int(b["labels"][j]) == int(gold["labels"][i])
    assert n_batches == 3 and sorted(seen) == [0, 1, 2, 3, 4]
    assert [len(b["labels"]) for b in ds.batches(2, drop_last=True)] == [2, 2]
    assert [b["sample_idx"].tolist() for b in ds.batches(5, shuffle=False, pattern="i")] == [[0, 1, 2, 3, 4]]
    # staging buffers rotate: with rotate=2 the third batch reuses the first batch's storage
    bs = []
    for b in ds.batches(2, shuffle=False, rotate=2):
        bs.append(b["audio_original"])
    assert bs[0].data_ptr() != bs[1].data_ptr() and bs[0][:1].data_ptr() == bs[2].data_ptr()


# ---- MOSI / MOSEI (MML_Suite/data/mosi.py) -----------------------------------------------------------------------------------------
MOSI_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mosi_loader.npz")


@pytest.fixture(scope="module")
def mgold():
    return np.load(MOSI_GOLD)


@pytest.fixture(scope="module")
def pkl(mgold, tmp_path_factory):
    import pickle

    raw = {}
    for k in mgold.files:
        if k.startswith("raw_"):
            _, split, name = k.split("_", 2)
            raw.setdefault(split, {})[name] = mgold[k]
    path = str(tmp_path_factory.mktemp("mosi_files") / "mosi.pkl")
    with open(path, "wb") as f:
        pickle.dump(raw, f)
    return path


def _check_mosi_items(gold, prefix, items, ds):
    assert [str(k) for k in items[0].keys()] == list(gold[f"{prefix}_keys"])
    lab = torch.stack([it["label"] for it in items])
    assert lab.numpy().dtype == gold[f"{prefix}_label"].dtype and np.array_equal(lab.numpy(), gold[f"{prefix}_label"])
    assert [it["sample_idx"] for it in items] == list(gold[f"{prefix}_sample_idx"])
    assert [it["pattern_name"] for it in items] == list(gold[f"{prefix}_pattern"])
    if f"{prefix}_audio_length" in gold.files:
        assert [float(it["audio_length"]) for it in items] == list(gold[f"{prefix}_audio_length"])
        assert [float(it["video_length"]) for it in items] == list(gold[f"{prefix}_video_length"])
    else:
        assert "audio_length" not in items[0]
    for mod in ("audio", "video", "text"):
        assert [float(it[f"{mod}_missing_index"]) for it in items] == list(gold[f"{prefix}_{mod}_missing_index"])
        if f"{prefix}_{mod}" not in gold.files:
            assert ds.keys[mod] not in items[0] and f"{mod}_original" not in items[0]
            continue
        for suffix, key in (("", ds.keys[mod]), ("_original", f"{mod}_original"), ("_reverse", f"{mod}_reverse")):
            got = torch.stack([it[key] for it in items])
            assert got.dtype == torch.float32 and np.array_equal(_bits(got), gold[f"{prefix}_{mod}{suffix}"]), (prefix, mod, suffix)


def test_mosi_items_equal_the_reference(mgold, pkl):
    from mml_b200.datasets import MOSEI, MOSI

    ds = MOSI(pkl, "valid", masks=_masks(mgold, "valid"))
    assert len(ds) == 21 and ds.selected_patterns == ["a", "at", "atv", "av", "t", "tv", "v"] and ds.get_full_modality() == "atv"
    assert MOSI.get_num_classes() == 3 and MOSEI.get_num_classes(False) == 1 and ds.NUM_CLASSES == 3
    _check_mosi_items(mgold, "valid", [ds[i] for i in range(len(ds))], ds)
    mp = {"atv": {"audio": 0.8, "text": 1.0, "video": 1.0}, "t": {"audio": 0.0, "text": 1.0, "video": 0.0}}
    tr = MOSI(pkl, "train", missing_patterns=mp, selected_patterns=["atv", "t"], aligned=True, length=6, masks=_masks(mgold, "train"))
    assert len(tr) == 4 and tr.length == 6
    random.seed(5)
    _check_mosi_items(mgold, "train", [tr[i] for i in (3, 0, 1, 1, 2)], tr)
    te = MOSI(pkl, "test", "text", selected_patterns=["atv"], labels_key="regression_labels", masks=_masks(mgold, "testt"))
    _check_mosi_items(mgold, "testt", [te[i] for i in range(len(te))], te)
    with pytest.raises(KeyError):
        MOSI(pkl, "valid", labels_key="nope")
    # torch's default collation (what the reference's loaders use: its own collate_fn indexes b[""], data/mosi.py:231)
    from torch.utils.data import DataLoader

    b = next(iter(DataLoader(ds, batch_size=5)))
    assert b["label"].shape == (5,) and b["audio_original"].shape == (5, 6, 3) and b[ds.keys["text"]].shape == (5, 6, 8) and b["pattern_name"] == ["a"] * 3 + ["at"] * 2


def test_mosi_batches_equal_the_item_path(mgold, pkl):
    from mml_b200.datasets import MOSI

    ds = MOSI(pkl, "valid", masks=_masks(mgold, "valid"), pin=False)
    items = [ds[i] for i in range(len(ds))]
    got = [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.items()} for b in ds.batches(8, rotate=1)]
    assert [len(b["label"]) for b in got] == [8, 8, 5]
    assert sum((b["pattern_name"] for b in got), []) == [it["pattern_name"] for it in items]
    cat = {k: torch.cat([b[k] for b in got]) for k in got[0] if k != "pattern_name"}
    assert torch.equal(cat["label"], torch.stack([it["label"] for it in items])) and cat["sample_idx"].tolist() == [it["sample_idx"] for it in items]
    assert torch.equal(cat["audio_length"], torch.stack([it["audio_length"] for it in items]))
    for mod in ("audio", "video", "text"):
        assert np.array_equal(_bits(cat[f"{mod}_original"]), _bits(torch.stack([it[f"{mod}_original"] for it in items])))
        assert cat[f"{mod}_missing_index"].tolist() == [float(it[f"{mod}_missing_index"]) for it in items]
        masked = cat[f"{mod}_original"] * cat[f"{mod}_missing_index"].reshape(-1, 1, 1)
        assert np.array_equal(_bits(masked), _bits(torch.stack([it[ds.keys[mod]] for it in items])))
    mono = MOSI(pkl, "test", "text", selected_patterns=["atv"], pin=False)
    b = next(iter(mono.batches(2)))
    assert "text_original" in b and "audio_original" not in b and b["label"].dtype == torch.long


def test_background_batches_equal_inline_batches(gold, csv):
    ds = AVMNIST(csv, "train", cmap=gold["table"], selected_patterns=["ai", "i"], pin=False)
    inline = [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.items()} for b in ds.batches(2, generator=torch.Generator().manual_seed(4))]
    for ahead in (1, 3):
        n = 0
        for want, got in zip(inline, ds.background_batches(2, ahead=ahead, generator=torch.Generator().manual_seed(4))):
            assert set(want) == set(got) and want["pattern_name"] == got["pattern_name"]
            assert all(torch.equal(want[k], got[k]) for k in want if k != "pattern_name")
            n += 1
        assert n == len(inline) == 3
    with pytest.raises(ValueError, match="rotate"):
        next(ds.background_batches(2, ahead=3, rotate=4))
    with pytest.raises(ValueError, match="image_form"):  # an error on the worker thread surfaces in the consumer
        next(ds.background_batches(2, image_form="png"))
    it = ds.background_batches(1)  # abandoning the iterator stops the worker
    next(it)
    it.close()


# ---- MMIMDb (MML_Suite/data/mmimdb.py) ----------------------------------------------------------------------------------------------
MMIMDB_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mmimdb_loader.npz")


class _DictH5(dict):
    """What ``h5py.File`` is to the dataset class: keys(), [key] -> array, close() (h5py is not installed in this image)."""

    def close(self):
        self.closed = True


def _check_mmimdb_items(gold, prefix, items, ds):
    assert [str(k) for k in items[0].keys()] == list(gold[f"{prefix}_keys"])
    lab = torch.stack([it["label"] for it in items])
    assert lab.dtype == torch.float32 and np.array_equal(lab.numpy(), gold[f"{prefix}_label"])
    assert [it["sample_idx"] for it in items] == list(gold[f"{prefix}_sample_idx"])
    assert [it["pattern_name"] for it in items] == list(gold[f"{prefix}_pattern"])
    assert [ds._load_id(it["sample_idx"]) for it in items] == list(gold[f"{prefix}_ids"])
    for mod in ("image", "text"):
        assert [float(it[f"{mod}_missing_index"]) for it in items] == list(gold[f"{prefix}_{mod}_missing_index"])
        if f"{prefix}_{mod}" not in gold.files:
            assert ds.keys[mod] not in items[0] and f"{mod}_original" not in items[0]
            continue
        for suffix, key in (("", ds.keys[mod]), ("_original", f"{mod}_original"), ("_reverse", f"{mod}_reverse")):
            got = torch.stack([it[key] for it in items])
            assert got.dtype == torch.float32 and np.array_equal(_bits(got), gold[f"{prefix}_{mod}{suffix}"]), (prefix, mod, suffix)


def test_mmimdb_items_and_batches(monkeypatch, tmp_path):
    import mml_b200.datasets as D

    gold = np.load(MMIMDB_GOLD)
    raw = _DictH5({k[4:]: gold[k] for k in gold.files if k.startswith("raw_")})
    opened = []
    monkeypatch.setattr(D, "_open_h5", lambda path: (opened.append(str(path)), raw)[1])
    fp = tmp_path / "val.h5"
    fp.write_bytes(b"")
    ds = D.MMIMDb(str(fp), "val", masks=_masks(gold, "val"), pin=False)
    assert opened == [str(fp)] and raw.closed and len(ds) == 15 and ds.selected_patterns == ["i", "it", "t"] and ds.get_full_modality() == "it"
    items = [ds[i] for i in range(len(ds))]
    _check_mmimdb_items(gold, "val", items, ds)
    mp = {"it": {"image": 0.3, "text": 0.7}, "t": {"image": 0.0, "text": 1.0}}
    tr = D.MMIMDb(str(fp), "train", missing_patterns=mp, selected_patterns=["it", "t"], masks=_masks(gold, "train"), pin=False)
    random.seed(8)
    _check_mmimdb_items(gold, "train", [tr[i] for i in (1, 4, 4, 0, 2)], tr)
    te = D.MMIMDb(str(fp), "test", "text", selected_patterns=["it"], masks=_masks(gold, "testt"), pin=False)
    _check_mmimdb_items(gold, "testt", [te[i] for i in range(len(te))], te)
    with pytest.raises(AssertionError, match="Labels key"):
        D.MMIMDb(str(fp), "val", labels_key="nope")
    with pytest.raises(FileNotFoundError):
        D.MMIMDb(str(tmp_path / "missing.h5"), "val")
    with pytest.raises(AssertionError):
        D.MMIMDb(str(fp), "valid")  # this dataset's evaluation split is called "val" (data/mmimdb.py:26)
    # batch path == item path; from_arrays == the file path
    same = D.MMIMDb.from_arrays(gold["raw_genres"], gold["raw_vgg_features"], gold["raw_features"], None, "val", masks=_masks(gold, "val"), pin=False)
    for dset in (ds, same):
        got = [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.items()} for b in dset.batches(4, rotate=1)]
        assert [len(b["label"]) for b in got] == [4, 4, 4, 3] and sum((b["pattern_name"] for b in got), []) == [it["pattern_name"] for it in items]
        cat = {k: torch.cat([b[k] for b in got]) for k in got[0] if k != "pattern_name"}
        assert torch.equal(cat["label"], torch.stack([it["label"] for it in items])) and cat["label"].shape == (15, 23)
        for mod in ("image", "text"):
            assert np.array_equal(_bits(cat[f"{mod}_original"]), _bits(torch.stack([it[f"{mod}_original"] for it in items])))
            masked = cat[f"{mod}_original"] * cat[f"{mod}_missing_index"].reshape(-1, 1)
            assert np.array_equal(_bits(masked), _bits(torch.stack([it[ds.keys[mod]] for it in items])))


def test_mmimdb_without_h5py_says_so(tmp_path):
    import importlib.util

    import mml_b200.datasets as D

    import sys

    if "h5py" in sys.modules or importlib.util.find_spec("h5py") is not None:
        pytest.skip("h5py is installed (or stubbed by oracle/ref_import.py earlier in this process)")
    fp = tmp_path / "train.h5"
    fp.write_bytes(b"")
    with pytest.raises(ImportError, match="h5py"):
        D.MMIMDb(str(fp), "train")


def test_data_parallel_shards_partition_every_global_batch(gold, csv):
    """rank r takes rows [r*B, (r+1)*B) of each global batch; same generator state on every rank => same epoch order."""
    n = 23
    g = torch.Generator().manual_seed(0)
    ds = AVMNIST.from_arrays(torch.arange(n) % 10, torch.rand(n, 3, 4, generator=g), torch.randint(0, 256, (n, 2, 2), dtype=torch.uint8, generator=g),
                             "train", selected_patterns=["ai", "a"], cmap=gold["table"], generator=g, pin=False)
    B, world = 3, 2
    one = [b["sample_idx"].tolist() for b in ds.batches(B * world, drop_last=True, generator=torch.Generator().manual_seed(7))]
    one_p = [b["pattern_name"] for b in ds.batches(B * world, drop_last=True, generator=torch.Generator().manual_seed(7))]
    per_rank = [[(b["sample_idx"].tolist(), b["pattern_name"], b["audio_missing_index"].tolist())
                 for b in ds.batches(B, drop_last=True, generator=torch.Generator().manual_seed(7), rank=r, world=world)] for r in range(world)]
    assert len(one) == 3 and all(len(x) == 3 for x in per_rank)      # 23 samples -> three global batches of 6, five dropped
    for k in range(3):                                                 # the ranks' batches concatenate to the single-process global batch
        assert per_rank[0][k][0] + per_rank[1][k][0] == one[k] and per_rank[0][k][1] + per_rank[1][k][1] == one_p[k]
        for r in range(world):
            idx, pats, masks = per_rank[r][k]
            assert masks == [float(ds.masks[p]["audio"][i]) for i, p in zip(idx, pats)]
    with pytest.raises(ValueError, match="drop_last"):
        next(ds.batches(B, rank=0, world=world))
    with pytest.raises(ValueError, match="rank"):
        next(ds.batches(B, rank=2, world=2, drop_last=True))
    # evaluation: nothing is dropped, the ragged tail is split into near-equal contiguous shards, dataset order is kept
    ev = AVMNIST.from_arrays(torch.arange(n) % 10, torch.rand(n, 3, 4), torch.zeros(n, 2, 2, dtype=torch.uint8), "valid", selected_patterns=["ai"],
                             cmap=gold["table"], pin=False)
    for world in (2, 3, 4):
        shards = [[b["sample_idx"].tolist() for b in ev.batches(4, rank=r, world=world)] for r in range(world)]
        flat = sorted(i for sh in shards for b in sh for i in b)
        assert flat == list(range(n)), world
        nb = -(-n // (4 * world))
        assert all(len(sh) in (nb, nb - 1) for sh in shards) and all(len(b) <= 4 for sh in shards for b in sh)
    # a tail smaller than the world: the last ranks get nothing for it (validation has no collective, so step counts may differ)
    tiny = AVMNIST.from_arrays(torch.arange(9) % 10, torch.rand(9, 3, 4), torch.zeros(9, 2, 2, dtype=torch.uint8), "test", selected_patterns=["ai"],
                               cmap=gold["table"], pin=False)
    assert [[b["sample_idx"].tolist() for b in tiny.batches(2, rank=r, world=4)] for r in range(4)] == [[[0, 1], [8]], [[2, 3]], [[4, 5]], [[6, 7]]]


def test_loader_fixtures_regenerate_from_the_reference(tmp_path):
    """Build container only (the reference is not on the GPU box): ``oracle/make_golden_loader.py`` run afresh over the unmodified reference
    classes writes exactly the three committed ``*_loader.npz`` fixtures."""
    import subprocess
    import sys

    import ref_import

    if not ref_import.reference_available():
        pytest.skip("reference not mounted")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MML_GOLDEN_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(root, "oracle", "make_golden_loader.py")], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    for name in ("avmnist_loader.npz", "mosi_loader.npz", "mmimdb_loader.npz"):
        new, old = np.load(str(tmp_path / name)), np.load(os.path.join(root, "tests", "golden", name))
        assert sorted(new.files) == sorted(old.files), name
        for k in old.files:
            assert new[k].dtype == old[k].dtype and np.array_equal(new[k], old[k]), (name, k)
