"""``mml_b200.datasets.AVMNIST.fused_loader`` end to end on the GPU: worker-thread batches -> pinned staging ->
DevicePrefetcher (uint8 image bytes over PCIe, luminance table on the copy stream) -> the fused train / validation step.  The host
side of the loader is pinned against the reference class in tests/test_datasets_cpu.py; here the staged tensors must equal the host
batches and training from the loader must follow the same trajectory as training from host fp32 batches."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class Term:
    def __init__(self):
        self.loss_fn, self.weight = torch.nn.CrossEntropyLoss(), 1.0


LOSS = {"cross_entropy": Term()}


def _dataset(split, n=40, hw=(32, 94)):
    from mml_b200.datasets import AVMNIST

    g = torch.Generator().manual_seed(7)
    audio = torch.rand(n, *hw, generator=g)
    image = torch.randint(0, 256, (n, 28, 28), dtype=torch.uint8, generator=g)
    labels = torch.randint(0, 10, (n,), generator=g)
    table = np.random.default_rng(3).random((256, 4))
    mp = {"ai": {"audio": 0.6, "image": 1.0}, "i": {"audio": 0.0, "image": 1.0}}
    return AVMNIST.from_arrays(labels, audio, image, split, missing_patterns=mp, selected_patterns=["ai", "i"], cmap=table, generator=g)


def _model():
    from mml_b200.avmnist import AVMNIST
    from mml_b200.resnet import ResNet18, ResNet34

    torch.manual_seed(0)
    return AVMNIST(ResNet18(1, 64), ResNet34(1, 128), 128, dropout=0.0).to(DEV)


def _snapshot(b):
    return {k: (v.clone() if torch.is_tensor(v) else list(v)) for k, v in b.items()}


def test_fused_loader_stages_what_the_host_path_yields_and_trains_the_same():
    ds = _dataset("train")
    B = 16
    host = [_snapshot(b) for b in ds.batches(B, image_form="f32", generator=torch.Generator().manual_seed(1))]
    assert [len(b["labels"]) for b in host] == [16, 16, 8] and {p for b in host for p in b["pattern_name"]} == {"ai", "i"}
    loader = ds.fused_loader(DEV, B, generator=torch.Generator().manual_seed(1))
    n = 0
    for h, d in zip(host, loader):
        assert set(d) == set(h)
        for k, v in h.items():
            if torch.is_tensor(v):
                assert d[k].is_cuda and d[k].dtype == v.dtype and d[k].shape == v.shape, k
                assert torch.equal(d[k].cpu(), v), k      # image: expanded on the device == table lookup on the host
            else:
                assert d[k] == v, k
        n += 1
    assert n == 3
    per_batch = lambda b: sum(v.numel() * (1 if k == "image_original" else v.element_size()) for k, v in b.items() if torch.is_tensor(v))
    assert loader.h2d_bytes == sum(per_batch(b) for b in host)  # one byte per image pixel crosses PCIe

    def run(fused):
        model = _model()
        opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
        out = []
        for epoch in range(2):  # two epochs of two full batches each, same shuffles and patterns on both paths
            kw = dict(drop_last=True, generator=torch.Generator().manual_seed(2 + epoch))
            it = ds.fused_loader(DEV, B, **kw) if fused else ds.batches(B, image_form="f32", **kw)
            out += [model.train_step(b, opt, LOSS, torch.device(DEV), None)["loss"] for b in it]
        return model, opt, out

    _, _, from_host = run(False)
    model, opt, from_loader = run(True)
    assert len(from_host) == len(from_loader) == 4 and all(np.isfinite(from_host))
    assert np.allclose(from_host, from_loader, rtol=2e-2, atol=2e-2), (from_host, from_loader)

    with pytest.raises(TypeError, match="luminance table"):
        model.train_step(next(iter(ds.batches(B))), opt, LOSS, torch.device(DEV), None)  # raw uint8 pixels without the table


def test_validation_split_through_the_fused_loader():
    ds = _dataset("valid", n=16)
    assert len(ds) == 32
    model = _model()
    seen = []
    for b in ds.fused_loader(DEV, 16):
        out = model.validation_step(b, LOSS, torch.device(DEV), None, return_test_info=True)
        assert np.isfinite(out["loss"]) and len(out["predictions"]) == 16
        assert np.array_equal(out["labels"], ds.labels.numpy())
        seen.append(sorted(set(out["miss_types"].tolist())))
        if seen[-1] == ["i"]:
            assert not b["audio_missing_index"].any()  # pattern "i": the audio modality is masked out on the device
    assert seen == [["ai"], ["i"]]


def test_mosi_fused_loader_trains_like_the_host_path(tmp_path):
    """``datasets.MOSI`` -> worker thread -> DevicePrefetcher -> ``UttFusionModel.train_step`` (config 4), against the same batches fed from
    the host; the dropout keep-masks are given so that both runs draw the same ones."""
    import pickle

    from mml_b200.datasets import MOSI
    from mml_b200.utt_fusion import FcClassifier, LSTMEncoder, TextCNN, UttFusionModel

    rng = np.random.default_rng(0)
    n, T, B = 70, 50, 32
    raw = {"audio": rng.normal(size=(n, T, 5)).astype(np.float32), "vision": rng.normal(size=(n, T, 20)).astype(np.float32),
           "text": rng.normal(size=(n, T, 768)).astype(np.float32), "classification_labels": rng.integers(0, 3, size=n),
           "audio_lengths": np.full(n, T), "vision_lengths": np.full(n, T)}
    fp = str(tmp_path / "mosi.pkl")
    with open(fp, "wb") as f:
        pickle.dump({"train": raw}, f)
    mp = {"atv": {"audio": 0.8, "text": 1.0, "video": 1.0}, "t": {"audio": 0.0, "text": 1.0, "video": 0.0}, "av": {"audio": 1.0, "text": 0.0, "video": 1.0}}
    ds = MOSI(fp, "train", missing_patterns=mp, selected_patterns=["atv", "t", "av"], generator=torch.Generator().manual_seed(3))
    host = [_snapshot(b) for b in ds.batches(B, generator=torch.Generator().manual_seed(1))]
    assert [len(b["label"]) for b in host] == [32, 32, 6] and {p for b in host for p in b["pattern_name"]} == {"atv", "t", "av"}
    n_batches = 0
    for h, d in zip(host, ds.fused_loader(DEV, B, generator=torch.Generator().manual_seed(1))):
        assert set(d) == set(h)
        for k, v in h.items():
            if torch.is_tensor(v):
                assert d[k].is_cuda and d[k].dtype == v.dtype and torch.equal(d[k].cpu(), v), k
            else:
                assert d[k] == v, k
        n_batches += 1
    assert n_batches == 3

    g = torch.Generator().manual_seed(9)
    keeps = [(torch.rand(B, k, generator=g) >= 0.5).float() for k in (384, 192, 64, 32)]

    def run(fused):
        torch.manual_seed(0)
        model = UttFusionModel(LSTMEncoder(5, 64, "last"), LSTMEncoder(20, 64, "last"),
                               TextCNN(768, embd_size=64, dropout=0.5, in_channels=1, out_channels=128, kernel_heights=[3, 4, 5]),
                               FcClassifier(192, [192, 64, 32], 3, dropout=0.5), clip=1.0).to(DEV)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
        out = []
        for epoch in range(2):
            kw = dict(drop_last=True, generator=torch.Generator().manual_seed(2 + epoch))
            it = ds.fused_loader(DEV, B, **kw) if fused else ds.batches(B, **kw)
            out += [model.train_step(b, opt, LOSS, torch.device(DEV), None, dropout_masks=keeps)["loss"] for b in it]
        return out

    from_host, from_loader = run(False), run(True)
    assert len(from_host) == len(from_loader) == 4 and all(np.isfinite(from_host))
    assert np.allclose(from_host, from_loader, rtol=2e-2, atol=2e-2), (from_host, from_loader)
