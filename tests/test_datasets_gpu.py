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
