"""GPU parity of the device input path (csrc/staging.cu through the C ABI) against oracle/staging_oracle.py: BIT-EXACT
(integer / byte work).  Mask draw: base_dataset.py:46-59; mask lookup: data/avmnist.py:193-224; image conversion: :188-191."""
import numpy as np
import pytest
import torch

import staging_oracle as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,first", [(1, 0), (5, 3), (4096, 0), (60000, 0), (5001, 3333), (2, 39999), (7, 2 ** 34 + 1)])
def test_mask_draw_bit_exact(n, first):
    from mml_b200 import ops

    p = torch.tensor([0.8, 1.0, 0.0, 0.37], device="cuda")
    seed = 0x1234_5678_9ABC_DEF0
    got = ops.missing_mask_draw(p, n, seed, stream_id=2, first_sample=first)
    want = S.draw_masks(p.cpu().numpy(), n, seed, 2, first_sample=first)
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))


def test_mask_draw_empty_and_shards_agree():
    from mml_b200 import ops

    p = torch.tensor([0.5, 0.9], device="cuda")
    assert ops.missing_mask_draw(p, 0, 1).shape == (2, 0)
    whole = ops.missing_mask_draw(p, 10000, 99, stream_id=1)
    parts = [ops.missing_mask_draw(p, 2500, 99, stream_id=1, first_sample=2500 * r) for r in range(4)]  # four ranks' slices
    assert torch.equal(torch.cat(parts, dim=1), whole)
    assert not torch.equal(ops.missing_mask_draw(p, 10000, 99, stream_id=0), whole)


def test_mask_table_and_gather():
    from mml_b200.data import DeviceMaskTable, generate_patterns

    pats = generate_patterns({"audio": (0.2, None), "image": (0.4, ["i"])})
    tab = DeviceMaskTable(pats, 5000, seed=11, device="cuda")
    for k, (pat, probs) in enumerate(pats.items()):
        want = S.draw_masks(list(probs.values()), 5000, 11, k)
        assert np.array_equal(tab.masks[pat].cpu().numpy(), want), pat
    idx = torch.randint(0, 5000, (256,), generator=torch.Generator().manual_seed(0))
    b = tab.batch("ai", idx)
    want = S.gather_masks(S.draw_masks(list(pats["ai"].values()), 5000, 11, list(pats).index("ai")), idx.numpy())
    assert list(b) == list(pats["ai"])
    for j, m in enumerate(b):
        assert b[m].shape == (256,) and np.array_equal(b[m].cpu().numpy(), want[j])
    tab.check_indices()
    tab.batch("ai", torch.tensor([0, 5000]))
    with pytest.raises(IndexError):
        tab.check_indices()


@pytest.mark.parametrize("shape", [(0,), (1,), (15,), (16,), (17,), (256, 28, 28), (3, 5, 7)])
def test_u8_lut_bit_exact(shape):
    from mml_b200 import ops
    from mml_b200.data import luma_lut

    rng = np.random.default_rng(3)
    table = rng.random((256, 4))
    lut = luma_lut(table)
    src = torch.from_numpy(rng.integers(0, 256, size=shape, dtype=np.uint8))
    got = ops.u8_lut(src.cuda(), lut.cuda())
    want = S.u8_lut(src.numpy(), S.luma_lut(table))
    assert got.shape == src.shape and got.dtype == torch.float32
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))


def test_prefetcher_expands_uint8_images_on_the_device():
    from mml_b200.data import DevicePrefetcher, luma_lut

    rng = np.random.default_rng(9)
    lut = luma_lut(rng.random((256, 3)))
    batches = [{"image_original": torch.from_numpy(rng.integers(0, 256, size=(32, 28, 28), dtype=np.uint8)).pin_memory(),
                "labels": torch.arange(32), "pattern_name": ["ai"] * 32} for _ in range(5)]
    pf = DevicePrefetcher(batches, "cuda", luts={"image_original": lut})
    n = 0
    for ref, out in zip(batches, pf):
        x = out["image_original"]
        assert x.is_cuda and x.dtype == torch.float32
        assert torch.equal(x.cpu(), lut[ref["image_original"].long()])
        assert torch.equal(out["labels"].cpu(), ref["labels"]) and out["pattern_name"] == ref["pattern_name"]
        n += 1
    assert n == 5
    assert pf.h2d_bytes == 5 * (32 * 28 * 28 + 32 * 8)  # one byte per pixel crosses PCIe


def test_device_mask_table_equals_the_host_philox_table():
    """One mask table per seed: what ``DeviceMaskTable`` draws on the GPU == what ``data.philox_missing_masks`` computes on the host (the
    table ``datasets.*(mask_seed=...)`` uses), both == oracle/staging_oracle.py."""
    from mml_b200.data import DeviceMaskTable, generate_patterns, philox_missing_masks

    pats = generate_patterns({"audio": (0.2, None), "image": (0.4, ["i"])})
    tab = DeviceMaskTable(pats, 5000, seed=11, device="cuda")
    host = philox_missing_masks(pats, 5000, 11)
    for pat, probs in pats.items():
        dev = tab.masks[pat].cpu()
        for j, m in enumerate(probs):
            assert torch.equal(dev[j], host[pat][m]), (pat, m)
