"""CPU side of the device input path: the oracle's Philox4x32-10 against the published Random123 known-answer
vectors, the draw rule, and the host-built luminance table against PIL + torchvision themselves (the reference's
``AVMNIST._load_image`` chain, MML_Suite/data/avmnist.py:188-191)."""
import numpy as np
import pytest
import torch

import staging_oracle as S

# Random123 (Salmon et al., SC'11) kat_vectors, philox4x32 with 10 rounds: (counter, key) -> output
PHILOX_KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_philox_known_answers():
    for ctr, key, want in PHILOX_KAT:
        got = S.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array(key, dtype=np.uint32))[0]
        assert tuple(int(v) for v in got) == want


def test_draw_rule_and_shard_invariance():
    p = [0.8, 1.0, 0.0, 0.37]
    m = S.draw_masks(p, 40001, seed=0x1234_5678_9ABC_DEF0, stream_id=2)
    assert m.dtype == np.float32 and set(np.unique(m)) <= {0.0, 1.0}
    assert m[1].all() and not m[2].any()                       # P = 1 always present, P = 0 never
    assert abs(m[0].mean() - 0.8) < 0.01 and abs(m[3].mean() - 0.37) < 0.01
    # a shard of the sample range (one data-parallel rank's slice) reproduces the single draw, at any (unaligned) offset
    for first, n in ((0, 5), (3, 1), (3333, 5001), (39999, 2)):
        assert np.array_equal(S.draw_masks(p, n, 0x1234_5678_9ABC_DEF0, 2, first_sample=first), m[:, first:first + n])
    # streams (patterns), modalities and seeds are independent
    assert not np.array_equal(S.draw_masks(p, 4096, 7, 0)[0], S.draw_masks(p, 4096, 7, 1)[0])
    assert not np.array_equal(S.draw_masks([0.5, 0.5], 4096, 7, 0)[0], S.draw_masks([0.5, 0.5], 4096, 7, 0)[1])
    assert not np.array_equal(S.draw_masks(p, 4096, 7, 0)[0], S.draw_masks(p, 4096, 8, 0)[0])
    assert S.draw_masks(p, 0, 1).shape == (4, 0)


def _reference_image_chain(img_u8: np.ndarray, table: np.ndarray, scale: str) -> torch.Tensor:
    """MML_Suite/data/avmnist.py:188-191 with the colormap call replaced by what matplotlib does for integer input: index the
    256-colour table (``Colormap.__call__`` takes integer arrays as table indices)."""
    from PIL import Image
    from torchvision.transforms.v2 import PILToTensor, ToDtype

    rgba = table[img_u8]                                                        # == cm.<name>(img_u8)
    img = Image.fromarray(np.uint8(rgba * 255)).convert("L")
    if scale == "mul":
        return ToDtype(torch.float32, scale=True)(PILToTensor()(img))[0]        # data/avmnist.py:93-94,190-191
    return torch.from_numpy(np.array(img)).float() / 255.0                      # train_monomodal.py:55-62


@pytest.mark.parametrize("scale", ["mul", "div"])
@pytest.mark.parametrize("channels", [4, 3])
def test_luma_lut_matches_pil_torchvision(scale, channels):
    from mml_b200.data import luma_lut

    rng = np.random.default_rng(5)
    # a smooth multi-segment colour table (the shape of a LinearSegmentedColormap) plus noise, alpha = 1
    x = np.linspace(0, 1, 256)
    table = np.stack([np.clip(np.interp(x, [0, .3, .7, 1], [0, .2, .9, 1]) + rng.normal(0, .02, 256), 0, 1),
                      np.clip(np.interp(x, [0, .5, 1], [0, .7, 1]) + rng.normal(0, .02, 256), 0, 1),
                      np.clip(np.interp(x, [0, .2, 1], [.4, .1, 1]) + rng.normal(0, .02, 256), 0, 1),
                      np.ones(256)], axis=1)[:, :channels]
    img = rng.integers(0, 256, size=(28, 28), dtype=np.uint8)
    img[0, :4] = [0, 255, 1, 254]
    want = _reference_image_chain(img, table, scale)
    lut = luma_lut(table, scale)
    assert lut.dtype == torch.float32 and lut.shape == (256,)
    assert np.array_equal(lut.numpy(), S.luma_lut(table, scale))               # product host code == oracle
    got = torch.from_numpy(S.u8_lut(img, lut.numpy()))
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))           # bit for bit with PIL + torchvision


def test_luma_lut_rejects_bad_tables():
    from mml_b200.data import luma_lut

    with pytest.raises(ValueError):
        luma_lut(np.zeros((255, 3)))
    with pytest.raises(ValueError):
        luma_lut(np.zeros((256, 3)), scale="x")


def test_host_philox_table_equals_the_oracle_draw_and_feeds_the_datasets():
    """``data.philox_missing_masks`` (product, host) against ``staging_oracle.draw_masks`` (which tests/test_staging_gpu.py holds the GPU
    kernel to, bit for bit): host datasets, data-parallel ranks and ``DeviceMaskTable`` share one mask table per seed."""
    from mml_b200.data import generate_patterns, philox_missing_masks
    from mml_b200.datasets import AVMNIST

    pats = generate_patterns({"audio": (0.2, None), "image": (0.4, ["i"])})
    seed = 0xFEDC_BA98_7654_3210
    tab = philox_missing_masks(pats, 5003, seed)
    assert list(tab) == list(pats)
    for k, (pat, probs) in enumerate(pats.items()):
        want = S.draw_masks(list(probs.values()), 5003, seed, k)
        for j, m in enumerate(probs):
            assert tab[pat][m].dtype == torch.float32 and np.array_equal(tab[pat][m].numpy().view(np.uint32), want[j].view(np.uint32)), (pat, m)
    # a shard of the sample range reproduces the slice (unaligned offsets, offsets beyond 2^32 blocks)
    for first, n in ((3, 1), (1001, 2002), (2 ** 34 + 1, 7)):
        part = philox_missing_masks(pats, n, seed, first_sample=first)
        want = S.draw_masks(list(pats["ai"].values()), n, seed, list(pats).index("ai"), first_sample=first)
        assert np.array_equal(part["ai"]["audio"].numpy(), want[0]) and np.array_equal(part["ai"]["image"].numpy(), want[1])
    assert philox_missing_masks(pats, 0, seed)["ai"]["audio"].shape == (0,)
    # datasets: mask_seed => the same table in every process, whatever the torch / Python RNG state
    n = 37
    mk = lambda: AVMNIST.from_arrays(torch.arange(n) % 10, torch.rand(n, 2, 3), torch.zeros(n, 2, 2, dtype=torch.uint8), "train", missing_patterns=pats,
                                     selected_patterns=["ai", "i"], cmap=np.zeros((256, 3)), pin=False, mask_seed=99)
    a, b = mk(), mk()
    full = philox_missing_masks(pats, n, 99)
    for pat in pats:
        for m in ("audio", "image"):
            assert torch.equal(a.masks[pat][m], b.masks[pat][m]) and torch.equal(a.masks[pat][m], full[pat][m])
    assert 0.6 < float(a.masks["ai"]["audio"].mean()) <= 1.0 and not a.masks["i"]["audio"].any()
