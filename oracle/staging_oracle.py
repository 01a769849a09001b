"""CPU oracle of the device-side input path.  TEST INFRASTRUCTURE ONLY: imported by tests/, never by the
product package (mml_b200/) -- the product path fails loudly without libmml_b200.so.

What it restates:

* ``draw_masks``: ``MultimodalBaseDataset._initialise_missing_masks`` (MML_Suite/data/base_dataset.py:46-59) draws, per pattern,
  ``create_missing_mask(n_modalities, n_samples, [P(present) per modality])`` -- one independent Bernoulli per (sample, modality).
  ``create_missing_mask`` lives in the un-vendored ``modalities`` package (git dependency without a pinned rev,
  MML_Suite/pyproject.toml:13), so the reference's *random stream* cannot be reproduced: **sampling parity is unpinned**.  What is
  pinned is the generator itself -- Philox4x32-10 (Salmon et al., SC'11; the Random123 library's known-answer vectors are in
  tests/test_oracle_golden.py) -- and the draw rule ``present iff u < P(present)``, the rule of ``torch.bernoulli``.
* ``luma_lut`` / ``u8_lut``: ``AVMNIST._load_image`` (MML_Suite/data/avmnist.py:188-191):
  ``np.uint8(cm.gist_earth(img) * 255)`` -> ``Image.fromarray(...).convert("L")`` -> ``PILToTensor`` -> ``ToDtype(float32, scale=True)``.
  For uint8 pixels matplotlib indexes the colormap's 256-entry table with the pixel value, so the chain is a table of the pixel
  value.  Pinned here against PIL and torchvision themselves (tests/test_oracle_golden.py::test_luma_lut_matches_pil_torchvision);
  the ``gist_earth`` table is matplotlib's (not installed in this image): the caller passes ``cm.gist_earth(np.arange(256))``.
"""
from __future__ import annotations

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Vectorised Philox4x32-10.  counter: uint32 [..., 4], key: uint32 [..., 2] (broadcastable) -> uint32 [..., 4]."""
    c = [counter[..., i].astype(np.uint64) for i in range(4)]
    k = [np.broadcast_to(key[..., i], counter.shape[:-1]).astype(np.uint64) for i in range(2)]
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k[0], p1 & MASK32, (p0 >> np.uint64(32)) ^ c[3] ^ k[1], p0 & MASK32]
        k = [(k[0] + np.uint64(W0)) & MASK32, (k[1] + np.uint64(W1)) & MASK32]
    return np.stack(c, axis=-1).astype(np.uint32)


def draw_masks(p_present, num_samples: int, seed: int, stream_id: int = 0, first_sample: int = 0) -> np.ndarray:
    """fp32 [n_modalities, num_samples]; sample i = first_sample + j uses word i % 4 of the Philox block with
    counter (lo32(i // 4), hi32(i // 4), modality, stream_id) and key (lo32(seed), hi32(seed))."""
    p = np.asarray(p_present, dtype=np.float32)
    i = np.arange(first_sample, first_sample + num_samples, dtype=np.uint64)
    q = i >> np.uint64(2)
    out = np.zeros((p.size, num_samples), dtype=np.float32)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    for m in range(p.size):
        ctr = np.stack([(q & MASK32), (q >> np.uint64(32)), np.full_like(q, m), np.full_like(q, stream_id & 0xFFFFFFFF)], axis=-1).astype(np.uint32)
        r = philox4x32_10(ctr, key)
        bits = r[np.arange(num_samples), (i & np.uint64(3)).astype(np.int64)]
        u = (bits >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
        out[m] = (u < p[m]).astype(np.float32)
    return out


def gather_masks(masks: np.ndarray, sample_idx) -> np.ndarray:
    return np.ascontiguousarray(masks[:, np.asarray(sample_idx, dtype=np.int64)])


def luma_lut(cmap_table, scale: str = "mul") -> np.ndarray:
    """fp32 [256]: pixel value v -> colormap colour -> ``np.uint8(rgba * 255)`` (truncation) -> PIL "L" (ITU-R 601-2 luma in 16.16
    fixed point: (19595 R + 38470 G + 7471 B + 0x8000) >> 16, alpha ignored) -> float32 scaled to [0, 1].
    ``scale``: "mul" = torchvision ``ToDtype(scale=True)`` (x * (1/255) in fp32, data/avmnist.py:93-94,191); "div" = ``x.float() / 255.0``
    (the loader of train_monomodal.py:55-62)."""
    t = np.asarray(cmap_table, dtype=np.float64)
    assert t.shape[0] == 256 and t.shape[1] in (3, 4), t.shape
    rgb = np.uint8(t[:, :3] * 255).astype(np.uint32)
    L = (rgb[:, 0] * 19595 + rgb[:, 1] * 38470 + rgb[:, 2] * 7471 + 0x8000) >> 16
    Lf = L.astype(np.float32)
    if scale == "mul":
        return (Lf * np.float32(1.0 / 255.0)).astype(np.float32)
    if scale == "div":
        return (Lf / np.float32(255.0)).astype(np.float32)
    raise ValueError(scale)


def u8_lut(src: np.ndarray, lut: np.ndarray) -> np.ndarray:
    return lut[np.asarray(src, dtype=np.uint8)]
