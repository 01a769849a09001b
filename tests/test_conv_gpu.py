"""GPU parity of the tcgen05 implicit-GEMM convolutions (through the C ABI) against torch on the same bf16 operands.

Every (C, K, H, W, stride) below is a layer shape of the reference's ResNet18 @112x112 / ResNet34 @28x28 encoders
(MML_Suite/models/msa/networks/resnet.py:25,30,176) plus the real 32x94 spectrogram shapes and
ragged batch sizes.  Tolerance: operands are identical bf16 values, accumulation is fp32 in both, the result is stored
as bf16 => |err| <= 2^-8 relative to the output magnitude plus fp32 reordering noise.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False  # the torch reference convolution must be true fp32
torch.backends.cuda.matmul.allow_tf32 = False

# (N, H, W, C, K, R, stride, pad)
SHAPES = [
    (8, 28, 28, 64, 64, 3, 1, 1),     # R18 layer1
    (8, 28, 28, 64, 128, 3, 2, 1),    # R18 layer2.0.conv1
    (8, 28, 28, 64, 128, 1, 2, 0),    # R18 layer2.0.downsample
    (8, 14, 14, 128, 128, 3, 1, 1),   # R18 layer2
    (8, 14, 14, 128, 256, 3, 2, 1),   # R18 layer3.0.conv1
    (8, 14, 14, 128, 256, 1, 2, 0),
    (8, 7, 7, 256, 256, 3, 1, 1),     # R18 layer3 / R34 layer1 spatial
    (8, 7, 7, 256, 512, 3, 2, 1),     # R18 layer4.0.conv1 (7 -> 4, odd input)
    (8, 7, 7, 256, 512, 1, 2, 0),
    (16, 4, 4, 512, 512, 3, 1, 1),    # R18 layer4
    (6, 7, 7, 64, 64, 3, 1, 1),       # R34 layer1
    (6, 7, 7, 64, 128, 3, 2, 1),      # R34 layer2.0 (7 -> 4)
    (6, 4, 4, 128, 128, 3, 1, 1),
    (6, 4, 4, 128, 256, 3, 2, 1),     # 4 -> 2
    (40, 2, 2, 256, 256, 3, 1, 1),    # R34 layer3
    (40, 2, 2, 256, 512, 3, 2, 1),    # 2 -> 1
    (40, 2, 2, 256, 512, 1, 2, 0),
    (130, 1, 1, 512, 512, 3, 1, 1),   # R34 layer4 (only the centre tap reaches the input), ragged N
    (3, 8, 24, 64, 64, 3, 1, 1),      # real 32x94 audio: layer1 is 8x24
    (3, 8, 24, 64, 128, 3, 2, 1),
    (5, 2, 6, 256, 512, 3, 2, 1),     # -> 1x3
    (5, 1, 3, 512, 512, 3, 1, 1),
    # BASELINE.json configs[1] (batch 256 per GPU): the shapes bench.py runs.  Only at this size do the persistent kernels walk
    # several tiles per CTA (halo fprop / dgrad: ~12 tiles per CTA -> TMEM double-buffer parity, patch-ring wrap; halo wgrad:
    # multi-tile TMEM accumulation) and do the split-K kernels use every split.
    (256, 28, 28, 64, 64, 3, 1, 1),   # R18 layer1 (halo kernel, weights resident)
    (256, 14, 14, 128, 128, 3, 1, 1), # R18 layer2 (halo kernel, weight ring)
    (256, 7, 7, 256, 256, 3, 1, 1),   # R18 layer3
    (256, 14, 14, 128, 256, 3, 2, 1), # R18 layer3.0.conv1 (stride 2)
    (256, 4, 4, 512, 512, 3, 1, 1),   # R18 layer4
    (256, 28, 28, 64, 128, 3, 2, 1),  # R18 layer2.0.conv1
    (256, 7, 7, 64, 64, 3, 1, 1),     # R34 layer1
    (256, 2, 2, 256, 256, 3, 1, 1),   # R34 layer3
    (256, 1, 1, 512, 512, 3, 1, 1),   # R34 layer4
    (250, 28, 28, 64, 64, 3, 1, 1),   # ragged batch on the persistent path (tile count not a multiple of the grid)
    # stride-2 dgrad = ONE launch over the four output phases (grid.z): equal phases (14x14 -> 7x7 each), unequal phases of an odd
    # tensor (7x7 -> 4x4, 4x3, 3x4, 3x3: different tile shapes and tile counts per phase), ragged batch
    (256, 7, 7, 256, 512, 3, 2, 1),   # R18 layer4.0.conv1
    (256, 7, 7, 64, 128, 3, 2, 1),    # R34 layer2.0.conv1
    (256, 4, 4, 128, 256, 3, 2, 1),   # R34 layer3.0.conv1
    (250, 2, 2, 256, 512, 3, 2, 1),   # R34 layer4.0.conv1 (2x2 -> phases of 1x1)
    (37, 7, 5, 64, 64, 3, 2, 1),      # odd x odd, ragged
    (256, 7, 7, 256, 512, 1, 2, 0),   # 1x1 stride 2: three of the four phases receive no tap (zero-filled)
]


def _mk(shape, seed):
    N, H, W, C, K, R, st, pad = shape
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(K, R, R, C, device="cuda", generator=g) * (2.0 / (C * R * R)) ** 0.5).to(torch.bfloat16)
    return x, w


def _ref_conv(x, w, st, pad):
    return torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), stride=st, padding=pad)


def _report(name, got, ref, tol):
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, f"{name}: max|err|={err:.4g} vs max|ref|={scale:.4g} (tol {tol})"


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_fprop_and_stats(shape):
    from mml_b200 import ops

    N, H, W, C, K, R, st, pad = shape
    x, w = _mk(shape, 1)
    geom = ops.make_geom(N, H, W, C, K, R, R, st, pad)
    P, Q = ops.conv_out_hw(H, W, R, R, st, pad)
    y = torch.full((N, P, Q, K), float("nan"), device="cuda", dtype=torch.bfloat16)
    stats = ops.bn_stats_buffer(K, "cuda")
    ops.conv_fprop(geom, x, w, y, stats)
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, st, pad).permute(0, 2, 3, 1)
    _report("fprop", y.float(), ref, 2.0 ** -7)
    yf = y.double().reshape(-1, K)
    stats = stats.sum(0)
    assert torch.allclose(stats[:, 0], yf.sum(0), rtol=1e-5, atol=1e-3 * yf.abs().max().item()), "BN sum"
    assert torch.allclose(stats[:, 1], (yf * yf).sum(0), rtol=1e-5, atol=1e-4), "BN sum of squares"


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_dgrad(shape):
    from mml_b200 import ops

    N, H, W, C, K, R, st, pad = shape
    x, w = _mk(shape, 2)
    geom = ops.make_geom(N, H, W, C, K, R, R, st, pad)
    P, Q = ops.conv_out_hw(H, W, R, R, st, pad)
    g = torch.Generator(device="cuda").manual_seed(3)
    dy = torch.randn(N, P, Q, K, device="cuda", generator=g).to(torch.bfloat16)
    dx = torch.full((N, H, W, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.conv_dgrad(geom, dy, w, dx)  # same K,R,S,C weights as fprop
    torch.cuda.synchronize()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    out = torch.nn.functional.conv2d(xr, w.float().permute(0, 3, 1, 2), stride=st, padding=pad)
    out.backward(dy.float().permute(0, 3, 1, 2))
    _report("dgrad", dx.float(), xr.grad.permute(0, 2, 3, 1), 2.0 ** -7)


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_wgrad(shape):
    from mml_b200 import ops

    N, H, W, C, K, R, st, pad = shape
    x, w = _mk(shape, 4)
    geom = ops.make_geom(N, H, W, C, K, R, R, st, pad)
    P, Q = ops.conv_out_hw(H, W, R, R, st, pad)
    g = torch.Generator(device="cuda").manual_seed(5)
    dy = torch.randn(N, P, Q, K, device="cuda", generator=g).to(torch.bfloat16)
    dw = torch.full((K, R, R, C), float("nan"), device="cuda")  # overwritten, never read
    ws = torch.empty(max(ops.conv_wgrad_workspace(geom, "cuda") // 4, 1), device="cuda")
    ops.conv_wgrad(geom, x, dy, dw, ws)
    torch.cuda.synchronize()
    wr = w.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    out = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wr, stride=st, padding=pad)
    out.backward(dy.float().permute(0, 3, 1, 2))
    ref = wr.grad.permute(0, 2, 3, 1)
    _report("wgrad", dw, ref, 1e-3)
    # deterministic (fixed-order sum of the split partials, no atomics): a second call reproduces the result bit for bit
    dw2 = torch.full((K, R, R, C), float("nan"), device="cuda")
    ops.conv_wgrad(geom, x, dy, dw2, ws)
    torch.cuda.synchronize()
    assert torch.equal(dw, dw2), "wgrad is not bit-reproducible"


SPLITK_SHAPES = [
    (256, 2, 2, 256, 256, 3, 1, 1),   # R34 layer3 at batch 256: 8 pixel tiles x 2 channel tiles
    (256, 1, 1, 512, 512, 3, 1, 1),   # R34 layer4: one reachable tap, 8 K chunks
    (130, 1, 1, 512, 512, 3, 1, 1),   # ragged batch: rows beyond the tensor are neither stored nor counted
    (256, 4, 4, 128, 128, 3, 1, 1),   # R34 layer2
    (6, 4, 4, 128, 256, 3, 2, 1),     # stride 2: dgrad writes strided phase views of dx
    (40, 2, 2, 256, 512, 1, 2, 0),    # 1x1 stride-2 downsample: 4 K iterations only
    (8, 7, 7, 256, 256, 3, 1, 1),
    (5, 1, 3, 512, 512, 3, 1, 1),
]


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
@pytest.mark.parametrize("shape", SPLITK_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_splitk_cluster_variant(shape, cluster):
    """conv_igemm_splitk_kernel (K loop split over a thread-block cluster, fixed-order DSMEM reduction) for every cluster size,
    fprop + BatchNorm sums and dgrad, and bit-reproducibility of two runs."""
    from mml_b200 import ops

    N, H, W, C, K, R, st, pad = shape
    x, w = _mk(shape, 21)
    geom = ops.make_geom(N, H, W, C, K, R, R, st, pad)
    P, Q = ops.conv_out_hw(H, W, R, R, st, pad)
    g = torch.Generator(device="cuda").manual_seed(22)
    dy = torch.randn(N, P, Q, K, device="cuda", generator=g).to(torch.bfloat16)
    ops.debug_set(2, cluster)
    try:
        outs = []
        for rep in range(2):
            y = torch.full((N, P, Q, K), float("nan"), device="cuda", dtype=torch.bfloat16)
            stats = ops.bn_stats_buffer(K, "cuda")
            ops.conv_fprop(geom, x, w, y, stats)
            dx = torch.full((N, H, W, C), float("nan"), device="cuda", dtype=torch.bfloat16)
            ops.conv_dgrad(geom, dy, w, dx)
            torch.cuda.synchronize()
            outs.append((y, dx, stats.sum(0)))
    finally:
        ops.debug_set(2, 1)
    y, dx, stats = outs[0]
    assert torch.equal(y, outs[1][0]) and torch.equal(dx, outs[1][1]), "not bit-reproducible"
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = torch.nn.functional.conv2d(xr, w.float().permute(0, 3, 1, 2), stride=st, padding=pad)
    ref.backward(dy.float().permute(0, 3, 1, 2))
    _report("fprop", y.float(), ref.detach().permute(0, 2, 3, 1), 2.0 ** -7)
    _report("dgrad", dx.float(), xr.grad.permute(0, 2, 3, 1), 2.0 ** -7)
    yf = y.double().reshape(-1, K)
    assert torch.allclose(stats[:, 0], yf.sum(0), rtol=1e-5, atol=1e-3 * yf.abs().max().item()), "BN sum"
    assert torch.allclose(stats[:, 1], (yf * yf).sum(0), rtol=1e-5, atol=1e-4), "BN sum of squares"
