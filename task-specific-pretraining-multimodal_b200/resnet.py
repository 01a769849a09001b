"""ResNet encoders for 1-channel inputs -- host-side mirror of the reference's encoder classes.

Drop-in surface (MML_Suite/models/msa/networks/resnet.py:113-239): ``ResNetEncoder(block, layers, in_channels,
hidden_dim)``, the ``ResNet18`` / ``ResNet34`` factories, ``get_embedding_size()``, ``forward(x)`` accepting
``[B,H,W]`` or ``[B,1,H,W]``, and -- what checkpoints depend on -- the exact ``state_dict()`` (names, shapes, dtypes;
e.g. ``layer2.0.downsample.1.running_var``).  Initialisation draws from torch's RNG in the same order as the reference
constructors, so the same seed produces the same weights (tests/test_modules_cpu.py checks it against the oracle).

The ``nn.Conv2d`` / ``nn.BatchNorm2d`` / ``nn.Linear`` sub-modules are PARAMETER CONTAINERS only: their ``forward`` is
never called.  All arithmetic runs in libmml_b200.so through ``engine.EncoderPlan`` (tcgen05 implicit-GEMM
convolutions, fused BN/ReLU/residual/pooling kernels).  There is no PyTorch fallback: calling ``forward`` on a CPU
tensor raises.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

STAGE_WIDTHS = (64, 128, 256, 512)


class BasicBlock(nn.Module):
    """conv3x3-BN-ReLU-conv3x3-BN (+ identity or 1x1-conv-BN shortcut) -> add -> ReLU (resnet.py:8-54)."""

    expansion = 1

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: Optional[nn.Module] = None, norm_layer=None):
        super().__init__()
        norm_layer = norm_layer or nn.BatchNorm2d
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = norm_layer(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = norm_layer(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):  # pragma: no cover - containers only
        raise RuntimeError("mml_b200.BasicBlock holds parameters only; run the encoder (fused CUDA path)")


class ResNetEncoder(nn.Module):
    def __init__(self, block=BasicBlock, layers: Sequence[int] = (2, 2, 2, 2), in_channels: int = 1, hidden_dim: int = 128,
                 zero_init_residual: bool = False, norm_layer=None):
        super().__init__()
        if block is not BasicBlock and getattr(block, "expansion", 1) != 1:
            raise NotImplementedError("mml_b200 implements the BasicBlock encoders (ResNet18/34) of the late-fusion path")
        if in_channels != 1:
            raise NotImplementedError("mml_b200 stem kernel is specialised for 1-channel inputs (AVMNIST audio / image)")
        if norm_layer not in (None, nn.BatchNorm2d):
            raise NotImplementedError("only nn.BatchNorm2d is supported")
        self.hidden_dim = hidden_dim
        self.layers_cfg = tuple(int(n) for n in layers)
        self.inplanes = 64
        self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = self._stage(64, self.layers_cfg[0], 1)
        self.layer2 = self._stage(128, self.layers_cfg[1], 2)
        self.layer3 = self._stage(256, self.layers_cfg[2], 2)
        self.layer4 = self._stage(512, self.layers_cfg[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512, hidden_dim)
        for m in self.modules():  # resnet.py:153-158
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, BasicBlock):
                    nn.init.constant_(m.bn2.weight, 0)
        self._plan_cache = {}

    def _stage(self, planes: int, n_blocks: int, stride: int) -> nn.Sequential:
        shortcut = None
        if stride != 1 or self.inplanes != planes:
            # built before the block, like the reference (RNG order), registered after conv/bn of the block
            shortcut = nn.Sequential(nn.Conv2d(self.inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
        blocks: List[nn.Module] = [BasicBlock(self.inplanes, planes, stride, shortcut)]
        self.inplanes = planes
        for _ in range(1, n_blocks):
            blocks.append(BasicBlock(planes, planes))
        return nn.Sequential(*blocks)

    def get_embedding_size(self) -> int:
        return self.hidden_dim

    def blocks(self) -> List[BasicBlock]:
        return [b for stage in (self.layer1, self.layer2, self.layer3, self.layer4) for b in stage]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Inference/feature forward of a stand-alone encoder: [B,H,W] or [B,1,H,W] fp32 -> [B, hidden_dim] fp32.

        Uses BatchNorm batch statistics in train() mode and running statistics in eval() mode, like the reference;
        no autograd graph is built (training goes through ``AVMNIST.train_step``).
        """
        from .engine import StandaloneEncoder

        if x.dim() == 4:
            if x.shape[1] != 1:
                raise ValueError("expected a 1-channel input")
            x = x[:, 0]
        if not x.is_cuda:
            raise RuntimeError("mml_b200 encoders run on a B200 GPU only (no CPU fallback)")
        runner = self._plan_cache.get("standalone")
        if runner is None:
            runner = self._plan_cache["standalone"] = StandaloneEncoder(self)
        return runner.forward(x.float().contiguous(), self.training)


def ResNet18(in_channels: int = 1, hidden_dim: int = 128) -> ResNetEncoder:
    return ResNetEncoder(BasicBlock, (2, 2, 2, 2), in_channels=in_channels, hidden_dim=hidden_dim)


def ResNet34(in_channels: int = 1, hidden_dim: int = 128) -> ResNetEncoder:
    return ResNetEncoder(BasicBlock, (3, 4, 6, 3), in_channels=in_channels, hidden_dim=hidden_dim)
