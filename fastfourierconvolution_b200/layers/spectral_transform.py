"""SELayer and SpectralTransform -- drop-ins for layers/ffc/spectral_transform.py:12-110."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from . import _util
from .fourier_unity import FourierUnitSN


class SELayer(nn.Module):
    """Squeeze-excite gate (spectral_transform.py:12-28); ``fc`` holds the two bias-free Linears."""

    def __init__(self, channel, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(
            nn.Linear(channel, channel // reduction, bias=False),
            nn.ReLU(inplace=True),
            nn.Linear(channel // reduction, channel, bias=False),
            nn.Sigmoid(),
        )

    def forward(self, x):
        return self._run(x, ops.RESAMPLE_NONE)

    def _run(self, x, mode):
        return ops.se_resample(x, self.fc[0].weight, self.fc[2].weight, mode)


class SpectralTransform(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, stride: int = 1, groups: int = 1,
                 enable_lfu: bool = False, upsample: bool = False, num_classes: int = 1):
        super().__init__()
        if groups != 1:
            raise NotImplementedError("SpectralTransform: groups != 1 is not supported by the sm_100a kernels")
        self.enable_lfu = enable_lfu
        self.stride = stride
        # resampling in front of the transform (reference :42-47); kept as a module for state/API parity
        self._mode = ops.RESAMPLE_NONE
        self.downsample = nn.Identity()
        if stride == 2 and upsample:
            self.downsample = nn.Upsample(scale_factor=2, mode="nearest")
            self._mode = ops.RESAMPLE_UP2
        if stride == 2 and not upsample:
            self.downsample = nn.AvgPool2d(kernel_size=(2, 2), stride=2)
            self._mode = ops.RESAMPLE_AVGPOOL2
        half = out_channels // 2
        self.conv1 = nn.Conv2d(in_channels, half, kernel_size=1, groups=groups, bias=False)
        self.bn1 = nn.BatchNorm2d(half)
        self.act1 = nn.ReLU(inplace=True)
        self.fu = FourierUnitSN(half, half, groups, num_classes=num_classes)
        if self.enable_lfu:
            # constructed but unused, as in the reference (:65-67 vs the commented block :94-105):
            # its parameters exist in every state_dict and never receive gradients
            self.lfu = FourierUnitSN(half, half, groups, num_classes=num_classes)
        self.conv2 = nn.Conv2d(half, out_channels, kernel_size=1, groups=groups, bias=False)
        self.se_block = SELayer(self.conv1.in_channels)

    def forward(self, x, y=None):
        return self._run(x, y, None)

    def _run(self, x, y, addend):
        """conv2(x1 + fu(x1)) [+ addend], x1 = relu(bn1(conv1(se(resample(x)))))."""
        return self._post(self._pre(x, y), addend)

    def _pre(self, x, y):
        """x1 + fu(x1): everything of the transform that does not depend on the local branches of the enclosing FFC."""
        xs = self.se_block._run(x, self._mode)                                       # :79, :87
        c1 = ops.conv2d(xs, _util.effective_weight(self.conv1))                       # :89
        x1 = _util.bn_act(c1, self.bn1, (ops.ACT_RELU, 0.0))
        return self.fu._run(x1, y, x1)                                                # :91 + the add of :108

    def _post(self, s, addend):
        return ops.conv2d(s, _util.effective_weight(self.conv2), addend=addend)       # :108
