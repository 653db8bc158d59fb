"""FFC_BN_ACT -- drop-in for layers/ffc/ffc_bn_act.py:11-83 (the block the models instantiate)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from . import _util
from .ffc import FFC, FFCTranspose


class FFC_BN_ACT(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, ratio_gin, ratio_gout,
                 stride=1, padding=0, dilation=1, groups=1, bias=False,
                 norm_layer: nn.Module = nn.Identity, activation_layer: nn.Module = nn.Identity,
                 enable_lfu=True, upsampling=False, out_padding=0,
                 uses_noise: bool = False, uses_sn: bool = False, num_classes: int = 1):
        super().__init__()
        self.uses_sn = uses_sn            # stored and ignored, as in the reference (:39)
        if upsampling:
            self.ffc = FFCTranspose(in_channels, out_channels, kernel_size, ratio_gin, ratio_gout, stride, padding,
                                    dilation, groups, bias, enable_lfu, out_padding=out_padding, num_classes=num_classes)
        else:
            self.ffc = FFC(in_channels, out_channels, kernel_size, ratio_gin, ratio_gout, stride, padding,
                           dilation, groups, bias, enable_lfu, num_classes=num_classes)
        # channel counts of the two norms follow the reference's own arithmetic (:49-50)
        ch_l = int(out_channels * (1 - ratio_gout))
        ch_g = int(out_channels * ratio_gout)
        norm_l = nn.Identity if ratio_gout == 1 else norm_layer
        norm_g = nn.Identity if ratio_gout == 0 else norm_layer
        if num_classes > 1:
            self.bn_l, self.bn_g = norm_l(ch_l, num_classes), norm_g(ch_g, num_classes)
        else:
            self.bn_l, self.bn_g = norm_l(ch_l), norm_g(ch_g)
        act_l = nn.Identity if ratio_gout == 1 else activation_layer
        act_g = nn.Identity if ratio_gout == 0 else activation_layer
        # LeakyReLU is always built with slope 0.1 (:66-67)
        self.act_l = act_l(0.1, inplace=True) if isinstance(act_l(), nn.LeakyReLU) else act_l()
        self.act_g = act_g(0.1, inplace=True) if isinstance(act_g(), nn.LeakyReLU) else act_g()

    @staticmethod
    def _norm_act(x, bn, act, y):
        if not torch.is_tensor(x):
            # an empty branch is the int 0 and both modules are Identity (:52-53, :63-64)
            return act(bn(x))
        fused_norm = isinstance(bn, (nn.Identity, nn.BatchNorm2d))
        code = _util.act_code(act)
        if y is not None and not isinstance(bn, nn.Identity):
            if fused_norm:
                raise TypeError("BatchNorm2d.forward() takes one input but a class label y was given "
                                "(the reference raises here as well)")
            x = bn(x, y)                                  # conditional norm stays the caller's module
            return _util.bn_act(x, nn.Identity(), code) if code else act(x)
        if not fused_norm:
            x = bn(x)
            return _util.bn_act(x, nn.Identity(), code) if code else act(x)
        if code is None:                                  # activation without a fused kernel
            return act(_util.bn_act(x, bn, (ops.ACT_IDENTITY, 0.0)))
        return _util.bn_act(x, bn, code)

    def forward(self, x, y=None):
        x_l, x_g = self.ffc(x, y)
        x_l = self._norm_act(x_l, self.bn_l, self.act_l, y)      # :73-76
        x_g = self._norm_act(x_g, self.bn_g, self.act_g, y)      # :78-81
        return x_l, x_g
