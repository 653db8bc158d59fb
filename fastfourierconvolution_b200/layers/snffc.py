"""SNFFC / SNFFCTranspose -- spectral-norm twins (layers/snffc/snffc.py:12-33, snffc_transpose.py:11-36).

``torch.nn.utils.spectral_norm`` is applied to the same parameter holders the reference wraps, so
``state_dict`` carries ``weight_orig / weight_u / weight_v`` exactly as the reference's does; the
sm_100a forward pulls the normalised weight through ``_util.effective_weight`` (one power
iteration per training forward, sigma = u^T W v, eps 1e-12).
"""
from __future__ import annotations

import torch.nn as nn
from torch.nn.utils import spectral_norm

from .ffc import FFC, FFCTranspose


def _wrap_spectral_transform(st):
    """SN on the SpectralTransform's direct-child convolutions only: conv1 and conv2
    (snffc.py:28-33); fu.conv_layer sits one level deeper and stays un-normalised."""
    if isinstance(st, nn.Identity):
        return
    for name, child in list(st.named_children()):
        if isinstance(child, (nn.Conv2d, nn.ConvTranspose2d)):
            st._modules[name] = spectral_norm(child)


class SNFFC(FFC):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: int,
                 ratio_gin: float, ratio_gout: float, stride: int = 1, padding: int = 0,
                 dilation: int = 1, groups: int = 1, bias: bool = False, enable_lfu: bool = True,
                 attention: bool = False):
        FFC.__init__(self, in_channels, out_channels, kernel_size, ratio_gin, ratio_gout, stride,
                     padding, dilation, groups, bias, enable_lfu, attention)
        # snffc.py:23 wraps convl2l unconditionally (an nn.Identity there raises KeyError('weight'),
        # which torch raises here too); the other two only when they are convolutions (:24-25)
        self.convl2l = spectral_norm(self.convl2l)
        if isinstance(self.convg2l, nn.Conv2d):
            self.convg2l = spectral_norm(self.convg2l)
        if isinstance(self.convl2g, nn.Conv2d):
            self.convl2g = spectral_norm(self.convl2g)
        _wrap_spectral_transform(self.convg2g)


class SNFFCTranspose(FFCTranspose):
    """The reference class cannot be constructed (snffc_transpose.py:28 reads a non-existent
    ``self.convg2gup`` and :18-19 passes ``attention`` into the ``num_classes`` slot).  This is its
    evident intent: FFCTranspose with spectral norm on the three transposed convolutions and on the
    spectral transform's conv1/conv2, i.e. the mirror image of SNFFC."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int,
                 ratio_gin: float, ratio_gout: float, stride: int = 1, padding: int = 0,
                 dilation: int = 1, groups: int = 1, bias: bool = False,
                 enable_lfu: bool = True, out_padding: int = 0, attention: bool = False):
        FFCTranspose.__init__(self, in_channels, out_channels, kernel_size, ratio_gin, ratio_gout, stride,
                              padding, dilation, groups, bias, enable_lfu, out_padding)
        self.convl2l = spectral_norm(self.convl2l)
        if isinstance(self.convg2l, nn.ConvTranspose2d):
            self.convg2l = spectral_norm(self.convg2l)
        if isinstance(self.convl2g, nn.ConvTranspose2d):
            self.convl2g = spectral_norm(self.convl2g)
        _wrap_spectral_transform(self.convg2g)
