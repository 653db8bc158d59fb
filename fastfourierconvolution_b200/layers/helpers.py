"""The small glue modules the reference's models pick up through ``from layers import *``
(layers/resizer.py, layers/noise_injection.py, layers/print_layer.py, layers/gaussian_noise.py).
NoiseInjection runs on the library's own kernel; the others are shape plumbing."""
from __future__ import annotations

import torch
import torch.nn as nn


class Print(nn.Module):
    """Shape-printing pass-through (layers/print_layer.py:15-32); silent unless debug=True."""

    def __init__(self, debug=False):
        super().__init__()
        self.debug = debug

    def forward(self, x):
        if self.debug:
            print(x[0].shape if type(x) is tuple else x.shape)
        return x


def debug_print(*args, **kwargs):
    """layers/print_layer.py:10-12 prints only when Config.shared().DEBUG is set; never here."""
    return None


class Resizer(nn.Module):
    """(x_l, x_g) -> tensor: channel concat, or x_l alone when the global branch is the int 0
    (layers/resizer.py:15-24)."""

    def __init__(self, debug=False):
        super().__init__()
        self.print_size = Print(debug=debug)

    def forward(self, x):
        if type(x) == tuple:
            if type(x[1]) == int:
                return x[0]
            return self.print_size(torch.cat(list(x), dim=1))
        return x


class NoiseInjection(nn.Module):
    """x + weight * N(0,1), one noise plane per image (layers/noise_injection.py:20-32)."""

    def __init__(self, channels):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1, channels, 1, 1))

    def forward(self, x, noise=None):
        if noise is None:
            b, _, h, w = x.shape
            noise = x.new_empty(b, 1, h, w).normal_()
        from .. import ops
        return ops.noise_add(x, self.weight, noise)           # x + weight * noise: one kernel (csrc/ffc_glue.cu)


class GaussianNoise(nn.Module):
    """Additive N(0, std) noise in training mode (layers/gaussian_noise.py)."""

    def __init__(self, stddev):
        super().__init__()
        self.stddev = stddev

    def forward(self, x):
        if self.training:
            return x + torch.empty_like(x).normal_(0, self.stddev)
        return x
