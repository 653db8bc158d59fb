"""FourierUnitSN -- drop-in for the reference's layers/ffc/fourier_unity.py:17-58.

Same constructor, same submodules (``conv_layer``, ``bn``, ``relu``; they only hold parameters and
buffers so ``state_dict`` / ``apply(weights_init)`` / optimizers see the reference layout), same
``forward(x, y=None)``.  The forward runs hand-written sm_100a kernels: a fused shared-memory
rfft2 -> channel mix -> BatchNorm + ReLU -> irfft2 kernel where one image's spectrum fits in a
CTA's shared memory, otherwise the general form rfft2 | 1x1 mix | BN statistics | (BN+ReLU)->irfft2 with the
spectrum staged through L2 in the reference's (B, 2C, H, W/2+1) channel layout (no layout copies).  Planes that are
not a square power of two (odd, non-square, the 48x48 of the mg = 6 scripts) take the general form on direct-DFT
plane kernels (csrc/ffc_dft2.cu), any H, W up to 128 -- the reference accepts every size, and so does this class.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from . import _util


def _prefers_staged(c: int, n: int, needs_grad: bool) -> bool:
    """Where both forms apply, which one is faster (measured, profiles/r03a_fu_paths.jsonl: batch 32..256 on one B200): the
    single-kernel form wins while an image is small (8 channels, 16 channels at 16x16); from 24 channels -- or 16 at 32x32 --
    the backward of the L2-staged form (tensor-core mix and weight gradient) is 1.2-1.8x faster, and from 24 channels at
    32x32 its forward is, too."""
    if needs_grad:
        return c >= 24 or (c >= 16 and n >= 32)
    return c >= 24 and n >= 32


class FourierUnitSN(nn.Module):
    def __init__(self, in_channels, out_channels, groups: int = 1, num_classes: int = 1):
        super().__init__()
        if groups != 1:
            raise NotImplementedError("FourierUnitSN: groups != 1 is not supported by the sm_100a kernels")
        self.groups = groups
        # parameter holders, reference layout: conv_layer.weight (2*Cout, 2*Cin, 1, 1), bn.* over 2*Cout
        self.conv_layer = nn.Conv2d(in_channels * 2, out_channels * 2, kernel_size=1, stride=1, padding=0,
                                    groups=groups, bias=False)
        self.bn = nn.BatchNorm2d(out_channels * 2)
        self.relu = nn.ReLU(inplace=True)
        # True: the single-kernel fused form where one image's spectrum fits a CTA, else the L2-staged form (csrc/ffc_fu3.cu);
        # (whichever is faster where both apply); "single" / "staged" force one of them, False the first-generation general form
        # (tests and benchmarks)
        self.fused = True

    def forward(self, x, y=None):
        return self._run(x, y, None)

    def _run(self, x, y, residual):
        """fu(x) [+ residual]; the residual add of spectral_transform.py:108 rides in the irfft2 epilogue."""
        if y is not None:
            raise NotImplementedError("class-conditional FourierUnitSN (y is not None) crashes in the reference "
                                      "(fourier_unity.py:46-47) and is not implemented")
        if x.dim() != 4:
            raise ValueError("FourierUnitSN expects (B, C, H, W)")
        h, w = x.shape[-2:]
        kind = ops.fft2_supported(h, w)       # 2: tuned power-of-two planes (every BASELINE config), 1: any plane up to 128x128
        if not kind:
            raise NotImplementedError(f"FourierUnitSN: planes up to 128x128 are supported, got {h}x{w}")
        weight = _util.effective_weight(self.conv_layer)
        bn = self.bn
        cin, cout = x.shape[1], weight.shape[0] // 2
        plain_bn = bn.affine and bn.track_running_stats and bn.momentum is not None
        single = kind == 2 and plain_bn and self.fused in (True, "single") and ops.fu_fused_supported(x.shape[0], cin, cout, h, w)
        if single and self.fused is True and _prefers_staged(max(cin, cout), h, torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad)) \
                and ops.fu_staged_supported(x.shape[0], cin, cout, h, w):
            single = False
        staged = kind == 2 and plain_bn and self.fused and not single and ops.fu_staged_supported(x.shape[0], cin, cout, h, w)
        if single or staged:
            _util.count_batch(bn)
            return ops.fourier_unit_fused(x, weight.view(2 * cout, 2 * cin), bn.weight, bn.bias, bn.running_mean,
                                          bn.running_var, residual, bn.training, bn.eps, bn.momentum, staged=bool(staged))
        spec = ops.rfft2(x)                                            # fourier_unity.py:38-42
        mixed = ops.conv2d(spec, weight)                               # :45
        if plain_bn:                                                   # :49 applied inside the load of :51-56
            _util.count_batch(bn)
            return ops.bn_relu_irfft2(mixed, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual,
                                      bn.training, bn.eps, bn.momentum, width=w)
        act = _util.bn_act(mixed, self.bn, (ops.ACT_RELU, 0.0))        # :49
        return ops.irfft2(act, residual, width=w)                      # :51-56
