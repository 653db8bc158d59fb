"""FFC and FFCTranspose -- drop-ins for layers/ffc/ffc.py:9-99 and layers/ffc/ffc_transpose.py:10-110."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from . import _util
from .spectral_transform import SpectralTransform


def _split(in_channels, out_channels, ratio_gin, ratio_gout):
    in_cg = int(in_channels * ratio_gin)
    out_cg = int(out_channels * ratio_gout)
    return in_channels - in_cg, in_cg, out_channels - out_cg, out_cg


class _FFCBase(nn.Module):
    """Shared forward of FFC / FFCTranspose: both are 'three local convs + spectral transform'."""

    _transposed = False

    def _local_sum(self, pairs, addend_fn=None):
        """sum_i conv_i(x_i) over (holder, input) pairs in ONE kernel launch.

        nn.Identity holders pass their input through (ffc.py:46-47): an int 0 is dropped, a tensor
        is added afterwards, exactly what ``Identity(x) + conv(...)`` does in the reference.
        """
        segs, passthrough = [], []
        for conv, inp in pairs:
            if isinstance(conv, nn.Identity):
                passthrough.append(inp)
            else:
                if not torch.is_tensor(inp):
                    raise TypeError(f"{type(conv).__name__} needs a tensor input, got {inp!r} "
                                    "(the reference raises here as well)")
                segs.append((inp, conv))
        out = 0
        if segs:
            c0 = segs[0][1]
            k, s, p = c0.kernel_size[0], c0.stride[0], c0.padding[0]
            if c0.kernel_size[0] != c0.kernel_size[1] or c0.dilation != (1, 1) or c0.groups != 1:
                raise NotImplementedError("only square kernels, dilation 1, groups 1 are supported")
            op = c0.output_padding[0] if self._transposed else 0
            ws = [_util.effective_weight(c) for _, c in segs]
            biases = [c.bias for _, c in segs if c.bias is not None]
            bias = None
            if biases:
                bias = biases[0] if len(biases) == 1 else biases[0] + biases[1]
            x1, w1 = (segs[1][0], ws[1]) if len(segs) > 1 else (None, None)
            if addend_fn is not None:
                # the spectral transform adds itself onto this result in its last kernel
                base = ops.conv2d(segs[0][0], ws[0], x1, w1, bias, None, s, p, self._transposed, op)
                out = addend_fn(base)
                addend_fn = None
            else:
                out = ops.conv2d(segs[0][0], ws[0], x1, w1, bias, None, s, p, self._transposed, op)
        if addend_fn is not None:
            out = addend_fn(None)
        for t in passthrough:
            if torch.is_tensor(t) or t != 0:
                out = out + t
        return out

    def _local_block(self, x_l, x_g):
        """(convl2l(x_l) + convg2l(x_g), convl2g(x_l)) in ONE launch when all of them are convolutions of one geometry
        (every FFC layer of the reference's generators / discriminators except the first and the last); else None."""
        l2l, l2g, g2l = self.convl2l, self.convl2g, self.convg2l
        if isinstance(l2l, nn.Identity) or isinstance(l2g, nn.Identity) or not torch.is_tensor(x_l):
            return None
        has_g = not isinstance(g2l, nn.Identity)
        if has_g != torch.is_tensor(x_g) or (not has_g and not (type(x_g) is int and x_g == 0)):
            return None
        geo = lambda c: (c.kernel_size, c.stride, c.padding, c.dilation, c.groups, getattr(c, "output_padding", (0, 0)))
        if geo(l2l) != geo(l2g) or (has_g and geo(g2l) != geo(l2l)):
            return None
        if l2l.kernel_size[0] != l2l.kernel_size[1] or l2l.dilation != (1, 1) or l2l.groups != 1:
            raise NotImplementedError("only square kernels, dilation 1, groups 1 are supported")
        w00, w01 = _util.effective_weight(l2l), _util.effective_weight(l2g)
        w10 = _util.effective_weight(g2l) if has_g else None
        bias0 = l2l.bias
        if has_g and g2l.bias is not None:
            bias0 = g2l.bias if bias0 is None else bias0 + g2l.bias
        op = l2l.output_padding[0] if self._transposed else 0
        return ops.conv2d_block(x_l, w00, w01, x_g if has_g else None, w10, bias0, l2g.bias,
                                l2l.stride[0], l2l.padding[0], self._transposed, op)

    def forward(self, x, y=None):
        x_l, x_g = x if type(x) is tuple else (x, 0)
        out_xl, out_xg = 0, 0
        st = self.convg2g if type(self.convg2g) is not nn.Identity else None
        fork = None
        if st is not None and torch.is_tensor(x_g) and x_g.is_cuda and not isinstance(self.convl2l, nn.Identity):
            # The spectral transform up to its last convolution (SE, conv1, BN+ReLU, Fourier unit: small kernels) does not
            # depend on the local convolutions (one large launch): queue it on a side stream, join before conv2 adds the
            # l2g output.  autograd runs each backward node on its forward stream, so the backward forks the same way, and a
            # CUDA-graph capture of the step records the two chains as parallel branches.
            fork = ops._Fork(x_g.device, 1)
            with fork:
                s = st._pre(x_g, y)
        blk = self._local_block(x_l, x_g)
        if fork is not None:
            ops._join(x_g.device)
        if blk is not None:
            out_xl, base = blk
            if st is not None:
                return out_xl, (st._post(s, base) if fork is not None else st._run(x_g, y, base))
            return out_xl, base
        if fork is not None:
            return self._local_sum([(self.convl2l, x_l), (self.convg2l, x_g)]) if self.ratio_gout != 1 else 0, \
                self._local_sum([(self.convl2g, x_l)], addend_fn=lambda base: st._post(s, base))
        if self.ratio_gout != 1:
            out_xl = self._local_sum([(self.convl2l, x_l), (self.convg2l, x_g)])
        if self.ratio_gout != 0:
            if type(self.convg2g) is not nn.Identity:
                out_xg = self._local_sum([(self.convl2g, x_l)],
                                         addend_fn=lambda base: self.convg2g._run(x_g, y, base))
            else:
                out_xg = self._local_sum([(self.convl2g, x_l)])
        return out_xl, out_xg


class FFC(_FFCBase):
    _transposed = False

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int,
                 ratio_gin: float, ratio_gout: float, stride: int = 1, padding: int = 0,
                 dilation: int = 1, groups: int = 1, bias: bool = False, enable_lfu: bool = True,
                 attention: bool = False, num_classes: int = 1):
        super().__init__()
        assert stride == 1 or stride == 2, "Stride should be 1 or 2."
        if groups != 1 or dilation != 1:
            raise NotImplementedError("FFC: groups/dilation other than 1 are not supported by the sm_100a kernels")
        self.stride = stride
        in_cl, in_cg, out_cl, out_cg = _split(in_channels, out_channels, ratio_gin, ratio_gout)
        self.ratio_gin = ratio_gin
        self.ratio_gout = ratio_gout

        def local(cin, cout):
            if cin == 0 or cout == 0:
                return nn.Identity()
            return nn.Conv2d(cin, cout, kernel_size, stride, padding, dilation, groups, bias)

        self.convl2l = local(in_cl, out_cl)
        self.convl2g = local(in_cl, out_cg)
        self.convg2l = local(in_cg, out_cl)
        if in_cg == 0 or out_cg == 0:
            self.convg2g = nn.Identity()
        else:
            self.convg2g = SpectralTransform(in_cg, out_cg, stride, 1, enable_lfu, False, num_classes)


class FFCTranspose(_FFCBase):
    _transposed = True

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int,
                 ratio_gin: float, ratio_gout: float, stride: int = 1, padding: int = 0,
                 dilation: int = 1, groups: int = 1, bias: bool = False,
                 enable_lfu: bool = True, out_padding: int = 0, num_classes: int = 1):
        super().__init__()
        assert stride == 1 or stride == 2, "Stride should be 1 or 2."
        if groups != 1 or dilation != 1:
            raise NotImplementedError("FFCTranspose: groups/dilation other than 1 are not supported by the sm_100a kernels")
        self.stride = stride
        in_cl, in_cg, out_cl, out_cg = _split(in_channels, out_channels, ratio_gin, ratio_gout)
        self.ratio_gin = ratio_gin
        self.ratio_gout = ratio_gout

        def local(cin, cout):
            return self.convtransp2d(cin == 0 or cout == 0, cin, cout, kernel_size, stride, padding,
                                     output_padding=out_padding, groups=groups, bias=bias, dilation=dilation)

        self.convl2l = local(in_cl, out_cl)
        self.convl2g = local(in_cl, out_cg)
        self.convg2l = local(in_cg, out_cl)
        if in_cg == 0 or out_cg == 0:
            self.convg2g = nn.Identity()
        else:
            self.convg2g = SpectralTransform(in_cg, out_cg, stride, 1, enable_lfu, True, num_classes)

    def convtransp2d(self, condition: bool, in_ch: int, out_ch: int, kernel_size: int, stride: int,
                     padding: int, output_padding: int, groups: int, bias: int, dilation: int):
        """Same helper name/signature as ffc_transpose.py:80-86 (subclasses may call it)."""
        if condition:
            return nn.Identity()
        return nn.ConvTranspose2d(in_ch, out_ch, kernel_size, stride, padding, output_padding=output_padding,
                                  groups=groups, bias=bias, dilation=dilation)
