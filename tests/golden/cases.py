"""The golden cases, shared by make_golden.py (run on the reference's ``layers``), the oracle tests
(run on oracle/ffc_ref.py) and the parity tests (run on fastfourierconvolution_b200.layers).

CASES[name] = (ctor(L) -> module, input shapes, training, oracle(R, P, xs, training) -> outputs)
where L is a namespace with the reference's class names and R is the oracle module.
"""
import torch.nn as nn

_BN_GELU = dict(norm_layer=nn.BatchNorm2d, activation_layer=nn.GELU)
_LRELU = dict(bias=True, activation_layer=nn.LeakyReLU)


def _cfg(R, *a, **k):
    return R.FFCConfig(*a, **k)


def _tup(xs):
    return xs[0] if len(xs) == 1 else tuple(xs)


def _bnact(R, cfg):
    return lambda P, xs, tr: R.ffc_bn_act(_tup(xs), P, "", cfg(R), tr)


CASES = {
    # row a2/a3
    "fu_c4_16": (lambda L: L.FourierUnitSN(4, 4), [(2, 4, 16, 16)], True,
                 lambda R: lambda P, xs, tr: R.fourier_unit(xs[0], P, "", tr)),
    "fu_c6to4_8_eval": (lambda L: L.FourierUnitSN(6, 4), [(3, 6, 8, 8)], False,
                        lambda R: lambda P, xs, tr: R.fourier_unit(xs[0], P, "", tr)),
    "fu_c2_64": (lambda L: L.FourierUnitSN(2, 2), [(1, 2, 64, 64)], True,
                 lambda R: lambda P, xs, tr: R.fourier_unit(xs[0], P, "", tr)),
    # row a4-a6
    "se_c32": (lambda L: L.SELayer(32), [(2, 32, 4, 4)], True,
               lambda R: lambda P, xs, tr: R.se_layer(xs[0], P, "")),
    "st_s1": (lambda L: L.SpectralTransform(32, 16), [(2, 32, 8, 8)], True,
              lambda R: lambda P, xs, tr: R.spectral_transform(xs[0], P, "", 1, False, tr)),
    "st_up": (lambda L: L.SpectralTransform(32, 16, 2, 1, True, True), [(2, 32, 4, 4)], True,
              lambda R: lambda P, xs, tr: R.spectral_transform(xs[0], P, "", 2, True, tr)),
    "st_down": (lambda L: L.SpectralTransform(16, 32, 2, 1, True, False), [(2, 16, 16, 16)], True,
                lambda R: lambda P, xs, tr: R.spectral_transform(xs[0], P, "", 2, False, tr)),
    # row a7-a9
    "ffcbn_up": (lambda L: L.FFC_BN_ACT(32, 16, 4, .25, .25, 2, 1, upsampling=True, **_BN_GELU),
                 [(2, 24, 8, 8), (2, 8, 8, 8)], True,
                 lambda R: _bnact(R, lambda R: _cfg(R, 32, 16, 4, .25, .25, 2, 1, norm="bn", act="gelu", upsampling=True))),
    "ffcbn_up_first": (lambda L: L.FFC_BN_ACT(40, 32, 4, 0.0, .25, stride=2, padding=1, upsampling=True, **_BN_GELU),
                       [(2, 40, 4, 4)], True,
                       lambda R: _bnact(R, lambda R: _cfg(R, 40, 32, 4, 0.0, .25, 2, 1, norm="bn", act="gelu", upsampling=True))),
    "ffcbn_gen0": (lambda L: L.FFC_BN_ACT(10, 32, 4, 0, .5, 1, 0, activation_layer=nn.LeakyReLU, upsampling=True),
                   [(3, 10, 1, 1)], True,
                   lambda R: _bnact(R, lambda R: _cfg(R, 10, 32, 4, 0, .5, 1, 0, act="leaky_relu", upsampling=True))),
    "ffcbn_last": (lambda L: L.FFC_BN_ACT(32, 3, 3, .25, 0.0, stride=1, padding=1, activation_layer=nn.Tanh),
                   [(2, 24, 8, 8), (2, 8, 8, 8)], True,
                   lambda R: _bnact(R, lambda R: _cfg(R, 32, 3, 3, .25, 0.0, 1, 1, act="tanh"))),
    "ffcbn_down": (lambda L: L.FFC_BN_ACT(32, 64, 4, .25, .25, 2, 1, norm_layer=nn.BatchNorm2d, **_LRELU),
                   [(2, 24, 16, 16), (2, 8, 16, 16)], True,
                   lambda R: _bnact(R, lambda R: _cfg(R, 32, 64, 4, .25, .25, 2, 1, bias=True, norm="bn", act="leaky_relu"))),
    "ffcbn_d0": (lambda L: L.FFC_BN_ACT(3, 32, 3, 0.0, .25, 1, 1, norm_layer=nn.Identity, **_LRELU),
                 [(2, 3, 8, 8)], True,
                 lambda R: _bnact(R, lambda R: _cfg(R, 3, 32, 3, 0.0, .25, 1, 1, bias=True, act="leaky_relu"))),
    "ffcbn_dlast": (lambda L: L.FFC_BN_ACT(32, 1, 4, .5, 0, 1, 0, activation_layer=nn.Sigmoid),
                    [(2, 16, 4, 4), (2, 16, 4, 4)], True,
                    lambda R: _bnact(R, lambda R: _cfg(R, 32, 1, 4, .5, 0, 1, 0, act="sigmoid"))),
    "ffcbn_up_eval": (lambda L: L.FFC_BN_ACT(32, 16, 4, .25, .25, 2, 1, upsampling=True, **_BN_GELU),
                      [(2, 24, 4, 4), (2, 8, 4, 4)], False,
                      lambda R: _bnact(R, lambda R: _cfg(R, 32, 16, 4, .25, .25, 2, 1, norm="bn", act="gelu", upsampling=True))),
    "ffc_plain": (lambda L: L.FFC(16, 16, 3, .5, .5, 1, 1), [(2, 8, 8, 8), (2, 8, 8, 8)], True,
                  lambda R: lambda P, xs, tr: R.ffc(_tup(xs), P, "", _cfg(R, 16, 16, 3, .5, .5, 1, 1), tr)),
    # row a10
    "snffc": (lambda L: L.SNFFC(32, 64, 4, .25, .25, 2, 1), [(2, 24, 8, 8), (2, 8, 8, 8)], True,
              lambda R: lambda P, xs, tr: R.ffc(_tup(xs), P, "", _cfg(R, 32, 64, 4, .25, .25, 2, 1, spectral_norm=True), tr)),
    "snffc_eval": (lambda L: L.SNFFC(32, 32, 3, .25, .25, 1, 1, bias=True), [(2, 24, 8, 8), (2, 8, 8, 8)], False,
                   lambda R: lambda P, xs, tr: R.ffc(_tup(xs), P, "", _cfg(R, 32, 32, 3, .25, .25, 1, 1, bias=True, spectral_norm=True), tr)),
    # planes that are not a square power of two (the reference accepts every size; here: direct-DFT kernels)
    "fu_c4to6_7x9": (lambda L: L.FourierUnitSN(4, 6), [(2, 4, 7, 9)], True,
                     lambda R: lambda P, xs, tr: R.fourier_unit(xs[0], P, "", tr)),
    "st_12": (lambda L: L.SpectralTransform(16, 16), [(2, 16, 12, 12)], True,
              lambda R: lambda P, xs, tr: R.spectral_transform(xs[0], P, "", 1, False, tr)),
    "ffcbn_up_6to12": (lambda L: L.FFC_BN_ACT(32, 16, 4, .25, .25, 2, 1, upsampling=True, **_BN_GELU),
                       [(2, 24, 6, 6), (2, 8, 6, 6)], True,
                       lambda R: _bnact(R, lambda R: _cfg(R, 32, 16, 4, .25, .25, 2, 1, norm="bn", act="gelu", upsampling=True))),
}
