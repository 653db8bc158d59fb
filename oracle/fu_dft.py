"""Independent explicit-DFT restatement of FourierUnitSN -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Follows layers/ffc/fourier_unity.py:32-58 of the reference without using any FFT library: dense
DFT matrices in float64 (numpy only).  It pins down the conventions that the CUDA kernels must
reproduce and that torch.fft hides:

  * :38  rfftn(norm="ortho"):  S[u,v] = (HW)^-1/2 sum_{h,w} x[h,w] e^{-2 pi i (uh/H + vw/W)},  v < W/2+1
  * :40-42  real channel 2c = Re S_c, 2c+1 = Im S_c
  * :45  Y = Wmix . S per bin (Wmix = conv_layer.weight[:, :, 0, 0], real 2Cout x 2Cin)
  * :49  BatchNorm2d over (b,u,v) with biased variance, eps 1e-5, affine; ReLU
  * :56  irfftn(s=(H,W), norm="ortho") == complex inverse DFT along H, then a one-sided inverse along W
         that takes Re only, counts interior columns twice and therefore ignores Im of columns 0 and W/2.

Only tests/ may import this module.  Pinned against the reference's outputs by
tests/test_oracle_golden.py (fixtures tests/golden/fu_*.npz).
"""
import numpy as np


def _dft_mats(h, w):
    wf = w // 2 + 1
    fh = np.exp(-2j * np.pi * np.outer(np.arange(h), np.arange(h)) / h)          # [u, h]
    fw = np.exp(-2j * np.pi * np.outer(np.arange(w), np.arange(wf)) / w)         # [w, v]
    return fh, fw


def rfft2_ortho(x):
    """x (..., H, W) real -> (..., H, Wf) complex."""
    h, w = x.shape[-2:]
    fh, fw = _dft_mats(h, w)
    return np.einsum("uh,...hw,wv->...uv", fh, x.astype(np.float64), fw) / np.sqrt(h * w)


def irfft2_ortho(z, w):
    """z (..., H, Wf) complex (not necessarily Hermitian) -> (..., H, W) real, c2r semantics."""
    h, wf = z.shape[-2:]
    fh, fw = _dft_mats(h, w)
    t = np.einsum("hu,...uv->...hv", np.conj(fh), z)   # inverse along H
    alpha = np.full(wf, 2.0)
    alpha[0] = 1.0
    if w % 2 == 0:
        alpha[-1] = 1.0
    out = np.einsum("...hv,v,wv->...hw", t, alpha, np.conj(fw)).real
    return out / np.sqrt(h * w)


def fourier_unit(x, wmix, gamma, beta, running_mean=None, running_var=None, training=True, eps=1e-5):
    """x (B,Cin,H,W); wmix (2Cout,2Cin); returns (B,Cout,H,W) float64."""
    b, c, h, w = x.shape
    s = rfft2_ortho(x)                                                   # (B,C,H,Wf) complex
    sr = np.stack([s.real, s.imag], axis=2).reshape(b, 2 * c, h, -1)      # channel 2c+{0,1}
    y = np.einsum("oi,bihw->bohw", wmix.astype(np.float64), sr)
    if training:
        mean = y.mean(axis=(0, 2, 3))
        var = y.var(axis=(0, 2, 3))                                       # biased
    else:
        mean, var = running_mean.astype(np.float64), running_var.astype(np.float64)
    yn = (y - mean[None, :, None, None]) / np.sqrt(var[None, :, None, None] + eps)
    yn = yn * gamma[None, :, None, None] + beta[None, :, None, None]
    r = np.maximum(yn, 0.0)
    r = r.reshape(b, -1, 2, h, r.shape[-1])
    z = r[:, :, 0] + 1j * r[:, :, 1]
    return irfft2_ortho(z, w)
