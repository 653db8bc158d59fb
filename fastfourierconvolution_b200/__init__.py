"""fastfourierconvolution_b200 -- the Fast Fourier Convolution layer stack of
phbgomes22/FastFourierConvolution (layers/ffc, layers/snffc) as drop-in torch.nn.Modules running on
hand-written sm_100a CUDA kernels behind a plain C ABI (include/ffc_b200.h).  No CPU fallback."""
from .layers import (FFC, FFC_BN_ACT, FFCTranspose, FourierUnitSN, SELayer, SNFFC, SNFFCTranspose,
                     SpectralTransform, Resizer, NoiseInjection, GaussianNoise, Print, debug_print)

__version__ = "0.1.0"
