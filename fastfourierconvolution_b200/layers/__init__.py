"""Drop-in replacement for the reference's ``layers`` package (layers/__init__.py:2-22): the same
class names, constructor arguments, forward signatures and state_dict layout, running on the
hand-written sm_100a kernels of libffc_b200.so."""
from .fourier_unity import FourierUnitSN
from .spectral_transform import SELayer, SpectralTransform
from .ffc import FFC, FFCTranspose
from .ffc_bn_act import FFC_BN_ACT
from .snffc import SNFFC, SNFFCTranspose
from .helpers import Print, debug_print, Resizer, NoiseInjection, GaussianNoise

__all__ = ["FourierUnitSN", "SELayer", "SpectralTransform", "FFC", "FFCTranspose", "FFC_BN_ACT",
           "SNFFC", "SNFFCTranspose", "Print", "debug_print", "Resizer", "NoiseInjection", "GaussianNoise"]
