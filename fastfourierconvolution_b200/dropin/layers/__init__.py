"""Shadow ``layers`` package: put ``fastfourierconvolution_b200/dropin`` on ``sys.path`` AHEAD of the reference checkout and
the reference's ``models/*`` and ``*_complete.py`` scripts (``from layers import *``, models/ffcmodel.py:8,
sngan_complete.py:5) pick up the sm_100a implementations of the hot path without being edited:

    sys.path[:0] = [".../fastfourierconvolution_b200/dropin", "/path/to/FastFourierConvolution"]
    import models                     # the reference's own model classes, now built from these layers

Exports the names of the reference's ``layers/__init__.py:2-22``.  The hot-path classes (``FFC``, ``FFCTranspose``,
``FFC_BN_ACT``, ``SpectralTransform``, ``FourierUnitSN``, ``SELayer``, ``SNFFC``, ``SNFFCTranspose``) and the glue around them
(``Resizer``, ``Print``, ``debug_print``, ``NoiseInjection``, ``GaussianNoise``) come from ``fastfourierconvolution_b200.layers``;
the helpers outside the hot path (``ConditionalBatchNorm2d``, layers/cond/cond_bn.py; ``aw_method``, layers/aw_loss.py) are
passed through UNCHANGED: they are loaded by file path from the reference checkout found on ``sys.path`` (or under
``$FFC_REFERENCE_ROOT``), never copied.  Without a reference checkout those two names raise on use.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

from fastfourierconvolution_b200.layers import (FFC, FFC_BN_ACT, FFCTranspose, FourierUnitSN, GaussianNoise, NoiseInjection,  # noqa: F401
                                                Print, Resizer, SELayer, SNFFC, SNFFCTranspose, SpectralTransform, debug_print)

_HERE = _os.path.dirname(_os.path.abspath(__file__))


def _reference_layers_dir():
    roots = [_os.environ["FFC_REFERENCE_ROOT"]] if _os.environ.get("FFC_REFERENCE_ROOT") else []
    roots += [p or "." for p in _sys.path]
    for root in roots:
        d = _os.path.join(root, "layers")
        if _os.path.isdir(d) and _os.path.abspath(d) != _HERE and _os.path.exists(_os.path.join(d, "aw_loss.py")):
            return d
    return None


def _load(rel, modname):
    d = _reference_layers_dir()
    if d is None:
        return None
    spec = _ilu.spec_from_file_location("_ffc_ref_passthrough_" + modname, _os.path.join(d, rel))
    mod = _ilu.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _missing(name, rel):
    class _Missing:
        def __init__(self, *a, **k):
            raise ImportError(f"layers.{name} is passed through from the reference checkout (layers/{rel}); put the checkout "
                              "on sys.path after fastfourierconvolution_b200/dropin or set FFC_REFERENCE_ROOT")
    _Missing.__name__ = name
    return _Missing


_m = _load(_os.path.join("cond", "cond_bn.py"), "cond_bn")
ConditionalBatchNorm2d = _m.ConditionalBatchNorm2d if _m else _missing("ConditionalBatchNorm2d", "cond/cond_bn.py")
_m = _load("aw_loss.py", "aw_loss")
aw_method = _m.aw_method if _m else _missing("aw_method", "aw_loss.py")
del _m

__all__ = ["FFC", "FFCTranspose", "FFC_BN_ACT", "SpectralTransform", "FourierUnitSN", "SELayer", "SNFFC", "SNFFCTranspose",
           "Resizer", "Print", "debug_print", "NoiseInjection", "GaussianNoise", "ConditionalBatchNorm2d", "aw_method"]
