"""The C-ABI library loads and exports every symbol include/ffc_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ffc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ffc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = header_symbols()
    for s in ("ffc_rfft2", "ffc_irfft2", "ffc_conv2d_fwd", "ffc_conv2d_fwd_ws", "ffc_conv2d_workspace_bytes", "ffc_conv2d_wgrad", "ffc_bn_act_fwd", "ffc_bn_act_bwd",
              "ffc_se_fwd", "ffc_se_bwd", "ffc_bias_grad", "ffc_version", "ffc_last_error", "ffc_workspace_bytes"):
        assert s in syms


def test_cuda_library_builds_loads_and_exports_header_symbols():
    from fastfourierconvolution_b200 import build
    path = build.build()                     # nvcc cross-compiles sm_100a without a GPU
    lib = ctypes.CDLL(path)
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in include/ffc_b200.h but not exported by {path}"
    lib.ffc_version.restype = ctypes.c_int
    assert lib.ffc_version() >= 100
    assert lib.ffc_is_emulation() == 0


def test_binding_signatures_cover_header():
    from fastfourierconvolution_b200 import _C
    assert set(header_symbols()) <= set(_C._SIGNATURES), set(header_symbols()) - set(_C._SIGNATURES)


def test_bad_arguments_return_error_codes_not_crashes():
    """Argument validation happens on the host before any launch, so it can be exercised without a GPU."""
    from fastfourierconvolution_b200 import _C
    L = _C.Library(_C.LIB_PATH)
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert L.ffc_rfft2(p, p, 1, 6, 200, 0, None) == 1         # planes beyond 128 are not supported (6x6 is, on the direct-DFT kernels)
    assert b"unsupported plane" in L.ffc_last_error()
    assert L.ffc_fft2_supported(6, 6) == 1 and L.ffc_fft2_supported(32, 32) == 2 and L.ffc_fft2_supported(6, 200) == 0
    assert L.ffc_rfft2(None, p, 1, 8, 8, 0, None) == 1
    assert L.ffc_conv2d_fwd(p, p, 4, None, None, 0, None, None, p, 1, 4, 8, 8, 9, 9, 3, 1, 1, 0, None) == 1
    assert L.ffc_bn_act_fwd(p, p, p, p, None, None, p, p, 1, 4, 16, 1, 1, 1e-5, 0.1, 3, 0.1, None, 0, None) == 1
