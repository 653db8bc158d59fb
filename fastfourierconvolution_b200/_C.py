"""ctypes binding of libffc_b200.so (include/ffc_b200.h).

There is exactly one compute backend: the hand-written sm_100a CUDA library.  If it cannot be
loaded, or a tensor is not a CUDA tensor, the call fails loudly -- there is no CPU or library
fallback in the product.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libffc_b200.so")

c_int, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/ffc_b200.h one to one
_SIGNATURES = {
    "ffc_version": (c_int, []),
    "ffc_last_error": (ctypes.c_char_p, []),
    "ffc_is_emulation": (c_int, []),
    "ffc_launch_count": (ctypes.c_ulonglong, []),
    "ffc_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ffc_rfft2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ffc_irfft2": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ffc_conv2d_fwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
                       + [c_int] * 10 + [c_void_p]),
    "ffc_conv2d_fwd_ws": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
                          + [c_int] * 10 + [c_void_p, c_size_t, c_void_p]),
    "ffc_conv2d_workspace_bytes": (c_size_t, [c_int] * 7),
    "ffc_conv2d_block_fwd_ws": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                        c_void_p, c_int, c_void_p, c_int] + [c_int] * 9 + [c_void_p, c_size_t, c_void_p]),
    "ffc_conv2d_act_fwd_ws": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p] + [c_int] * 9 + [c_int, c_float, c_void_p, c_size_t, c_void_p]),
    "ffc_conv2d_wgrad": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 10 + [c_void_p]),
    "ffc_bias_grad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ffc_bn_act_fwd": (c_int, [c_void_p] * 8 + [c_int] * 5 + [c_float, c_float, c_int, c_float, c_void_p, c_size_t, c_void_p]),
    "ffc_bn_act_bwd": (c_int, [c_void_p] * 9 + [c_int] * 6 + [c_float, c_void_p, c_size_t, c_void_p]),
    "ffc_bn_stats": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_float, c_float, c_void_p, c_size_t, c_void_p]),
    "ffc_irfft2_bn_relu": (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_void_p] * 5),
    "ffc_spectral_norm_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ffc_spectral_norm_fwd": (c_int, [c_void_p] * 7 + [c_int, c_int, c_int, c_int, c_float, c_void_p, c_size_t, c_void_p]),
    "ffc_spectral_norm_bwd": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ffc_se_fwd": (c_int, [c_void_p] * 7 + [c_int] * 6 + [c_void_p, c_size_t, c_void_p]),
    "ffc_se_bwd": (c_int, [c_void_p] * 10 + [c_int] * 6 + [c_void_p, c_size_t, c_void_p]),
    "ffc_fu_fwd": (c_int, [c_void_p] * 10 + [c_int] * 6 + [c_float, c_float, c_void_p, c_size_t, c_void_p]),
    "ffc_fu_fused_supported": (c_int, [c_int] * 5),
    "ffc_fu3_fwd": (c_int, [c_void_p] * 10 + [c_int] * 6 + [c_float, c_float, c_void_p, c_size_t, c_void_p]),
    "ffc_fu3_fwd_keep": (c_int, [c_void_p] * 12 + [c_int] * 6 + [c_float, c_float, c_void_p, c_size_t, c_void_p]),
    "ffc_fu3_bwd": (c_int, [c_void_p] * 12 + [c_int] * 6 + [c_void_p, c_size_t, c_void_p]),
    "ffc_fu3_bwd_workspace_bytes": (c_size_t, [c_int] * 5),
    "ffc_fu3_supported": (c_int, [c_int] * 5),
    "ffc_fu3_workspace_bytes": (c_size_t, [c_int] * 6),
    "ffc_debug_fu3_simt_mix": (None, [c_int]),
    "ffc_debug_fu3_chunk_bytes": (None, [c_size_t]),
    "ffc_fu_bwd": (c_int, [c_void_p] * 11 + [c_int] * 6 + [c_void_p, c_size_t, c_void_p]),
    "ffc_fu_bwd_supported": (c_int, [c_int] * 5),
    "ffc_noise_add_fwd": (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p]),
    "ffc_noise_add_bwd_w": (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p]),
    "ffc_gemm_f32": (c_int, [c_void_p] * 4 + [c_int] * 3 + [ctypes.c_longlong] * 6 + [c_void_p]),
    "ffc_colsum_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "ffc_adam_step": (c_int, [c_void_p] * 4 + [ctypes.c_longlong, c_void_p, c_void_p] + [c_float] * 5 + [c_int, c_void_p]),
    "ffc_adam_step_table": (c_int, [c_void_p] * 8 + [c_int, c_int, c_void_p, c_void_p] + [c_float] * 5 + [c_int, c_void_p]),
    "ffc_gather_table": (c_int, [c_void_p] * 6 + [c_int, c_void_p]),
    "ffc_to_uint8": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_float, c_float, c_void_p]),
    "ffc_debug_conv_reference": (None, [c_int]),
    "ffc_debug_fu_two_pass": (None, [c_int]),
    "ffc_debug_fu4": (None, [c_int]),
    "ffc_fft2_supported": (c_int, [c_int, c_int]),
    "ffc_fu_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ffc_fu_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
}
_OPTIONAL = set()


class Library:
    """A loaded libffc_b200 with typed entry points."""

    def __init__(self, path: str):
        self.path = path
        self.cdll = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(self.cdll, name)
            except AttributeError:
                if name in _OPTIONAL:
                    continue
                raise
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def has(self, name: str) -> bool:
        return hasattr(self, name)

    def last_error(self) -> str:
        return self.ffc_last_error().decode("utf-8", "replace")


_lib: Optional[Library] = None
_lock = threading.Lock()


ABI_VERSION = 200          # ffc_version() of the library this binding was written for (csrc/ffc_api.cu)


def lib() -> Library:
    """The CUDA library; (re)built in-tree with nvcc on first use when it is missing or was built from other sources
    than the ones present (content hash, build.is_stale).  A stale library that cannot be rebuilt is an error: the
    ctypes signatures below would otherwise be applied to whatever symbols the old binary has."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                from . import build as _build
                if _build.is_stale():
                    try:
                        _build.build()
                    except Exception as e:  # no nvcc either: nothing can run
                        raise RuntimeError(
                            f"libffc_b200.so is missing or stale ({LIB_PATH}) and could not be built: {e}. "
                            "This package has no CPU or PyTorch fallback.") from e
                L = Library(LIB_PATH)
                if L.ffc_version() != ABI_VERSION:
                    raise RuntimeError(f"{LIB_PATH} reports ABI version {L.ffc_version()}, this binding needs {ABI_VERSION}: "
                                       "rebuild with `python -m fastfourierconvolution_b200.build --force`")
                _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"libffc_b200 error {rc}: {lib().last_error()}")


def require_device(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("fastfourierconvolution_b200 runs on CUDA (sm_100a) tensors only; "
                               "got a tensor on %s (there is no CPU fallback)" % t.device)
        if t is not None and t.dtype != torch.float32:
            raise RuntimeError("fastfourierconvolution_b200 computes in float32; got %s" % t.dtype)


def current_stream(device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t: Optional[torch.Tensor]) -> c_void_p:
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


_workspaces = {}
_retired = []      # buffers a CUDA-graph capture has seen: replays keep using them, so they are never freed


def workspace(nbytes: int, device) -> torch.Tensor:
    """Scratch buffer owned by PyTorch's allocator, one per (device, stream, capturing?), grown on demand.

    Buffers requested while a CUDA graph is being captured live in the graph's private pool and are baked into the graph:
    they get their own cache entries (an eager call never writes into them) and are kept alive for the life of the
    process when a later, larger request replaces them."""
    capturing = torch.cuda.is_current_stream_capturing() if torch.cuda.is_available() else False
    key = (str(device), current_stream(device).value, capturing)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None and capturing:
            _retired.append(ws)
        ws = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


_side_streams = {}


def side_streams(device, n: int = 2):
    """Side streams of ``device`` for work that is independent of what follows on the current stream (weight gradient vs data
    gradient of one convolution, spectral transform vs local convolutions of one FFC layer).  Always used fork / join:
    ``side.wait_stream(cur)`` ... ``cur.wait_stream(side)``, so that a CUDA-graph capture of the caller records parallel
    branches and eager execution overlaps the small kernels of a batch shard.  FFC_B200_SINGLE_STREAM=1 disables the forks."""
    if os.environ.get("FFC_B200_SINGLE_STREAM") == "1":
        return None
    key = str(device)
    st = _side_streams.get(key)
    if st is None or len(st) < n:
        st = [torch.cuda.Stream(device=device) for _ in range(n)]
        _side_streams[key] = st
    return st
