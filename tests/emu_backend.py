"""TEST INFRASTRUCTURE: runs the product's Python modules on the host *emulation* build of the
CUDA sources (g++ -DFFC_EMU, see csrc/ffc_common.cuh) so that the autograd wiring and the kernels'
index arithmetic can be checked in a container without a GPU.  The product never imports this; it
works by monkeypatching fastfourierconvolution_b200._C inside a pytest fixture."""
import ctypes
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_LIB = os.path.join(HERE, "emu", "libffc_emu.so")


def build():
    from fastfourierconvolution_b200 import build as b
    return b.build_emulation(EMU_LIB)


class patched:
    """Context manager: route fastfourierconvolution_b200 through the emulation library on CPU tensors."""

    def __enter__(self):
        from fastfourierconvolution_b200 import _C
        self._C = _C
        self.saved = (_C._lib, _C.require_device, _C.current_stream, _C.workspace)
        _C._lib = _C.Library(build())
        assert _C._lib.ffc_is_emulation() == 1

        def require_device(*tensors):
            for t in tensors:
                if t is not None and (t.is_cuda or t.dtype != torch.float32):
                    raise RuntimeError("emulation backend expects CPU float32 tensors")

        _C.require_device = require_device
        _C.current_stream = lambda device: ctypes.c_void_p(0)
        ws = {}

        def workspace(nbytes, device):
            if "t" not in ws or ws["t"].numel() < nbytes:
                ws["t"] = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8)
            return ws["t"]

        _C.workspace = workspace
        return self

    def __exit__(self, *exc):
        _C = self._C
        _C._lib, _C.require_device, _C.current_stream, _C.workspace = self.saved
        return False
