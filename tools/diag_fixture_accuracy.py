"""Where does the forward error of a whole-model fixture come from?  Runs tests/golden/model_*.npz forward on the GPU under
different kernel families and prints max|out - float64 oracle| / max|oracle| for each.
usage: python tools/diag_fixture_accuracy.py model_fgan128_G fgan128"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import parity
import test_layers_emu as T
import fastfourierconvolution_b200 as ffc
from fastfourierconvolution_b200 import _C

name, fn = sys.argv[1], sys.argv[2]
L = _C.lib()


def run(conv_mode, fu_mode):
    L.ffc_debug_conv_reference(conv_mode)
    orig = ffc.FourierUnitSN.__init__

    def init(self, *a, **k):
        orig(self, *a, **k)
        self.fused = fu_mode
    ffc.FourierUnitSN.__init__ = init
    ffc.layers.fourier_unity.FourierUnitSN.__init__ = init
    try:
        got, fx, oracle_run = T.run_model_fixture(name, fn, "cuda:0")
    finally:
        ffc.FourierUnitSN.__init__ = orig
        L.ffc_debug_conv_reference(5)
    return got, oracle_run


ref = None
for conv_mode, fu_mode, tag in ((5, True, "default (tcgen05 convs, fused/staged FU)"), (5, False, "tcgen05 convs, first-generation general FU"),
                                (3, True, "mma.sync 3xTF32 convs"), (2, True, "FP32 SIMT convs")):
    got, oracle_run = run(conv_mode, fu_mode)
    if ref is None:
        ref = oracle_run({})[0]
        r32 = oracle_run({}, torch.float32)[0]
        print(f"{name}: reference FP32 (CPU) out err {parity.relerr(r32['out0'], ref['out0']):.2e}  din {parity.relerr(r32['din0'], ref['din0']):.2e}")
    print(f"{tag:55s} out err {parity.relerr(got['out0'], ref['out0']):.2e}  din err {parity.relerr(got['din0'], ref['din0']):.2e}", flush=True)
