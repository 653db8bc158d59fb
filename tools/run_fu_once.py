"""Runs the FourierUnit forward a few times on one shape (for ncu captures).
usage: python tools/run_fu_once.py B C N [train|eval] [general|staged] [bwd] [chunk=<MB>]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import fastfourierconvolution_b200 as ffc

B, C, N = (int(a) for a in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "eval"
m = ffc.FourierUnitSN(C, C).to("cuda:0").train(mode == "train")
m.fused = False if "general" in sys.argv else ("staged" if "staged" in sys.argv else True)
for a in sys.argv:
    if a.startswith("chunk="):
        from fastfourierconvolution_b200 import _C
        _C.lib().ffc_debug_fu3_chunk_bytes(int(a[6:]) << 20)
xs = [torch.randn(B, C, N, N, device="cuda:0") for _ in range(4)]
if "bwd" in sys.argv:
    for i in range(4):
        x = xs[i].requires_grad_(True)
        m(x).square().sum().backward()
else:
    with torch.no_grad():
        for i in range(8):
            m(xs[i % 4])
torch.cuda.synchronize()
print("ok")
