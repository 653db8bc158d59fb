"""Runs one convolution shape a few times (for ncu captures). usage: python tools/run_conv_once.py [mode]"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fastfourierconvolution_b200 import _C, ops
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_C.lib().ffc_debug_conv_reference(mode)
x = torch.randn(256, 192, 8, 8, device="cuda:0"); w = torch.randn(192, 96, 4, 4, device="cuda:0") * 0.05
for _ in range(4):
    y = ops.conv2d(x, w, stride=2, pad=1, transposed=True)
torch.cuda.synchronize(); print("ok", y.shape)
