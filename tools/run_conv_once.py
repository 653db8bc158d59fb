"""Runs one convolution shape (forward + backward) a few times, for ncu captures.
usage: python tools/run_conv_once.py [mode] [shape]   shape in {conv2, conv3, conv4, conv6}"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fastfourierconvolution_b200 import _C, ops
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 5
shape = sys.argv[2] if len(sys.argv) > 2 else "conv3"
B, cin, cout, Hi = {"conv2": (256, 512, 192, 4), "conv3": (256, 192, 96, 8), "conv4": (256, 96, 48, 16), "conv6": (64, 64, 64, 64)}[shape]
_C.lib().ffc_debug_conv_reference(mode)
x = torch.randn(B, cin, Hi, Hi, device="cuda:0", requires_grad=True)
w = (torch.randn(cin, cout, 4, 4, device="cuda:0") * 0.05).requires_grad_(True)
for _ in range(3):
    y = ops.conv2d(x, w, stride=2, pad=1, transposed=True)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize(); print("ok", y.shape)
