"""Runs one SN-discriminator stage (conv + LeakyReLU, forward + backward) a few times, for ncu captures.
usage: python tools/run_dconv_once.py [shape]   shape in {d2, d3, d4, d5, d6, d7} (fgan32 discriminator, batch 256)"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fastfourierconvolution_b200 import ops
shape = sys.argv[1] if len(sys.argv) > 1 else "d3"
B = 256
cin, cout, Hi, k, s = {"d2": (64, 64, 32, 4, 2), "d3": (64, 128, 16, 3, 1), "d4": (128, 128, 16, 4, 2), "d5": (128, 256, 8, 3, 1),
                       "d6": (256, 256, 8, 4, 2), "d7": (256, 512, 4, 3, 1)}[shape]
x = torch.randn(B, cin, Hi, Hi, device="cuda:0", requires_grad=True)
w = (torch.randn(cout, cin, k, k, device="cuda:0") * 0.05).requires_grad_(True)
b = torch.zeros(cout, device="cuda:0", requires_grad=True)
for _ in range(3):
    y = ops.conv2d_act(x, w, b, s, 1)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize(); print("ok", y.shape)
