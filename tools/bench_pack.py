"""Times the weight-packing launch of the tcgen05 convolutions (pack_v5_kernel) on the weight shapes of the fgan32 step:
ncu --metrics gpu__time_duration.sum -k regex:pack_v5 python tools/bench_pack.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fastfourierconvolution_b200 import ops

dev = "cuda:0"
torch.manual_seed(0)
with torch.no_grad():
    for (cin, cout, k, s, p, tr, hw) in [(512, 256, 4, 2, 1, True, 4), (256, 128, 4, 2, 1, True, 8), (64, 64, 4, 2, 1, False, 32), (128, 128, 4, 2, 1, False, 16),
                                          (256, 256, 4, 2, 1, False, 8), (256, 512, 3, 1, 1, False, 4), (64, 128, 3, 1, 1, False, 16)]:
        x = torch.randn(32, cin, hw, hw, device=dev)
        w = torch.randn((cin, cout, k, k) if tr else (cout, cin, k, k), device=dev)
        for _ in range(2):
            ops.conv2d(x, w, stride=s, pad=p, transposed=tr)
    torch.cuda.synchronize()
print("ok")
