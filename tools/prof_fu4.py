"""Two launches of the small-plane fused Fourier unit for an ncu capture: FourierUnitSN(8,8)@32x32, training at batch 256
(cooperative forward kernel + backward kernel) and eval at batch 2048.  usage: ncu --set full -k regex:fu4 ... python tools/prof_fu4.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import fastfourierconvolution_b200 as ffc

torch.manual_seed(0)
for B, training in ((256, True), (2048, False)):
    mod = ffc.FourierUnitSN(8, 8).to("cuda:0").train(training)
    mod.fused = "single"
    x = torch.randn(B, 8, 32, 32, device="cuda:0", requires_grad=training)
    for _ in range(3):
        if training:                      # forward (cooperative fu4_kernel<2>) + backward (fu4_bwd_kernel)
            mod(x).square().sum().backward()
        else:
            with torch.no_grad():
                mod(x)
    torch.cuda.synchronize()
print("ok")
