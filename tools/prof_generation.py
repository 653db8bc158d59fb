"""Launches of one generation batch (G.eval() forward to uint8, fgan_complete.py:413-427 path) for an ncu launch list, and the
graph-replay time of the same batch: python tools/prof_generation.py [batch] [variant]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fastfourierconvolution_b200 import harness as H

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
variant = sys.argv[2] if len(sys.argv) > 2 else "fgan32"
dev = torch.device("cuda:0")
torch.manual_seed(0)
G = H.FGenerator(128, 4, variant).to(dev); G.apply(H.weights_init); G.eval()
z = torch.randn(B, 128, device=dev)
with torch.no_grad():
    for _ in range(3):
        out = G(z)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = G(z)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
print(f"generation {variant} batch {B}: {ms * 1000:.1f} us per batch, {B / ms * 1000:.0f} img/s, out {tuple(out.shape)} {out.dtype}")
