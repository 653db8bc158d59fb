"""Weight-gradient micro-benchmark (wgrad_v5_kernel) on the SN-discriminator layer shapes of fgan32, batch 256.
usage: python tools/bench_wgrad.py      (FFC_LIB=<path to an alternative libffc_b200.so> to A/B builds)"""
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
from fastfourierconvolution_b200 import _C
if os.environ.get('FFC_LIB'):
    _C._lib = _C.Library(os.environ['FFC_LIB'])
L = _C.lib()
shapes = {"d2": (64, 64, 32, 4, 2), "d3": (64, 128, 16, 3, 1), "d4": (128, 128, 16, 4, 2), "d5": (128, 256, 8, 3, 1),
          "d6": (256, 256, 8, 4, 2), "d7": (256, 512, 4, 3, 1)}
B = 256
out = {}
for name, (cin, cout, Hi, k, s) in shapes.items():
    Ho = (Hi + 2 - k) // s + 1
    x = torch.randn(B, cin, Hi, Hi, device="cuda")
    dy = torch.randn(B, cout, Ho, Ho, device="cuda")
    dw = torch.empty(cout, cin, k, k, device="cuda")
    st = _C.current_stream(x.device)
    def run():
        _C.check(L.ffc_conv2d_wgrad(_C.ptr(dy), _C.ptr(x), _C.ptr(dw), B, cout, cin, Ho, Ho, Hi, Hi, k, s, 1, st))
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1000
    gf = 2.0 * B * Ho * Ho * cout * cin * k * k / 1e9
    out[name] = f"{us:.0f}us {gf / us * 1e3 / 1e3:.0f}TF"
print(os.environ.get("FFC_LIB", "in-tree"), json.dumps(out))
