"""Convolution micro-benchmark: the local-branch shapes of the BASELINE generators, forward / dgrad / wgrad,
for the three kernel families (0 tensor-core 3xTF32, 1 simple FP32, 2 tuned FP32 SIMT) and PyTorch (cuDNN).
usage: python tools/bench_conv.py"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fastfourierconvolution_b200 import _C, ops

DEV = "cuda:0"


def graph_time(fn, iters=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(4):
                fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * 4)


def main():
    L = _C.lib()
    # (name, B, cin, cout, Hi, k, stride, pad, transposed)
    shapes = [("fgan32 conv2 l2l", 256, 512, 192, 4, 4, 2, 1, True), ("fgan32 conv3 l2l", 256, 192, 96, 8, 4, 2, 1, True),
              ("fgan32 conv4 l2l", 256, 96, 48, 16, 4, 2, 1, True), ("fgan32 conv4 l2g", 256, 96, 16, 16, 4, 2, 1, True),
              ("fgan32 conv5 l2l", 256, 48, 3, 32, 3, 1, 1, False), ("fgan128 conv6 l2l", 64, 64, 64, 64, 4, 2, 1, True),
              ("sngan D main.2 l2l", 256, 96, 192, 16, 4, 2, 1, False), ("ST conv1 1x1", 256, 32, 8, 32, 1, 1, 0, False)]
    for name, B, cin, cout, Hi, k, s, p, tr in shapes:
        x = torch.randn(B, cin, Hi, Hi, device=DEV)
        w = torch.randn(*((cin, cout) if tr else (cout, cin)), k, k, device=DEV) * 0.05
        Ho = ops.conv_out_size(Hi, k, s, p, tr)
        dy = torch.randn(B, cout, Ho, Ho, device=DEV)
        taps = k * k / (s * s) if tr else k * k
        flops = 2.0 * B * Ho * Ho * cout * cin * taps
        row = {"shape": name, "GFLOP": round(flops / 1e9, 2)}
        ref = (F.conv_transpose2d(x.double(), w.double(), None, s, p) if tr else F.conv2d(x.double(), w.double(), None, s, p))
        for mode in (3, 4):
            L.ffc_debug_conv_reference(mode)
            y = ops.conv2d(x, w, stride=s, pad=p, transposed=tr)
            err = ((y.double() - ref).abs().max() / ref.abs().max()).item()
            xg, wg = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
            t_f = graph_time(lambda: ops.conv2d(x, w, stride=s, pad=p, transposed=tr))

            def fwd_bwd():      # forward and backward inside the capture (autograd replays on the capture stream)
                out = ops.conv2d(xg, wg, stride=s, pad=p, transposed=tr)
                return torch.autograd.grad(out, (xg, wg), dy)
            t_b = graph_time(fwd_bwd) - t_f
            row[f"m{mode}"] = f"fwd {1000 * t_f:.0f}us {flops / t_f / 1e9:.1f}TF err {err:.1e} | bwd {1000 * t_b:.0f}us {2 * flops / t_b / 1e9:.1f}TF"
        L.ffc_debug_conv_reference(5)
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            fn = (lambda: F.conv_transpose2d(x, w, None, s, p)) if tr else (lambda: F.conv2d(x, w, None, s, p))
            t = graph_time(fn)
            row["cudnn_tf32" if tf32 else "cudnn_fp32"] = f"fwd {1000 * t:.0f}us {flops / t / 1e9:.1f}TF"
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
