"""A/B of the small-plane fused Fourier unit forward: warp-private kernel (csrc/ffc_fu4.cu) vs the phase kernel
(csrc/ffc_fu2.cu) on FourierUnitSN(C,C)@32x32 -- accuracy against a float64 evaluation of the same formula on the GPU and
device time per call (CUDA graph replay over rotating inputs larger than L2).
usage: python tools/bench_fu4.py [--json out.jsonl]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import fastfourierconvolution_b200 as ffc
from fastfourierconvolution_b200 import _C
from bench_fu import timeit, ref_fourier_unit, PEAK

DEV = "cuda:0"


def main():
    out = None
    if "--json" in sys.argv:
        out = open(sys.argv[sys.argv.index("--json") + 1], "w")
    torch.manual_seed(0)
    L = _C.lib()
    for (C, B, training) in [(8, 256, True), (8, 256, False), (8, 2048, False), (8, 128, True), (8, 32, True), (8, 512, True), (4, 64, True)]:
        mod = ffc.FourierUnitSN(C, C).to(DEV).train(training)
        mod.fused = "single"
        with torch.no_grad():
            mod.bn.weight.uniform_(0.5, 1.5); mod.bn.bias.uniform_(-0.5, 0.5)
            mod.bn.running_mean.uniform_(-0.2, 0.2); mod.bn.running_var.uniform_(0.5, 1.5)
        nbuf = max(2, int(300e6 // (B * C * 32 * 32 * 4 * 2)) + 1)
        xs = [torch.randn(B, C, 32, 32, device=DEV) for _ in range(min(nbuf, 24))]
        P64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in mod.state_dict().items()}
        ref = ref_fourier_unit(xs[0].double(), P64, training)
        row = {"C": C, "B": B, "training": training}
        for name, on in (("fu4", 1), ("fu2", 0)):
            L.ffc_debug_fu4(on)
            sd = {k: v.clone() for k, v in mod.state_dict().items()}
            with torch.no_grad():
                y = mod(xs[0])
                err = ((y.double() - ref).abs().max() / ref.abs().max()).item()
                t = timeit(lambda x: mod(x), xs, iters=20)
                mod.load_state_dict(sd)
            if training:
                with torch.enable_grad():
                    xg = [x.clone().requires_grad_(True) for x in xs[:max(2, len(xs) // 3)]]
                    gy = torch.randn_like(xs[0])
                    x64 = xs[0].double().requires_grad_(True)
                    P64g = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in P64.items()}
                    g64 = torch.autograd.grad(ref_fourier_unit(x64, P64g, True), (x64, P64g["conv_layer.weight"], P64g["bn.weight"], P64g["bn.bias"]), gy.double())

                    def fwd_bwd(x):
                        return torch.autograd.grad(mod(x), (x, mod.conv_layer.weight, mod.bn.weight, mod.bn.bias), gy)
                    g = fwd_bwd(xg[0])
                    # relative L2 (a ReLU element on the other side of its kink moves the max norm)
                    row[name + "_gerr"] = [float(((a.double().reshape(b.shape) - b).norm() / b.norm()).item()) for a, b in zip(g, g64)]
                    mod.load_state_dict(sd)
                    tb = timeit(fwd_bwd, xg, iters=20)
                    mod.load_state_dict(sd)
                row[name + "_fwdbwd_us"] = round(tb * 1e3, 2)
                row[name + "_fwdbwd_frac"] = round(20.0 * B * C * 1024 / (tb * 1e-3) / 1e9 / PEAK, 4)
            row[name + "_err"] = err
            row[name + "_us"] = round(t * 1e3, 2)
            row[name + "_frac"] = round(8.0 * B * C * 1024 / (t * 1e-3) / 1e9 / PEAK, 4)
        L.ffc_debug_fu4(1)
        print(json.dumps(row), flush=True)
        if out:
            out.write(json.dumps(row) + "\n")


if __name__ == "__main__":
    with torch.no_grad():
        main()
