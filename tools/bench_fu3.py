"""FourierUnit forward on the GPU: L2-staged form (csrc/ffc_fu3.cu) vs the first-generation general form, per shape, plus a
sweep of the chunk size (bytes of spectrum per chunk of images).  CUDA events around CUDA-graph replays over rotating inputs
whose total exceeds the 126 MB L2.  usage: python tools/bench_fu3.py [--out file.jsonl] [--quick]"""
import json
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench
import fastfourierconvolution_b200 as ffc
from fastfourierconvolution_b200 import _C, ops

DEV = "cuda:0"
PEAK = bench.measured_peaks()[0]


def time_fu(B, C, N, mode, train, chunk=0, bwd=False):
    torch.manual_seed(0)
    fu = ffc.FourierUnitSN(C, C).to(DEV).train(train)
    fu.fused = mode
    nbuf = min(max(2, int(160e6 // (4 * B * C * N * N)) + 1), 24)
    xs = [torch.randn(B, C, N, N, device=DEV) for _ in range(nbuf)]
    _C.lib().ffc_debug_fu3_chunk_bytes(chunk)
    try:
        if bwd:
            gy = torch.randn(B, C, N, N, device=DEV)
            xg = [x.requires_grad_(True) for x in xs]
            ms = bench._graph_time(lambda x: torch.autograd.grad(fu(x), (x, fu.conv_layer.weight, fu.bn.weight, fu.bn.bias), gy), xg)
        else:
            with torch.no_grad():
                ms = bench._graph_time(lambda x: fu(x), xs)
    finally:
        _C.lib().ffc_debug_fu3_chunk_bytes(0)
    alg = (20.0 if bwd else 8.0) * B * C * N * N
    return {"us": 1000 * ms, "gbs": alg / ms / 1e6, "frac": alg / ms / 1e6 / PEAK}


def main():
    out = None
    if "--out" in sys.argv:
        out = open(sys.argv[sys.argv.index("--out") + 1], "w")
    quick = "--quick" in sys.argv
    shapes = [(64, 32, 128), (64, 32, 64), (128, 8, 64), (64, 64, 16), (32, 32, 128), (32, 32, 64), (32, 64, 32), (32, 64, 64), (32, 16, 128), (32, 8, 128)]
    if quick:
        shapes = shapes[:3]
    for (B, C, N) in shapes:
        row = {"B": B, "C": C, "N": N, "alg_MB_fwd": 8.0 * B * C * N * N / 1e6}
        for train in (True, False):
            tag = "train" if train else "eval"
            row["staged_" + tag] = time_fu(B, C, N, "staged", train)
            row["general_" + tag] = time_fu(B, C, N, False, train)
        row["staged_fwd_bwd"] = time_fu(B, C, N, "staged", True, bwd=True)
        line = json.dumps(row)
        print(line, flush=True)
        if out:
            out.write(line + "\n"); out.flush()
    # chunk-size sweep on the headline shape (fgan128's largest unit)
    B, C, N = 64, 32, 128
    for mb in (6, 12, 24, 48, 96, 4096):
        row = {"B": B, "C": C, "N": N, "chunk_MB": mb, "train": time_fu(B, C, N, "staged", True, chunk=mb << 20),
               "eval": time_fu(B, C, N, "staged", False, chunk=mb << 20)}
        line = json.dumps(row)
        print(line, flush=True)
        if out:
            out.write(line + "\n"); out.flush()


if __name__ == "__main__":
    main()
