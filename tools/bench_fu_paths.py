"""Single-kernel form (csrc/ffc_fu2*.cu) vs L2-staged form (csrc/ffc_fu3*.cu) of the Fourier unit on the shapes both support:
training forward, eval forward, forward + backward.  Feeds the routing rule in layers/fourier_unity.py.
usage: python tools/bench_fu_paths.py [out.jsonl]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import fastfourierconvolution_b200 as ffc
from fastfourierconvolution_b200 import ops
from bench_fu import timeit, timeit_fwd_bwd

DEV = "cuda:0"
SHAPES = [(256, 8, 32), (256, 16, 16), (128, 8, 32), (128, 16, 16), (64, 32, 32), (64, 16, 32), (32, 8, 32), (32, 8, 16), (32, 16, 16), (32, 16, 32),
          (32, 24, 16), (32, 24, 32), (32, 32, 16), (32, 32, 32), (128, 8, 16), (128, 16, 32), (128, 24, 16), (128, 24, 32), (128, 32, 16), (128, 32, 32)]


def main():
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None
    for B, C, N in SHAPES:
        if not (ops.fu_fused_supported(B, C, C, N, N) and ops.fu_staged_supported(B, C, C, N, N)):
            continue
        torch.manual_seed(0)
        m = ffc.FourierUnitSN(C, C).to(DEV)
        nbuf = min(max(2, int(200e6 // (4 * B * C * N * N)) + 1), 8)
        xs = [torch.randn(B, C, N, N, device=DEV, requires_grad=True) for _ in range(nbuf)]
        row = {"B": B, "C": C, "N": N}
        for tag, mode in (("single", "single"), ("staged", "staged")):
            m.fused = mode
            with torch.no_grad():
                m.train()
                row[tag + "_fwd_train_us"] = round(1000 * timeit(m, xs, iters=5), 1)
                m.eval()
                row[tag + "_fwd_eval_us"] = round(1000 * timeit(m, xs, iters=5), 1)
            m.train()
            row[tag + "_fwd_bwd_us"] = round(1000 * timeit_fwd_bwd(m, xs, iters=5), 1)
        print(json.dumps(row), flush=True)
        if out:
            out.write(json.dumps(row) + "\n"); out.flush()


if __name__ == "__main__":
    main()
