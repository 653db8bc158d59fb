"""FourierUnit micro-benchmark on the GPU: fused kernel vs general form vs PyTorch (cuFFT + cuDNN) for the
shapes of the BASELINE configs and the isolated sweep.  CUDA events, rotating inputs larger than L2.
usage: python tools/bench_fu.py [--json out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import fastfourierconvolution_b200 as ffc
from fastfourierconvolution_b200 import _C, ops
from oracle import ffc_ref as R

DEV = "cuda:0"
PEAK = 6545.6
if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]


def timeit(fn, xs, iters=10, warm=2):
    """Device time per call: one pass over the rotating inputs is captured into a CUDA graph and replayed,
    so host launch overhead (tens of us per call from Python) does not hide the kernel time."""
    for i in range(warm):
        fn(xs[i % len(xs)])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        with torch.cuda.graph(g, stream=stream):
            for x in xs:
                fn(x)
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * len(xs))


def torch_fu(mod):
    P = dict(mod.state_dict())
    def f(x, training):
        return R.fourier_unit(x, P, "", training)        # the reference's op sequence on cuFFT/cuDNN
    return f


def main():
    shapes = [(256, 8, 32), (256, 16, 16), (256, 32, 8), (128, 8, 64), (128, 16, 16), (64, 64, 16), (64, 32, 32),
              (64, 32, 64), (64, 32, 128), (32, 32, 16), (32, 16, 32), (32, 8, 32), (32, 24, 16), (32, 96, 32)]
    if "--quick" in sys.argv:
        shapes = [(256, 8, 32), (256, 16, 16), (256, 32, 8), (2048, 8, 32), (2048, 16, 16), (64, 32, 32), (512, 32, 32)]
    rows = []
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for B, C, N in shapes:
        torch.manual_seed(0)
        m = ffc.FourierUnitSN(C, C).to(DEV)
        nbytes = 4 * B * C * N * N
        nbuf = min(max(2, int(200e6 // nbytes) + 1), 16)
        xs = [torch.randn(B, C, N, N, device=DEV) for _ in range(nbuf)]
        alg = 2.0 * nbytes
        row = {"B": B, "C": C, "N": N, "MB_in+out": alg / 1e6, "fused_supported": ops.fu_fused_supported(B, C, C, N, N)}
        with torch.no_grad():
            for mode in ("train", "eval"):
                m.train(mode == "train")
                if row["fused_supported"]:
                    m.fused = True
                    row[f"fused_{mode}_us"] = 1000 * timeit(m, xs)
                m.fused = False
                row[f"general_{mode}_us"] = 1000 * timeit(m, xs)
                tf = torch_fu(m)
                row[f"torch_{mode}_us"] = 1000 * timeit(lambda x: tf(x, mode == "train"), xs)
        best = min(row.get("fused_eval_us", 1e30), row["general_eval_us"])
        row["best_eval_GBs"] = alg / best / 1e3
        row["best_eval_frac"] = row["best_eval_GBs"] / PEAK
        bt = min(row.get("fused_train_us", 1e30), row["general_train_us"])
        row["best_train_GBs"] = alg / bt / 1e3
        row["best_train_frac"] = row["best_train_GBs"] / PEAK
        rows.append(row)
        print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in row.items()}), flush=True)
    if "--json" in sys.argv:
        json.dump(rows, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
