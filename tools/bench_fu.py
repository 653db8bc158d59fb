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
import torch.nn.functional as F

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


def ref_fourier_unit(x, P, training):
    """The reference's op sequence (layers/ffc/fourier_unity.py:32-58) on PyTorch's GPU kernels (cuFFT, cuDNN): the
    comparison arm of this benchmark.  Restated here so that tools/ does not import the test oracle."""
    b, c, h, w = x.shape
    s = torch.fft.rfftn(x, dim=(-2, -1), norm="ortho")                                       # :38
    s = torch.stack((s.real, s.imag), dim=2).reshape(b, 2 * c, h, s.shape[-1])               # :40-42
    y = F.conv2d(s, P["conv_layer.weight"])                                                  # :45
    y = F.relu(F.batch_norm(y, P["bn.running_mean"], P["bn.running_var"], P["bn.weight"], P["bn.bias"], training, 0.1, 1e-5))   # :49
    y = y.reshape(b, -1, 2, h, y.shape[-1])
    yc = torch.complex(y[:, :, 0].contiguous(), y[:, :, 1].contiguous())                     # :51-53
    return torch.fft.irfftn(yc, s=(h, w), dim=(-2, -1), norm="ortho")                        # :56


def torch_fu(mod):
    P = dict(mod.state_dict())
    def f(x, training):
        return ref_fourier_unit(x, P, training)
    return f


def timeit_fwd_bwd(fn, xs, iters=5):
    """fn(x) -> output; times forward + backward (all gradients) per call, graph-captured like timeit."""
    gs = [torch.randn_like(fn(xs[0]).detach())]
    def step(x):
        x.grad = None
        fn(x).backward(gs[0])
    for i in range(2):
        step(xs[i % len(xs)])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        with torch.cuda.graph(g, stream=stream):
            for x in xs:
                step(x)
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


def ref_spectral_transform(x, P, training):
    """The reference's SpectralTransform.forward (layers/ffc/spectral_transform.py:77-110, stride 1) on PyTorch's GPU kernels."""
    b, c = x.shape[:2]
    g = torch.sigmoid(F.linear(F.relu(F.linear(x.mean(dim=(2, 3)), P["se_block.fc.0.weight"])), P["se_block.fc.2.weight"]))
    x = x * g.view(b, c, 1, 1)
    x = F.relu(F.batch_norm(F.conv2d(x, P["conv1.weight"]), P["bn1.running_mean"], P["bn1.running_var"], P["bn1.weight"], P["bn1.bias"],
                            training, 0.1, 1e-5))
    f = ref_fourier_unit(x, {k[3:]: v for k, v in P.items() if k.startswith("fu.")}, training)
    return F.conv2d(x + f, P["conv2.weight"])


def sweep_config5(out_path):
    """BASELINE.json configs[4]: FourierUnitSN(Cf, Cf) and SpectralTransform(Cg, Cg, stride 1), Cg = int(C * r), Cf = Cg // 2 for
    C in {64, 256, 512}, r in {.25, .5, .75}, H = W in {16, 32, 64, 128}, batch 32 and 128, training and eval mode, forward
    and forward + backward, against the reference's op sequence on the same GPU (cuFFT + cuDNN, TF32 off).
    Algorithmic bytes: FU 8*B*C*H*W forward, 20*B*C*H*W fwd + bwd; ST 8*B*Cg*H*W forward, 20*B*Cg*H*W fwd + bwd."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfs = sorted({int(C * r) // 2 for C in (64, 256, 512) for r in (0.25, 0.5, 0.75)})
    with open(out_path, "w") as f:
        for Cf in cfs:
            for N in (16, 32, 64, 128):
                for B in (32, 128):
                    for kind in ("fu", "st"):
                        C = Cf if kind == "fu" else 2 * Cf
                        nbytes = 4 * B * C * N * N
                        if nbytes > 1.7e9 or (B == 128 and nbytes > 0.9e9):
                            continue
                        torch.manual_seed(0)
                        m = (ffc.FourierUnitSN(C, C) if kind == "fu" else ffc.SpectralTransform(C, C, 1, 1, True, False)).to(DEV)
                        nbuf = min(max(2, int(200e6 // nbytes) + 1), 6)
                        xs = [torch.randn(B, C, N, N, device=DEV, requires_grad=True) for _ in range(nbuf)]
                        P = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k and "lfu" not in k)
                             for k, v in m.state_dict().items()}
                        reff = ref_fourier_unit if kind == "fu" else ref_spectral_transform
                        fu_c = C if kind == "fu" else C // 2
                        path = ("single-kernel" if ops.fu_fused_supported(B, fu_c, fu_c, N, N) else
                                "L2-staged" if ops.fu_staged_supported(B, fu_c, fu_c, N, N) else "general")
                        row = {"module": "FourierUnitSN" if kind == "fu" else "SpectralTransform", "B": B, "C": C, "N": N, "fu_path": path}
                        with torch.no_grad():
                            for mode in ("train", "eval"):
                                m.train(mode == "train")
                                row[f"ours_fwd_{mode}_us"] = 1000 * timeit(m, xs, iters=3)
                                row[f"torch_fwd_{mode}_us"] = 1000 * timeit(lambda x: reff(x, P, mode == "train"), xs, iters=3)
                        m.train()
                        row["ours_fwd_bwd_us"] = 1000 * timeit_fwd_bwd(m, xs, iters=3)
                        row["torch_fwd_bwd_us"] = 1000 * timeit_fwd_bwd(lambda x: reff(x, P, True), xs, iters=3)
                        row["fwd_train_frac_hbm"] = 2.0 * nbytes / row["ours_fwd_train_us"] / 1e3 / PEAK
                        row["fwd_eval_frac_hbm"] = 2.0 * nbytes / row["ours_fwd_eval_us"] / 1e3 / PEAK
                        row["fwd_bwd_frac_hbm"] = 5.0 * nbytes / row["ours_fwd_bwd_us"] / 1e3 / PEAK
                        row["speedup_fwd_train"] = row["torch_fwd_train_us"] / row["ours_fwd_train_us"]
                        row["speedup_fwd_eval"] = row["torch_fwd_eval_us"] / row["ours_fwd_eval_us"]
                        row["speedup_fwd_bwd"] = row["torch_fwd_bwd_us"] / row["ours_fwd_bwd_us"]
                        line = json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in row.items()})
                        print(line, flush=True)
                        f.write(line + "\n")
                        f.flush()
                        del xs, m, P
                        torch.cuda.empty_cache()


def main():
    if "--config5" in sys.argv:
        return sweep_config5(sys.argv[sys.argv.index("--config5") + 1])
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
