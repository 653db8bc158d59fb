#!/usr/bin/env python
"""Benchmark of the FFC hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload fgan32]

Workload (BASELINE.json configs[1]): fgan_complete FFC generator + spectral-norm discriminator GAN
training step on synthetic CIFAR-shaped data (3x32x32), global batch 256, AdamW, hinge losses --
G fwd x2, G bwd, D fwd x3, D bwd x2, two optimiser steps (fgan_complete.py:357-393).
metric = images/s (global batch / step time, max over ranks).

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM.  `e2e`: the same step through the
public module API with pinned HOST buffers, host->device copies of z / real images and the device->host
read of both losses inside the timed region.  `roofline`: the FourierUnit kernels of this workload against
the measured HBM peak.  `cpu_baseline`: the CPU restatement of the reference step (oracle/train_ref.py),
timed on this host's cores on a bounded sample.  `--impl reference` times that CPU step alone.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

WORKLOADS = {
    # name: (generator variant, discriminator convs, image size, default global batch, weak scaling?)
    "fgan32": ("fgan32", 7, 32, 256, False),     # configs[1]: global batch 256 on 1/2/4/8 GPUs (strong)
    "fgan64": ("fgan64", 8, 64, 128, True),      # configs[2]: 128 per GPU (weak)
    "fgan128": ("fgan128", 9, 128, 64, True),    # configs[3]: 64 per GPU (weak)
    # configs[2] with the "spectral-norm snffc discriminator" it names (harness.FDiscriminatorSN64):
    # the reference ships no such class; it is assembled from FFC_BN_ACT + SNFFC and pinned by tests/golden/model_fgan64_FD.npz
    "fgan64_snffc": ("fgan64", "fd64", 64, 128, True),
}
# FourierUnit instances inside each generator: (channels, plane size) -- SURVEY.md appendix A.2/A.3
FU_SHAPES = {"fgan32": [(16, 16), (8, 32)], "fgan64": [(16, 16), (8, 32), (8, 64)],
             "fgan128": [(64, 16), (32, 32), (32, 64), (32, 128)]}
FU_SHAPES["fgan64_snffc"] = FU_SHAPES["fgan64"]


def workload_config(workload, gb):
    """The ``config`` object of the JSON line: identical for our arm and the reference arm (same workload, same batch)."""
    _, n_convs, size, _, _ = WORKLOADS[workload]
    d = "SNFFC discriminator (FFC_BN_ACT + SNFFC stages)" if n_convs == "fd64" else f"SN conv Discriminator ({n_convs} convs)"
    return {"workload": f"{workload}: fgan_complete-style FGenerator + {d} GAN training step (G fwd x2, G bwd, D fwd x3, D bwd x2, "
                        f"2 optimiser steps), synthetic 3x{size}x{size}, global batch {gb}",
            "global_batch": gb,
            "l2": "per-step working set (activations + gradients of the global batch, > 400 MB) exceeds the 126 MB L2; "
                  "isolated-kernel timings rotate input buffers whose total exceeds L2"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU restatement of the reference training step
# ---------------------------------------------------------------------------------------------
def cpu_reference_step_rate(workload, batch, steps, warmup, seed=1234):
    from fastfourierconvolution_b200 import harness as H
    from oracle.train_ref import RefTrainer
    variant, n_convs, size, _, _ = WORKLOADS[workload]
    torch.manual_seed(seed)
    torch.set_num_threads(os.cpu_count() or 1)
    G = H.FGenerator(128, 4, variant); G.apply(H.weights_init)
    D = H.FDiscriminatorSN64(True, 4) if n_convs == "fd64" else H.SNDiscriminator(True, 4, n_convs)
    D.apply(H.weights_init)
    tr = RefTrainer(G.state_dict(), D.state_dict(), variant, n_convs)
    def data():
        return torch.randn(batch, 128), torch.randn(batch, 128), torch.rand(batch, 3, size, size) * 2 - 1
    for _ in range(warmup):
        tr.step(*data())
    times = []
    for _ in range(steps):
        d = data()
        t0 = time.perf_counter()
        tr.step(*d)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * steps / total, 1000.0 * total / steps, torch.get_num_threads()


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    workload = args.workload
    gb = args.batch or WORKLOADS[workload][3] * (args.gpus if WORKLOADS[workload][4] else 1)
    # the CPU step runs the workload's own global batch (same config as our arm); --cpu-batch N caps it for a quick look
    sample = min(gb, args.cpu_batch) if args.cpu_batch else gb
    rate, ms, cores = cpu_reference_step_rate(workload, sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "FFC-GAN training images/s", "value": rate, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak" if WORKLOADS[workload][4] else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(workload, gb),
        "notes": {"cpu_step_batch": sample, "threads": cores},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps at batch {sample} (of global batch {gb}) of the same G+D step, oracle/train_ref.py on PyTorch CPU kernels"},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def _graph_time(fn, xs, replays=10):
    """Device time per call of fn(x): one pass over the rotating inputs xs (total footprint > L2) is captured
    into a CUDA graph and replayed; CUDA events bracket the replays on the replay stream."""
    for x in xs[:2]:
        fn(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for x in xs:
                fn(x)
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (replays * len(xs))


def time_fourier_unit(workload, per_rank_batch, dev):
    """The FourierUnit kernels on the workload's largest unit, timed alone on the device."""
    C, N = max(FU_SHAPES[workload], key=lambda s: s[0] * s[1] * s[1])
    return time_fourier_unit_shape(C, N, per_rank_batch, dev)


def fu_roofline(fu, peak, peak_src, traffic):
    kern = (f"fused FourierUnit forward ({fu['launches_fwd_train']} launch(es))" if fu["fused"] else
            f"L2-staged FourierUnit forward ({fu['launches_fwd_train']} launches: rfft2 | tensor-core mix + BN statistics | BN+ReLU -> irfft2)")
    return {"bound": "hbm", "achieved": fu["gbs_fwd_train"], "peak": peak, "unit": "GB/s", "frac": fu["gbs_fwd_train"] / peak,
            "traffic": traffic, "peak_source": peak_src,
            "kernel": f"{kern}, FourierUnitSN({fu['C']},{fu['C']}) @ {fu['N']}x{fu['N']}, batch {fu['B']}, training mode",
            "algorithmic_bytes_per_launch": fu["alg_bytes_fwd"], "us_per_launch": 1000 * fu["ms_fwd_train"],
            "eval_mode_single_pass": {"us": 1000 * fu["ms_fwd_eval"], "achieved": fu["gbs_fwd_eval"], "frac": fu["gbs_fwd_eval"] / peak},
            "fwd_plus_bwd": {"us": 1000 * fu["ms_fwd_bwd_train"], "achieved": fu["gbs_fwd_bwd_train"], "frac": fu["gbs_fwd_bwd_train"] / peak,
                             "algorithmic_bytes": 20.0 * fu["B"] * fu["C"] * fu["N"] * fu["N"]},
            "timing": "CUDA events around CUDA-graph replays of the op over %d rotating inputs (> L2)" % fu["rotating_buffers"]}


def time_fourier_unit_shape(C, N, B, dev):
    import fastfourierconvolution_b200 as ffc
    from fastfourierconvolution_b200 import _C
    torch.manual_seed(0)
    fu = ffc.FourierUnitSN(C, C).to(dev)
    bytes_in = 4 * B * C * N * N
    nbuf = min(max(2, int(160e6 // bytes_in) + 1), 24)          # rotation larger than the 126 MB L2
    xs = [torch.randn(B, C, N, N, device=dev) for _ in range(nbuf)]
    L = _C.lib()
    out = {"C": C, "N": N, "B": B, "rotating_buffers": nbuf, "fused": bool(L.ffc_fu_fused_supported(B, C, C, N, N))}
    with torch.no_grad():
        fu.train()
        n0 = L.ffc_launch_count()
        fu(xs[0])
        out["launches_fwd_train"] = L.ffc_launch_count() - n0
        out["fused"] = out["launches_fwd_train"] <= 3            # single-kernel form (1 launch; + zero fill, + statistics pass in its two-pass variant); the L2-staged form has >= 4
        out["ms_fwd_train"] = _graph_time(lambda x: fu(x), xs)
        fu.eval()
        out["ms_fwd_eval"] = _graph_time(lambda x: fu(x), xs)
    fu.train()
    xg = [x.clone().requires_grad_(True) for x in xs[:max(2, nbuf // 3)]]
    gy = torch.randn(B, C, N, N, device=dev)

    def fwd_bwd(x):
        return torch.autograd.grad(fu(x), (x, fu.conv_layer.weight, fu.bn.weight, fu.bn.bias), gy)
    out["ms_fwd_bwd_train"] = _graph_time(fwd_bwd, xg)
    alg_f = 4.0 * B * N * N * (C + C)                      # fwd 4BHW(Cin+Cout)
    alg_fb = 20.0 * B * C * N * N                          # fwd + bwd
    out["alg_bytes_fwd"] = alg_f
    out["gbs_fwd_train"] = alg_f / out["ms_fwd_train"] / 1e6
    out["gbs_fwd_eval"] = alg_f / out["ms_fwd_eval"] / 1e6
    out["gbs_fwd_bwd_train"] = alg_fb / out["ms_fwd_bwd_train"] / 1e6
    return out


def time_generation(G, z, steps, dev, world):
    """Pure generation (G.eval() forward, uint8 images out for the fgan generators): batch shards, no collective.  The forward
    is captured once into a CUDA graph and replayed (like the training step); CUDA events, max over ranks."""
    import torch.distributed as dist
    was_training = G.training
    G.eval()
    with torch.no_grad():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                G(z)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = G(z)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(); torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    G.train(was_training)
    return {"value": z.shape[0] * world / ms * 1000.0, "unit": "images/s", "ms_per_batch": ms, "per_gpu_batch": z.shape[0],
            "out_dtype": str(out.dtype).replace("torch.", ""), "collectives": 0,
            "what": "G.eval() forward on resident latents (one CUDA-graph replay per batch), batch sharded over the ranks "
                    "(fgan_complete.py:413-427 path)"}


# ---------------------------------------------------------------------------------------------
# configs[0]: FFC-DCGAN generator (models/ffc_generator.py) forward + backward, batch 128 per GPU
# ---------------------------------------------------------------------------------------------
CFG1_BATCH = 128


def cfg1_config(gb):
    return {"workload": "ffcgen_cfg1: models/ffc_generator.py FFCGenerator(nz=100, nc=1, ngf=32, g_factor=0.5) forward + backward of "
                        f"out.mean(), z (B,100,1,1) -> (B,1,64,64), batch {gb}",
            "global_batch": gb,
            "l2": "activations + gradients of the batch (> 200 MB) exceed the 126 MB L2"}


def run_reference_cfg1(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from fastfourierconvolution_b200 import harness as H
    from oracle import ffc_ref as R
    gb = args.batch or CFG1_BATCH * args.gpus
    sample = min(gb, args.cpu_batch) if args.cpu_batch else gb
    torch.manual_seed(1234)
    torch.set_num_threads(os.cpu_count() or 1)
    G = H.FFCGenerator(100, 1, 32); G.apply(H.weights_init)
    P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in G.state_dict().items()}
    leaves = [v for v in P.values() if v.requires_grad]

    def step():
        z = torch.randn(sample, 100, 1, 1)
        t0 = time.perf_counter()
        out = R.ffc_generator(z, P, True)
        torch.autograd.grad(out.mean(), leaves, allow_unused=True)
        return time.perf_counter() - t0
    for _ in range(args.warmup):
        step()
    total = sum(step() for _ in range(args.steps))
    rate, ms, cores = sample * args.steps / total, 1000.0 * total / args.steps, torch.get_num_threads()
    print(json.dumps({
        "impl": "reference", "metric": "FFC-DCGAN generator fwd+bwd images/s", "value": rate, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg1_config(gb), "notes": {"cpu_step_batch": sample, "threads": cores},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} fwd+bwd passes at batch {sample}, oracle/ffc_ref.py ffc_generator on PyTorch CPU kernels"},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)


def run_ours_cfg1(args):
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: libffc_b200 has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    from fastfourierconvolution_b200 import _C, harness as H
    from fastfourierconvolution_b200.harness.train import FlatGradAllReduce
    L = _C.lib()
    gb = args.batch or CFG1_BATCH * world
    pb = gb // world
    torch.manual_seed(1234 + rank)
    G = H.FFCGenerator(100, 1, 32).to(dev).train(); G.apply(H.weights_init)
    if world > 1:
        for p in list(G.parameters()) + list(G.buffers()):
            dist.broadcast(p.data, 0)
    reduce_G = FlatGradAllReduce(G.parameters())
    h_z = torch.randn(pb, 100, 1, 1).pin_memory()
    z = h_z.to(dev)
    h_loss = torch.zeros(1).pin_memory()

    def step():
        G.zero_grad(set_to_none=True)
        loss = G(z).mean()
        loss.backward()
        reduce_G()
        return loss.detach()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize(dev)
    n0 = L.ffc_launch_count()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_loss = step()
    launches = L.ffc_launch_count() - n0

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(); torch.cuda.synchronize(dev)

    def resident():
        graph.replay()

    def e2e():
        z.copy_(h_z, non_blocking=True)
        graph.replay()
        h_loss.copy_(static_loss.reshape(1), non_blocking=False)
    for _ in range(max(args.warmup, 3)):
        resident()

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / args.steps
    with ClockSampler(local_rank) as clk:
        ms_res = timed(resident)
        ms_e2e = timed(e2e)
    gen = time_generation(G, z, args.steps, dev, world)
    if rank == 0:
        peak, peak_src = measured_peaks()
        fu = time_fourier_unit_shape(32, 8, pb, dev)
        line = {"metric": "FFC-DCGAN generator fwd+bwd images/s", "value": gb / ms_res * 1000.0, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg1_config(gb),
                "notes": {"per_gpu_batch": pb, "launch": "forward + backward replayed as one CUDA graph"},
                "e2e": {"value": gb / ms_e2e * 1000.0, "unit": "images/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": world * h_z.numel() * 4,
                        "d2h_bytes_per_step": world * 4},
                "gpu_launches": int(launches * args.steps), "gpu_launches_per_step": launches, "generation": gen,
                "roofline": fu_roofline(fu, peak, peak_src, None), "clocks": clk.summary()}
        if not args.no_cpu_baseline and world == 1:
            import subprocess as sp
            r = sp.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", "ffcgen_cfg1", "--steps", str(args.cpu_steps),
                        "--warmup", "1"], capture_output=True, text=True)
            try:
                line["cpu_baseline"] = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])["cpu_baseline"]
            except Exception:
                line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); torch.cuda.synchronize(dev); sys.stdout.flush(); os._exit(0)


# largest local-branch convolution of each generator: (cin, cout, input size) of conv2.ffc.convl2l, ConvTranspose2d k4 s2 p1
CONV_SHAPES = {"fgan32": (512, 192, 4), "fgan64": (512, 192, 4), "fgan128": (1024, 256, 4), "fgan64_snffc": (512, 192, 4)}


def tensor_peak_tf32():
    """Dense TF32 tensor peak: half of the measured BF16 cuBLAS burst figure (MEASURED_PEAKS.json), else of the fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return 0.5 * float(json.load(f)["bf16_tflops"]), "0.5 x measured bf16_tflops (MEASURED_PEAKS.json)"
    return 0.5 * 1590.0, "0.5 x fallback bf16 (B200_PROFILING.md)"


def time_conv_layer(workload, per_rank_batch, dev):
    """The tcgen05 implicit-GEMM convolution (ConvFwdV5, 3xTF32) on the generator's largest local convolution."""
    from fastfourierconvolution_b200 import ops
    cin, cout, hi = CONV_SHAPES[workload]
    B = per_rank_batch
    torch.manual_seed(0)
    w = torch.randn(cin, cout, 4, 4, device=dev) * 0.02
    nbuf = min(max(2, int(160e6 // (4 * B * cout * 4 * hi * hi)) + 1), 16)
    xs = [torch.randn(B, cin, hi, hi, device=dev) for _ in range(nbuf)]
    with torch.no_grad():
        ms = _graph_time(lambda x: ops.conv2d(x, w, stride=2, pad=1, transposed=True), xs)
    flops = 2.0 * B * (2 * hi) ** 2 * cout * cin * 4            # 16 taps / 4 parity classes per output pixel
    return {"ms": ms, "flops": flops, "shape": f"ConvTranspose2d({cin}->{cout}, k4 s2 p1) {hi}x{hi}->{2 * hi}x{2 * hi}, batch {B}",
            "rotating_buffers": nbuf}


def run_ours(args):
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: libffc_b200 has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        # NCCL writes its version banner (NCCL_DEBUG=VERSION and up) to stdout by default; stdout carries the JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"WORLD_SIZE={world} but --gpus {args.gpus}"
    from fastfourierconvolution_b200 import _C, harness as H
    L = _C.lib()
    workload = args.workload
    variant, n_convs, size, base_batch, weak = WORKLOADS[workload]
    gb = args.batch or (base_batch * world if weak else base_batch)
    assert gb % world == 0
    pb = gb // world
    torch.manual_seed(1234 + rank)
    # Everything on the step's critical path is FP32-accurate: the FFC generator AND the SN conv
    # discriminator run on libffc_b200's kernels.  FFC_BENCH_D=torch keeps the discriminator on nn.Conv2d.forward
    # (cuDNN, autotuned, NHWC, TF32 allowed -- PyTorch's GPU defaults) for an A/B comparison only.
    d_backend = os.environ.get("FFC_BENCH_D", "ffc_b200")
    G = H.FGenerator(128, 4, variant).to(dev).train(); G.apply(H.weights_init)
    if n_convs == "fd64":
        D = H.FDiscriminatorSN64(True, 4).to(dev).train()        # an FFC discriminator: on the product's kernels by construction
        d_backend = "ffc_b200"
    else:
        D = H.SNDiscriminator(True, 4, n_convs, backend="ffc_b200" if d_backend == "ffc_b200" else "torch").to(dev).train()
    D.apply(H.weights_init)
    if d_backend != "ffc_b200":                      # "torch" (TF32 allowed) or "torch_fp32" (cuDNN held to FP32 like the product)
        torch.backends.cudnn.allow_tf32 = d_backend != "torch_fp32"
        torch.backends.cuda.matmul.allow_tf32 = d_backend != "torch_fp32"
        torch.backends.cudnn.benchmark = True
        D = D.to(memory_format=torch.channels_last); D.channels_last = True
    if world > 1:                                    # identical replicas
        for p in list(G.parameters()) + list(D.parameters()) + list(G.buffers()) + list(D.buffers()):
            dist.broadcast(p.data, 0)
    use_graph = not args.no_graph
    tr = H.GanTrainer(G, D, capturable=use_graph)

    h_zg = torch.randn(pb, 128).pin_memory(); h_zd = torch.randn(pb, 128).pin_memory()
    h_real = (torch.rand(pb, 3, size, size) * 2 - 1).pin_memory()
    d_zg, d_zd, d_real = h_zg.to(dev), h_zd.to(dev), h_real.to(dev)
    h_loss = torch.zeros(2).pin_memory()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    import warnings
    warnings.filterwarnings("ignore", message=".*lr_scheduler.step.*")
    if use_graph:
        # the whole G+D step (our kernels, the PyTorch discriminator, both optimisers and, for N > 1, the NCCL
        # all-reduces) is captured once into a CUDA graph and replayed: same kernels, no host launch cost
        tr.capture(d_zg, d_zd, d_real)

        def step_resident():
            return tr.step_graphed(d_zg, d_zd, d_real)

        def step_e2e():
            h_loss.copy_(tr.step_graphed(h_zg, h_zd, h_real), non_blocking=False)   # H2D of inputs, D2H of both losses
            return h_loss
    else:
        def step_resident():
            return tr.step(d_zg, d_zd, d_real)

        def step_e2e():
            zg = h_zg.to(dev, non_blocking=True); zd = h_zd.to(dev, non_blocking=True); re = h_real.to(dev, non_blocking=True)
            lg, ld = tr.step(zg, zd, re)
            h_loss.copy_(torch.stack((lg, ld)), non_blocking=False)       # device -> host read of the step's result
            return h_loss

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sync_all()

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        n0 = L.ffc_launch_count()
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        per_step = tr.launches_per_step if use_graph else (L.ffc_launch_count() - n0) / args.steps
        return ms / args.steps, per_step

    with ClockSampler(local_rank) as clk:
        ms_res, launches = timed(step_resident)
        ms_e2e, _ = timed(step_e2e)
    clocks = clk.summary()

    if args.profile_step:                            # launch lists under ncu: the timed steps only, no micro-benchmarks
        if rank == 0:
            print(json.dumps({"profile_step": True, "ms_per_step": ms_res, "gpu_launches_per_step": launches}), flush=True)
        if world > 1:
            dist.barrier(); torch.cuda.synchronize(dev); sys.stdout.flush(); os._exit(0)
        return
    gen = time_generation(G, d_zg, args.steps, dev, world)         # every rank: generation shards the batch, no collective
    fu = time_fourier_unit(workload, pb, dev) if rank == 0 else None
    cv = time_conv_layer(workload, pb, dev) if rank == 0 else None
    fu_all = []
    if rank == 0:                                     # every Fourier unit of the generator at this batch, next to the headline one
        peak0, _ = measured_peaks()
        for (c, n) in FU_SHAPES[workload]:
            f = fu if (c, n) == (fu["C"], fu["N"]) else time_fourier_unit_shape(c, n, pb, dev)
            fu_all.append({"C": c, "N": n, "B": pb, "fused_single_kernel": f["fused"], "launches": f["launches_fwd_train"],
                           "us_fwd_train": 1000 * f["ms_fwd_train"], "frac_fwd_train": f["gbs_fwd_train"] / peak0,
                           "us_fwd_eval": 1000 * f["ms_fwd_eval"], "frac_fwd_eval": f["gbs_fwd_eval"] / peak0,
                           "us_fwd_bwd": 1000 * f["ms_fwd_bwd_train"], "frac_fwd_bwd": f["gbs_fwd_bwd_train"] / peak0})
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = min(gb, args.cpu_batch) if args.cpu_batch else gb
        rate, ms, cores = cpu_reference_step_rate(workload, cb, args.cpu_steps, 1)
        cpu = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_steps} steps at batch {cb} (the global batch) of the same G+D training step "
                         f"(oracle/train_ref.py, PyTorch CPU kernels, {ms:.0f} ms/step)"}
    if rank == 0:
        peak, peak_src = measured_peaks()
        traffic = None                                    # DRAM bytes per launch from the committed ncu capture (same shape only)
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic_fu.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("shape") == {"B": fu["B"], "C": fu["C"], "N": fu["N"]}:
                traffic = tj["traffic"]
        act_mb = sum(p.numel() for p in G.parameters()) * 4 / 1e6
        line = {
            "metric": "FFC-GAN training images/s", "value": gb / ms_res * 1000.0, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_res,
            "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(workload, gb),
            "notes": {"per_gpu_batch": pb, "parallelism": f"dp{world} (batch shards, per-rank BatchNorm, flat-grad NCCL all-reduce)",
                      "launch": "whole step replayed as one CUDA graph" if use_graph else "eager launches",
                      "discriminator": ("SNFFC discriminator (harness.FDiscriminatorSN64) on libffc_b200 kernels" if n_convs == "fd64" else "plain SN conv net on libffc_b200 kernels (tcgen05 conv / dgrad / wgrad at FP32 accuracy, fused bias, LeakyReLU kernel)"
                                        if d_backend == "ffc_b200" else "plain SN conv net on PyTorch kernels (cuDNN autotuned, NHWC, %s)" % ("FP32" if d_backend == "torch_fp32" else "TF32")),
                      "generator_params_MB": round(act_mb, 1)},
            "e2e": {"value": gb / ms_e2e * 1000.0, "unit": "images/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": world * (h_zg.numel() + h_zd.numel() + h_real.numel()) * 4,
                    "d2h_bytes_per_step": world * 8},
            "gpu_launches": int(round(launches * args.steps)),
            "gpu_launches_per_step": launches,
            "roofline": fu_roofline(fu, peak, peak_src, traffic),
            "fourier_units": fu_all,
            "generation": gen,
            "clocks": clocks,
        }
        tpeak, tpeak_src = tensor_peak_tf32()
        issued = 3.0 * cv["flops"] / cv["ms"] / 1e9          # three TF32 MMAs per FP32-accurate product (3xTF32)
        ctraffic = None                                   # DRAM bytes per launch from the committed ncu capture (same shape only)
        cpath = os.path.join(ROOT, "profiles", "roofline_traffic_conv.json")
        if os.path.exists(cpath):
            cj = json.load(open(cpath))
            if cj.get("shape") == {"workload": workload, "B": pb}:
                ctraffic = cj["traffic"]
        effective = cv["flops"] / cv["ms"] / 1e9             # useful (FP32-accurate) flops: the headline fraction
        line["roofline_tensor"] = {
            "bound": "tensor", "achieved": effective, "peak": tpeak, "unit": "TFLOP/s", "frac": effective / tpeak, "traffic": ctraffic,
            "peak_source": tpeak_src,
            "kernel": "conv_v5_kernel (tcgen05.mma kind::tf32, A in TMEM, 3xTF32 at FP32 accuracy) + pack_v5_kernel, " + cv["shape"],
            "effective_fp32_tflops": effective, "algorithmic_flops_per_launch": cv["flops"],
            "issued_tf32_tflops": issued, "issued_frac": issued / tpeak,
            "note": "achieved / frac count each FP32-accurate product once; the tensor cores execute three TF32 products for it (issued_*)",
            "us_per_launch": 1000 * cv["ms"],
            "timing": "CUDA events around CUDA-graph replays over %d rotating inputs; includes the per-call weight packing kernel" % cv["rotating_buffers"]}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        # NCCL communicator teardown can block while captured graphs still hold its kernels: leave without it
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--workload", default="fgan32", choices=sorted(WORKLOADS) + ["ffcgen_cfg1"])
    ap.add_argument("--batch", type=int, default=0, help="global batch (default: the workload's)")
    ap.add_argument("--cpu-batch", type=int, default=0, help="cap on the batch of the CPU step (default 0: the workload's global batch)")
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-step", action="store_true", help="run the warm-up and the timed steps only (for ncu launch lists)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.workload == "ffcgen_cfg1":
        (run_reference_cfg1 if args.impl == "reference" else run_ours_cfg1)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
