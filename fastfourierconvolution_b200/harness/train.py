"""The GAN training step of the reference scripts (fgan_complete.py:357-393) on synthetic data,
batch-sharded over one process per GPU with a flat-buffer gradient all-reduce.

Per-rank BatchNorm statistics are kept (the semantics of the reference's only multi-GPU mode,
nn.DataParallel at train_cond.py:67-68); generation needs no collective at all.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from .models import hinge_loss_dis, hinge_loss_gen


class FlatGradAllReduce:
    """Averages the gradients of a parameter list over the process group with ONE in-place
    all-reduce of a flat FP32 buffer (NCCL over NVLink on GPUs, gloo in the CPU tests).

    Parameters whose ``grad`` is None (the never-used ``lfu.*`` of SpectralTransform,
    spectral_transform.py:65-67) are skipped on every rank alike, so the buffers stay aligned.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        self.group = group
        self._flat: Optional[torch.Tensor] = None

    def __call__(self) -> int:
        if not (dist.is_available() and dist.is_initialized()):
            return 0
        world = dist.get_world_size(self.group)
        if world == 1:
            return 0
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads:
            return 0
        n = sum(g.numel() for g in grads)
        if self._flat is None or self._flat.numel() != n or self._flat.device != grads[0].device:
            self._flat = torch.empty(n, dtype=torch.float32, device=grads[0].device)
        views = []
        off = 0
        for g in grads:
            views.append(self._flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        torch._foreach_copy_(views, grads)
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group)
        self._flat.div_(world)
        torch._foreach_copy_(grads, views)
        return n * 4


class GanTrainer:
    """One object per rank.  ``step(z_g, z_d, real)`` is one iteration of the reference loop:
    generator update (G fwd, D fwd, backward through both, optimiser step), then discriminator
    update (G fwd without graph, D fwd x2, hinge loss, backward, optimiser step), then LR decay."""

    def __init__(self, G, D, lr=2e-4, betas=(0.5, 0.999), num_total_steps=100000, optimizer="adamw", capturable=False):
        self.G, self.D = G, D
        opt = torch.optim.AdamW if optimizer == "adamw" else torch.optim.Adam   # sngan_complete.py:247-248 uses Adam
        kw = {}
        if capturable:
            # whole-step CUDA-graph capture: the optimiser keeps step counters and the learning rate on the device
            dev = next(G.parameters()).device
            kw = dict(capturable=True)
            if dev.type == "cuda":
                kw["fused"] = True          # one multi-tensor kernel per optimiser step instead of ~10 per parameter
            lr = torch.tensor(float(lr), device=dev)
        self.optim_G = opt(G.parameters(), lr=lr.clone() if capturable else lr, betas=betas, **kw)
        self.optim_D = opt(D.parameters(), lr=lr.clone() if capturable else lr, betas=betas, **kw)
        self._graph = None
        decay = lambda step: 1.0 - step / num_total_steps                        # fgan_complete.py:318-319
        self.sched_G = torch.optim.lr_scheduler.LambdaLR(self.optim_G, decay)
        self.sched_D = torch.optim.lr_scheduler.LambdaLR(self.optim_D, decay)
        self.reduce_G = FlatGradAllReduce(G.parameters())
        self.reduce_D = FlatGradAllReduce(D.parameters())
        self.allreduce_bytes = 0

    def step(self, z_g, z_d, real):
        G, D = self.G, self.D
        # ---- generator update (fgan_complete.py:368-377)
        G.requires_grad_(True)
        D.requires_grad_(False)
        self.optim_D.zero_grad()
        self.optim_G.zero_grad()
        loss_G = hinge_loss_gen(D(G(z_g)))
        loss_G.backward()
        self.allreduce_bytes = self.reduce_G()
        self.optim_G.step()
        # ---- discriminator update (:380-393, num_dis_updates = 1)
        G.requires_grad_(False)
        D.requires_grad_(True)
        self.optim_D.zero_grad()
        self.optim_G.zero_grad()
        fake = G(z_d)
        loss_D = hinge_loss_dis(D(fake), D(real))
        loss_D.backward()
        self.allreduce_bytes += self.reduce_D()
        self.optim_D.step()
        if self._capturing:
            return loss_G.detach(), loss_D.detach()      # LR decay is applied outside the graph
        self.sched_G.step()
        self.sched_D.step()
        return loss_G.detach(), loss_D.detach()

    # ---- whole-step CUDA graph (launch-bound at these layer sizes: ~500 kernels of a few microseconds each)
    _capturing = False

    def capture(self, z_g, z_d, real, warmup=3):
        """Captures one training step into a CUDA graph on static input buffers shaped like the arguments.
        The step is the same sequence of kernels as ``step``; only the host-side launch cost is removed."""
        assert z_g.is_cuda, "graph capture needs CUDA tensors"
        assert warmup >= 1, "at least one eager step must run first: optimiser state has to exist before the capture"
        self._static = (z_g.clone(), z_d.clone(), real.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(*self._static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from .. import _C
        n0 = _C.lib().ffc_launch_count()
        self._graph = torch.cuda.CUDAGraph()
        self._capturing = True
        try:
            with torch.cuda.graph(self._graph):
                self._static_losses = torch.stack(self.step(*self._static))
        finally:
            self._capturing = False
        self.launches_per_step = _C.lib().ffc_launch_count() - n0
        return self

    def step_graphed(self, z_g, z_d, real):
        """Copies the inputs (device or pinned host tensors) into the static buffers and replays the graph."""
        for dst, src in zip(self._static, (z_g, z_d, real)):
            dst.copy_(src, non_blocking=True)
        self._graph.replay()
        self.sched_G.step()
        self.sched_D.step()
        return self._static_losses
