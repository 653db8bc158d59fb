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


class FlatAdam(torch.optim.Optimizer):
    """optim.AdamW (fgan_complete.py:315-319) / optim.Adam (sngan_complete.py:247-248) as ONE kernel per step for any number
    of parameter tensors (``ffc_adam_step_table``).

    ``adopt()`` -- called once, after the first backward -- moves every parameter that received a gradient into one flat
    FP32 buffer (``p.data`` becomes a view); both moments are flat too.  Gradients stay where autograd puts them (with
    ``grad = None`` before a backward autograd hands over its buffer without a copy or an accumulation kernel) and are
    read through a table of device pointers.  With several ranks the same table packs the gradients into one flat buffer
    (one launch), the all-reduce runs in place on it and its 1 / world is folded into the step.  Parameters that never
    receive a gradient (the unused ``lfu.*`` of SpectralTransform, spectral_transform.py:65-67) stay outside, untouched,
    exactly as torch's optimisers skip parameters whose ``grad`` is None.  The learning rate and the step count live on
    the device, so a CUDA-graph replay of the step sees what the LR scheduler wrote between replays."""

    PIECE = 4096          # elements per block of the table kernels (csrc/ffc_glue.cu: kAdamPiece)

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, decoupled=True, group=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.decoupled, self.group = decoupled, group
        self.flat_p = self.flat_g = self.m = self.v = None
        self._members: List[torch.nn.Parameter] = []
        self._reduced = False

    @property
    def adopted(self) -> bool:
        return self.flat_p is not None

    def adopt(self):
        if self.adopted:
            return self
        assert len(self.param_groups) == 1, "FlatAdam keeps one parameter group"
        members = [p for p in self.param_groups[0]["params"] if p.grad is not None]
        if not members:
            return self
        dev = members[0].device
        offs, n = [], 0
        for p in members:
            offs.append(n)
            n += (p.numel() + 63) // 64 * 64                  # every tensor starts on a 256-byte boundary
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m, self.v = torch.zeros_like(self.flat_p), torch.zeros_like(self.flat_p)
        for p, o in zip(members, offs):
            view = self.flat_p[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self._members = members
        blk_t, blk_p = [], []
        for t, p in enumerate(members):
            for piece in range((p.numel() + self.PIECE - 1) // self.PIECE):
                blk_t.append(t); blk_p.append(piece)
        self._offs = torch.tensor(offs, dtype=torch.int64, device=dev)
        self._sizes = torch.tensor([p.numel() for p in members], dtype=torch.int64, device=dev)
        self._blk_t = torch.tensor(blk_t, dtype=torch.int32, device=dev)
        self._blk_p = torch.tensor(blk_p, dtype=torch.int32, device=dev)
        self._nblocks = len(blk_t)
        self._ptrs_host = torch.zeros(len(members), dtype=torch.int64)
        if dev.type == "cuda":
            self._ptrs_host = self._ptrs_host.pin_memory()
        self._ptrs = torch.zeros(len(members), dtype=torch.int64, device=dev)
        g = self.param_groups[0]
        self._step_t = torch.zeros(len(members), dtype=torch.float32, device=dev)     # per tensor, like torch's state["step"]
        if not torch.is_tensor(g["lr"]):
            self._lr_t = torch.full((1,), float(g["lr"]), dtype=torch.float32, device=dev)
        return self

    def _upload_grad_pointers(self):
        for i, p in enumerate(self._members):
            gr = p.grad
            if gr is not None and not gr.is_contiguous():
                gr = p.grad = gr.contiguous()
            self._ptrs_host[i] = gr.data_ptr() if gr is not None else 0
        self._ptrs.copy_(self._ptrs_host, non_blocking=True)

    def all_reduce(self) -> int:
        """Packs the gradients into one flat buffer (one launch), SUMs it over the process group in place; the step then reads
        the flat buffer.  Returns the bytes reduced (0 without a process group: nothing is packed, the step reads the
        gradients in place)."""
        self._reduced = False
        if not (self.adopted and dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return 0
        from .. import _C
        if self.flat_g is None:
            self.flat_g = torch.zeros_like(self.flat_p)
        self._upload_grad_pointers()
        _C.check(_C.lib().ffc_gather_table(_C.ptr(self.flat_g), _C.ptr(self._ptrs), _C.ptr(self._offs), _C.ptr(self._sizes),
                                           _C.ptr(self._blk_t), _C.ptr(self._blk_p), self._nblocks, _C.current_stream(self.flat_p.device)))
        dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group)
        self._reduced = True
        return self.flat_g.numel() * 4

    @torch.no_grad()
    def step(self, closure=None):
        from .. import _C, ops
        self.adopt()
        if not self.adopted:
            return None
        g = self.param_groups[0]
        lr = g["lr"]
        if not torch.is_tensor(lr):
            self._lr_t.fill_(float(lr))
            lr = self._lr_t
        b1, b2 = g["betas"]
        if self._reduced:                 # gradients were packed and summed over the ranks: plain flat step, 1 / world folded in
            world = dist.get_world_size(self.group)
            # (every member has a gradient in the packed buffer, so one count serves all; the per-tensor counts follow it)
            ops.adam_step(self.flat_p, self.flat_g, self.m, self.v, lr, self._step_t, b1, b2, g["eps"], g["weight_decay"],
                          1.0 / world, self.decoupled)
            self._step_t[1:] = self._step_t[0]
            self._reduced = False
            return None
        _C.require_device(self.flat_p, lr)
        self._upload_grad_pointers()
        _C.check(_C.lib().ffc_adam_step_table(_C.ptr(self.flat_p), _C.ptr(self.m), _C.ptr(self.v), _C.ptr(self._ptrs), _C.ptr(self._offs),
                                              _C.ptr(self._sizes), _C.ptr(self._blk_t), _C.ptr(self._blk_p), len(self._members), self._nblocks,
                                              _C.ptr(lr), _C.ptr(self._step_t), float(b1), float(b2), float(g["eps"]),
                                              float(g["weight_decay"]), 1.0, int(self.decoupled), _C.current_stream(self.flat_p.device)))
        return None


class GanTrainer:
    """One object per rank.  ``step(z_g, z_d, real)`` is one iteration of the reference loop:
    generator update (G fwd, D fwd, backward through both, optimiser step), then discriminator
    update (G fwd without graph, D fwd x2, hinge loss, backward, optimiser step), then LR decay."""

    def __init__(self, G, D, lr=2e-4, betas=(0.5, 0.999), num_total_steps=100000, optimizer="adamw", capturable=False):
        self.G, self.D = G, D
        # AdamW (fgan scripts) or Adam (sngan_complete.py:247-248) over flat buffers: one kernel per optimiser step, gradients
        # accumulate straight into the buffer the all-reduce runs on (torch defaults: AdamW decays by 0.01, Adam by 0)
        kw = dict(betas=betas, decoupled=True) if optimizer == "adamw" else dict(betas=betas, decoupled=False, weight_decay=0.0)
        if capturable:
            # whole-step CUDA-graph capture: the learning rate lives on the device (the step counter always does)
            lr = torch.tensor(float(lr), device=next(G.parameters()).device)
        self.optim_G = FlatAdam(G.parameters(), lr=lr.clone() if capturable else lr, **kw)
        self.optim_D = FlatAdam(D.parameters(), lr=lr.clone() if capturable else lr, **kw)
        self._graph = None
        decay = lambda step: 1.0 - step / num_total_steps                        # fgan_complete.py:318-319
        self.sched_G = torch.optim.lr_scheduler.LambdaLR(self.optim_G, decay)
        self.sched_D = torch.optim.lr_scheduler.LambdaLR(self.optim_D, decay)
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
        self.allreduce_bytes = self.optim_G.adopt().all_reduce()
        self.optim_G.step()
        # ---- discriminator update (:380-393, num_dis_updates = 1)
        G.requires_grad_(False)
        D.requires_grad_(True)
        self.optim_D.zero_grad()
        self.optim_G.zero_grad()
        fake = G(z_d)
        d_fake, d_real = self._two_forwards(D, fake, real)
        loss_D = hinge_loss_dis(d_fake, d_real)
        loss_D.backward()
        self.allreduce_bytes += self.optim_D.adopt().all_reduce()
        self.optim_D.step()
        if self._capturing:
            return loss_G.detach(), loss_D.detach()      # LR decay is applied outside the graph
        self.sched_G.step()
        self.sched_D.step()
        return loss_G.detach(), loss_D.detach()

    @staticmethod
    def _two_forwards(D, fake, real):
        """D(fake), then D(real) (the order of fgan_complete.py:384-385: the second forward's power iterations start from the
        first's u / v).  On a GPU the second forward is queued on a side stream: its convolutions do not depend on the first
        forward's, so the two chains -- and, through autograd's per-node streams, their backward chains -- overlap (a captured
        step records them as parallel branches; the spectral-norm launches of both stay in order on the streams of
        ``ops.fork_map``).  FFC_B200_SINGLE_STREAM=1 keeps everything on one stream."""
        from .. import _C
        sides = _C.side_streams(real.device, 6) if (real.is_cuda and getattr(D, "concurrent_forwards_ok", False)) else None
        if not sides:
            return D(fake), D(real)
        cur, side = torch.cuda.current_stream(real.device), sides[5]
        side.wait_stream(cur)                 # fork BEFORE the first forward is queued
        d_fake = D(fake)
        with torch.cuda.stream(side):
            d_real = D(real)
        cur.wait_stream(side)
        if not torch.cuda.is_current_stream_capturing():
            d_real.record_stream(cur)
        return d_fake, d_real

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
