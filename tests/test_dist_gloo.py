"""N > 1 host logic on CPU: world_size-2 gloo run of the flat-buffer gradient all-reduce used by the
batch-sharded GAN trainer (fastfourierconvolution_b200/harness/train.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fastfourierconvolution_b200.harness import FlatGradAllReduce
    torch.manual_seed(0)                                    # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    unused = torch.nn.Parameter(torch.zeros(4))             # like SpectralTransform.lfu.*: never gets a gradient
    params = list(net.parameters()) + [unused]
    torch.manual_seed(100 + rank)                           # different batch shard per rank
    x = torch.randn(8, 6)
    net(x).square().mean().backward()
    local = [p.grad.clone() for p in net.parameters()]
    nbytes = FlatGradAllReduce(params)()
    got = [p.grad.clone() for p in net.parameters()]
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    want = [sum(g[i] for g in gathered) / world for i in range(len(local))]
    ok = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(got, want)) and unused.grad is None
    ok = ok and nbytes == 4 * sum(p.numel() for p in net.parameters())
    # sharded batch == full batch: mean-of-shard gradients equals the gradient of the concatenated batch
    xs = [None] * world
    dist.all_gather_object(xs, x)
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    ref(torch.cat(xs)).square().mean().backward()
    ok = ok and all(torch.allclose(a, p.grad, atol=1e-6) for a, p in zip(got, ref.parameters()))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_grad_allreduce_world_size_2():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_allreduce_is_a_noop_without_process_group():
    from fastfourierconvolution_b200.harness import FlatGradAllReduce
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    assert FlatGradAllReduce([p])() == 0 and torch.equal(p.grad, torch.full((3,), 2.0))


def _trainer_worker(rank, world, port, out):
    """One GAN step of the batch-sharded trainer on 2 ranks (gloo; kernels on the host emulation build)."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emu_backend
    from fastfourierconvolution_b200 import harness as H
    with emu_backend.patched():
        torch.manual_seed(0)                                # identical replicas
        G = H.FGenerator(128, 4, "fgan32").train(); G.apply(H.weights_init)
        D = H.SNDiscriminator(True, 4, 7).train(); D.apply(H.weights_init)
        tr = H.GanTrainer(G, D)
        g0, d0 = G.conv2.ffc.convl2l.weight.detach().clone(), D.conv3.weight_orig.detach().clone()
        torch.manual_seed(100 + rank)                       # a different shard of the global batch on every rank
        lg, ld = tr.step(torch.randn(2, 128), torch.randn(2, 128), torch.rand(2, 3, 32, 32) * 2 - 1)
    ok = bool(torch.isfinite(lg) and torch.isfinite(ld)) and tr.allreduce_bytes > 0
    # identical initial weights + all-reduced gradients -> identical weights after the optimiser steps, on every rank
    mine = {"G/" + k: v.detach().clone() for k, v in G.named_parameters()}
    mine.update({"D/" + k: v.detach().clone() for k, v in D.named_parameters()})
    mine.update({"D/" + k: v.detach().clone() for k, v in D.named_buffers() if k.endswith("weight_u")})
    # BatchNorm statistics stay per rank (DataParallel semantics, train_cond.py:67-68): they must differ between the shards
    rm = G.conv2.bn_l.running_mean.detach().clone()
    gathered, rms = [None] * world, [None] * world
    dist.all_gather_object(gathered, mine)
    dist.all_gather_object(rms, rm)
    same = all(torch.equal(gathered[0][k], gathered[1][k]) for k in mine)
    moved = not torch.equal(g0, G.conv2.ffc.convl2l.weight) and not torch.equal(d0, D.conv3.weight_orig)
    # lfu.* never receives a gradient (spectral_transform.py:65-67): AdamW leaves it exactly as initialised
    torch.manual_seed(0)
    G_init = H.FGenerator(128, 4, "fgan32"); G_init.apply(H.weights_init)
    init = dict(G_init.named_parameters())
    unused = all(torch.equal(p, init[k]) for k, p in G.named_parameters() if ".lfu." in k)
    out[rank] = bool(ok and same and moved and unused and not torch.equal(rms[0], rms[1]))
    if not out[rank]:
        print("rank", rank, dict(ok=ok, same=same, moved=moved, unused=unused, bn_differs=not torch.equal(rms[0], rms[1])), flush=True)
    dist.destroy_process_group()


def test_gan_trainer_step_world_size_2_keeps_replicas_identical():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_trainer_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
