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
