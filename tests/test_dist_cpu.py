"""Data-parallel host logic on CPU: world size 2, gloo backend (the N>1 path without GPUs)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from srcgan_b200 import dist as sdist
    sdist.init_from_env(backend="gloo")
    assert sdist.world() == world and sdist.rank() == rank
    torch.manual_seed(rank)                                    # different replicas ...
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4))
    sdist.broadcast_module_state([net])                        # ... start from rank 0's state
    w0 = net[0].weight.detach().clone()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    handles = sdist.attach([opt])
    assert len(handles) == 1
    g = torch.Generator().manual_seed(100 + rank)              # rank-local shard of the batch
    x = torch.rand(2, 3, 8, 8, generator=g)
    net(x).square().mean().backward()
    local_grad = net[0].weight.grad.detach().clone()
    opt.step()                                                 # pre-hook averages the gradients
    gathered = [torch.zeros_like(local_grad) for _ in range(world)]
    dist.all_gather(gathered, local_grad)
    mean_grad = sum(gathered) / world
    assert torch.allclose(net[0].weight.grad, mean_grad, atol=1e-7)
    ws = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(ws, net[0].weight.detach())
    assert torch.equal(ws[0], ws[1])                           # replicas stay in lock-step
    gw = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(gw, w0)
    assert torch.equal(gw[0], gw[1])
    # allreduce_mean_ on a ragged list of tensors
    ts = [torch.full((3,), float(rank)), torch.full((2, 2), float(rank + 1))]
    sdist.allreduce_mean_(ts)
    assert torch.allclose(ts[0], torch.full((3,), 0.5)) and torch.allclose(ts[1], torch.full((2, 2), 1.5))
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


def test_gradient_averaging_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"


def test_single_process_is_a_noop():
    from srcgan_b200 import dist as sdist
    assert sdist.world() == 1 and sdist.rank() == 0
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.1)
    assert sdist.attach([opt]) == []
    t = [torch.ones(2)]
    sdist.allreduce_mean_(t)
    assert torch.equal(t[0], torch.ones(2))
