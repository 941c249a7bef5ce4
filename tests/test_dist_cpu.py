"""Data-parallel host logic on CPU: world size 2, gloo backend (the N>1 path without GPUs)."""
import os
import socket

import pytest
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
    handles = sdist.attach([opt])                              # no bucketed nets: the flat-copy path
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


class _HostNet:
    """Built lazily (needs srcgan_b200.nn): a two-layer net on the REAL _NetBase / _NetFn / _GradBucket machinery whose
    kernels are torch CPU ops, so the bucket-view and overlapped-reducer logic runs under gloo without a GPU."""

    @staticmethod
    def make():
        from srcgan_b200 import nn as snn

        class HostNet(snn._NetBase):
            _host_test_double = True

            def __init__(self):
                super().__init__()
                self.a = torch.nn.Linear(4, 5)
                self.b = torch.nn.Linear(5, 3, bias=False)

            def forward(self, x):
                return snn._NetFn.apply(self, x, *self.parameters())

            def _forward_impl(self, x, st):
                h = x @ self.a.weight.detach().t() + self.a.bias.detach()
                st["x"], st["h"] = x, h
                return h @ self.b.weight.detach().t()

            def _backward_impl(self, st, g, sink, want, need_dx):
                W = lambda p: want.get(id(p), False)
                for p, val in ((self.b.weight, g.t() @ st["h"]),):
                    t, acc = sink.slot(p, W(p))
                    if t is not None:
                        t.add_(val) if acc else t.copy_(val)
                gh = g @ self.b.weight.detach()
                for p, val in ((self.a.weight, gh.t() @ st["x"]), (self.a.bias, gh.sum(0))):
                    t, acc = sink.slot(p, W(p))
                    if t is not None:
                        t.add_(val) if acc else t.copy_(val)
                return gh @ self.a.weight.detach() if need_dx else None

        return HostNet()


def _bucket_worker(rank, world, port, out, overlap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from srcgan_b200 import dist as sdist
    sdist.init_from_env(backend="gloo")
    torch.manual_seed(rank)
    net = _HostNet.make()
    extra = torch.nn.Linear(3, 1)                              # a plain torch module in the same optimizer: "loose" grads
    sdist.broadcast_module_state([net, extra])
    ref = _HostNet.make()                                      # autograd reference of the same maths, plain torch
    ref.load_state_dict(net.state_dict())
    opt = torch.optim.SGD(list(net.parameters()) + list(extra.parameters()), lr=0.1)
    red = sdist.attach([opt], [net], overlap=overlap)
    assert isinstance(red, sdist.BucketReducer)
    for it in range(3):
        g = torch.Generator().manual_seed(10 * it + rank)
        x1, x2 = torch.rand(6, 4, generator=g), torch.rand(6, 4, generator=g)
        opt.zero_grad()
        # the network runs TWICE in one backward pass (like G_A / G_B in loss_G.backward()): second node adds in place
        loss = extra(net(x1)).square().mean() + 2.0 * extra(net(x2)).square().mean()
        loss.backward()
        b = net.grad_bucket()
        assert all(b.holds(p) for p in net.parameters())       # param.grad IS the bucket view
        # plain-torch reference of the local gradient
        ws = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
        f = lambda x: (x @ ws["a.weight"].t() + ws["a.bias"]) @ ws["b.weight"].t()
        e = lambda h: h @ extra.weight.detach().t() + extra.bias.detach()
        (e(f(x1)).square().mean() + 2.0 * e(f(x2)).square().mean()).backward()
        local = {k: v.grad.clone() for k, v in ws.items()}
        launched = red.launched
        opt.step()                                             # pre-hook joins (overlap) or launches + joins the all-reduce
        if overlap:
            assert launched == it + 1 and red.launched == it + 1 and red.deferred == 0   # launched by the ready hook
        else:
            assert red.launched == 0 and red.deferred == it + 1 and red.fallbacks == it + 1   # (fallbacks: the loose grads)
        for k, p in net.named_parameters():
            parts = [torch.zeros_like(local[k]) for _ in range(world)]
            dist.all_gather(parts, local[k])
            assert torch.allclose(p.grad, sum(parts) / world, atol=1e-6), (it, k)
        eg = [torch.zeros_like(extra.weight.grad) for _ in range(world)]
        dist.all_gather(eg, extra.weight.grad)
        assert torch.equal(eg[0], eg[1])                       # loose gradients were averaged by the fallback path
    ws_ = [torch.zeros_like(net.a.weight) for _ in range(world)]
    dist.all_gather(ws_, net.a.weight.detach())
    assert torch.equal(ws_[0], ws_[1])                         # replicas in lock-step after three steps
    # frozen parameters (backward_G phase of the discriminators): no gradient, no bucket traffic, hook does not fire
    for p in net.parameters():
        p.requires_grad = False
    x = torch.rand(2, 4, requires_grad=True)
    net(x).sum().backward()
    assert x.grad is not None and red.launched == (3 if overlap else 0)
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [False, True])
def test_bucket_views_and_reducer_world2_gloo(overlap):
    """The bucket reducer in both modes: all-reduce launched in the optimizer step pre-hook (default) or from inside backward."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, out, overlap)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"


def test_gradient_accumulation_over_backward_passes_uses_the_bucket_in_place():
    """Two backward passes without zero_grad: the second adds into the bucket that p.grad already views."""
    net = _HostNet.make()
    x1, x2 = torch.rand(3, 4), torch.rand(3, 4)
    net(x1).sum().backward()
    g1 = {k: p.grad.clone() for k, p in net.named_parameters()}
    net(x2).sum().backward()
    ref = _HostNet.make()
    ref.load_state_dict(net.state_dict())
    for p in ref.parameters():
        p.grad = None
    import os as _os
    _os.environ["SRCGAN_B200_NO_GRAD_BUCKET"] = "1"
    try:
        ref(x1).sum().backward()
        ref(x2).sum().backward()
    finally:
        del _os.environ["SRCGAN_B200_NO_GRAD_BUCKET"]
    for (k, p), q in zip(net.named_parameters(), ref.parameters()):
        assert net.grad_bucket().holds(p) and not ref.grad_bucket().holds(q)
        assert torch.allclose(p.grad, q.grad, atol=1e-6), k
        assert not torch.allclose(p.grad, g1[k])
    # zero_grad(set_to_none) then a new pass: written fresh, not added
    for p in net.parameters():
        p.grad = None
    net(x1).sum().backward()
    for k, p in net.named_parameters():
        assert torch.allclose(p.grad, g1[k], atol=1e-6), k


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
