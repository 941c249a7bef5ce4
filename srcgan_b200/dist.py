"""Data parallelism for the G+D step: one process per GPU (torchrun), full replicas, local batches,
gradient averaging with NCCL over NVLink.

The batch shards by patch: every rank runs the whole step on its own patches, BatchNorm
normalises with rank-local batch statistics (no SyncBN - the reference has none to mirror) and the running
mean / variance buffers therefore drift apart per rank (torch DDP would re-broadcast rank 0's every forward;
here ``sync_buffers`` averages them on demand, and ``checkpoint.save_state`` callers should call it first) and
the only exchange is the all-reduce of the fp32 gradients right before each optimizer step.  The
trainers construct their optimizers themselves (train.py:191-192), so the exchange is attached as an
optimizer *step pre-hook*: nothing in the trainer has to change.  ~27 MB per step over NVSwitch is
<0.1% of a step (see BucketReducer for why it is NOT overlapped with the backward pass by default).
"""
from __future__ import annotations

import os
from typing import Iterable, List

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def init_from_env(backend: str | None = None) -> int:
    """Initialise the default process group from torchrun's environment; returns the local rank."""
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend)
    return local_rank


def broadcast_module_state(modules: Iterable[torch.nn.Module], src: int = 0) -> None:
    """Replicas start from rank ``src``'s parameters and buffers."""
    if world() == 1:
        return
    with torch.no_grad():
        for m in modules:
            for t in list(m.parameters()) + list(m.buffers()):
                dist.broadcast(t, src)          # through the tensor itself (not .data): bumps its version counter
            if hasattr(m, "invalidate_packs"):
                m.invalidate_packs()            # packed bf16 weight caches are derived from the parameters


def allreduce_mean_(tensors: List[torch.Tensor]) -> None:
    """In-place mean over ranks of a list of same-dtype tensors through one flat bucket (the fallback path for gradients
    that are not views of a network's bucket)."""
    n = world()
    if n == 1 or not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])          # one bucket: one NCCL call over NVLink
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.mul_(1.0 / n)
    views, off = [], 0
    for t in tensors:
        k = t.numel()
        views.append(flat[off:off + k].view_as(t))
        off += k
    torch._foreach_copy_(tensors, views)                         # one fused scatter back into the .grad tensors


class BucketReducer:
    """Bucketed gradient exchange.  Every network of ``srcgan_b200.nn`` writes its parameter gradients into
    ONE flat fp32 bucket (``nn._GradBucket``; ``param.grad`` is a view of it), so the exchange is one in-place NCCL
    ``all_reduce`` (AVG) per network - no ``torch.cat``, no copy back - joined by a stream-side wait (no host
    synchronisation) in the optimizer *step pre-hook*.

    ``overlap=False`` (default): the all-reduces are launched in the pre-hook itself, when the compute stream has nothing
    else to run.  ``overlap=True``: when the last pending backward of a network has been issued - inside
    ``loss.backward()``, on autograd's thread - its ``_grads_ready_hook`` fires and the bucket's all-reduce is launched
    right there: NCCL's stream waits for the compute stream at that point and the reduction runs beside the rest of the
    backward pass (G_B's 12.8 MB while G_A's last backward runs; D_A's while D_B's runs).  Measured on two B200s the
    overlap LOSES 13 ms of a 256 ms step (profiles/r2_bench_2gpu_*.json): 27 MB over NVLink take ~0.1 ms, there is nothing
    to hide, while the convolution kernels are persistent grids over all 74 CTA pairs with a static split of the work -
    every SM NCCL's kernels hold for a while turns one pair's share into a straggler.

    Gradients that are not bucket views at step time (a foreign ``.grad``, a module without buckets, a network whose
    backward never completed) go through the flat-copy fallback so the result is always the mean over ranks."""

    def __init__(self, nets, optimizers, overlap: bool = False):
        self.nets = [n for n in nets if hasattr(n, "grad_bucket")]
        self.pending = {}            # id(net) -> (work, writes at launch)
        self.launched = 0            # all-reduces launched from inside backward (overlap mode)
        self.deferred = 0            # all-reduces launched in the step pre-hook
        self.fallbacks = 0
        self.overlap = bool(overlap)
        self.avg = dist.get_backend() == "nccl"
        for n in self.nets:
            n.__dict__["_grads_ready_hook"] = self._ready
        self.handles = [opt.register_step_pre_hook(self._pre_step) for opt in optimizers]
        self._owner = {}
        for n in self.nets:
            for p in n.parameters():
                self._owner[id(p)] = n

    def _ready(self, net) -> None:
        b = net.grad_bucket()
        if b.flat is None or not self.overlap:
            return
        if id(net) in self.pending:
            raise RuntimeError("srcgan_b200.dist: a second backward pass wrote gradients after this network's all-reduce "
                               "was launched and before optimizer.step() - gradient accumulation over several backward "
                               "passes needs attach(..., overlap=False)")
        work = dist.all_reduce(b.flat, op=dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM, async_op=True)
        self.pending[id(net)] = (work, b.writes)
        self.launched += 1

    def _pre_step(self, optimizer, args, kwargs) -> None:
        n = world()
        params = [p for g in optimizer.param_groups for p in g["params"] if p.grad is not None]
        nets, loose = {}, []
        for p in params:
            net = self._owner.get(id(p))
            if net is not None and net.grad_bucket().holds(p):
                nets[id(net)] = net
            else:
                loose.append(p.grad)
        for net in nets.values():
            b = net.grad_bucket()
            got = self.pending.pop(id(net), None)
            if got is not None and got[1] != b.writes:
                raise RuntimeError("srcgan_b200.dist: gradients were written after the all-reduce was launched")
            if got is None:                                      # not launched from inside backward: reduce here
                work = dist.all_reduce(b.flat, op=dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM, async_op=True)
                if self.overlap:
                    self.fallbacks += 1                          # the ready hook should have fired
                else:
                    self.deferred += 1
            else:
                work = got[0]
            work.wait()                                          # the compute stream waits for NCCL's; the host does not
            if not self.avg:
                b.flat.mul_(1.0 / n)
            net.__dict__["_pending_bw"] = 0
        if loose:
            self.fallbacks += 1
            allreduce_mean_(loose)

    def detach(self) -> None:
        for h in self.handles:
            h.remove()
        for n in self.nets:
            n.__dict__["_grads_ready_hook"] = None


def sync_buffers(modules: Iterable[torch.nn.Module]) -> None:
    """Average the floating-point buffers (BatchNorm running statistics) over ranks - call before checkpointing."""
    n = world()
    if n == 1:
        return
    with torch.no_grad():
        for m in modules:
            for b in m.buffers():
                if b.is_floating_point():
                    dist.all_reduce(b, op=dist.ReduceOp.SUM)
                    b.mul_(1.0 / n)


def _step_pre_hook(optimizer, args, kwargs):
    grads = [p.grad for g in optimizer.param_groups for p in g["params"] if p.grad is not None]
    allreduce_mean_(grads)


def attach(optimizers: Iterable[torch.optim.Optimizer], nets: Iterable[torch.nn.Module] = (), overlap: bool = False,
           bucketed: bool = True):
    """Average gradients over ranks before every ``optimizer.step()``.  With ``nets`` (the ``srcgan_b200.nn`` modules the
    optimizers own) the exchange is the ``BucketReducer`` (one in-place all-reduce per network's gradient bucket; launched
    in the step pre-hook, or from inside backward with ``overlap``); without them, or with ``bucketed=False``, one flat
    cat / all-reduce / copy-back inside the step pre-hook.  Returns the reducer (or the hook handles)."""
    if world() == 1:
        return []
    optimizers = list(optimizers)
    nets = list(nets)
    if bucketed and nets:
        return BucketReducer(nets, optimizers, overlap)
    return [opt.register_step_pre_hook(_step_pre_hook) for opt in optimizers]


def make_data_parallel(model, overlap: bool = False):
    """``model``: a trainer.SRCycleGAN (or the reference's own, built on the drop-in modules).  Broadcast + hooks."""
    nets = [model.netG_A, model.netG_B, model.netD_A, model.netD_B]
    broadcast_module_state(nets)
    return attach([model.optimizer_G, model.optimizer_D], nets, overlap)
