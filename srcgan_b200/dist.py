"""Data parallelism for the G+D step: one process per GPU (torchrun), full replicas, local batches,
gradient averaging with NCCL over NVLink.

The batch shards by patch: every rank runs the whole step on its own patches, BatchNorm
normalises with rank-local batch statistics (no SyncBN - the reference has none to mirror) and the running
mean / variance buffers therefore drift apart per rank (torch DDP would re-broadcast rank 0's every forward;
here ``sync_buffers`` averages them on demand, and ``checkpoint.save_state`` callers should call it first) and
the only exchange is the all-reduce of the fp32 gradients right before each optimizer step.  The
trainers construct their optimizers themselves (train.py:191-192), so the exchange is attached as an
optimizer *step pre-hook*: nothing in the trainer has to change.  ~27 MB per step over NVSwitch is
<0.1% of a step; it is issued on the compute stream right after the last wgrad.
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
    """In-place mean over ranks of a list of same-dtype tensors through one flat bucket."""
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


def attach(optimizers: Iterable[torch.optim.Optimizer]):
    """Average gradients over ranks before every ``optimizer.step()``.  Returns the hook handles."""
    if world() == 1:
        return []
    return [opt.register_step_pre_hook(_step_pre_hook) for opt in optimizers]


def make_data_parallel(model) -> list:
    """``model``: a trainer.SRCycleGAN (or the reference's own).  Broadcast + hooks."""
    broadcast_module_state([model.netG_A, model.netG_B, model.netD_A, model.netD_B])
    return attach([model.optimizer_G, model.optimizer_D])
