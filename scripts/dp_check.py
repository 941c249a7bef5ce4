#!/usr/bin/env python
"""Data-parallel correctness on real GPUs (torchrun, NCCL): the bucketed, overlapped all-reduce must give every rank exactly the
mean over ranks of the local gradients, and the replicas must stay in lock-step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py

Rank-local model A steps with ``dist.make_data_parallel``; model B (same weights, same rank-local batch, same python RNG) steps
alone; B's local gradients are then averaged with a plain all_reduce and compared with A's."""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F

from srcgan_b200 import dist as sdist, nn as snn, trainer


def main():
    local_rank = sdist.init_from_env()
    rank, world = sdist.rank(), sdist.world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    snn.set_precision("bf16")
    opt = trainer.params()
    opt.device, opt.mode, opt.net = dev, "x4", "1"
    torch.manual_seed(0)
    A = trainer.SRCycleGAN(opt)
    torch.manual_seed(0)
    B = trainer.SRCycleGAN(opt)
    reducer = sdist.make_data_parallel(A, overlap=bool(int(os.environ.get('DP_CHECK_OVERLAP', '0'))))
    for n in ("G_A", "G_B", "D_A", "D_B"):
        getattr(B, "net" + n).load_state_dict(getattr(A, "net" + n).state_dict())
    g = torch.Generator().manual_seed(100 + rank)                  # rank-local shard
    real_B = torch.rand(4, 3, 128, 128, generator=g).to(dev)
    real_A = F.interpolate(real_B, scale_factor=0.25, mode="nearest")
    report = {"rank": rank, "world": world}
    worst = 0.0
    for it in range(2):
        random.seed(it)
        A.optimize_parameters(real_A, real_B)
        random.seed(it)
        B.optimize_parameters(real_A, real_B)
        for n in ("G_A", "G_B", "D_A", "D_B"):
            na, nb = getattr(A, "net" + n), getattr(B, "net" + n)
            for (k, pa), (_k, pb) in zip(na.named_parameters(), nb.named_parameters()):
                if pb.grad is None:
                    assert pa.grad is None, (n, k)
                    continue
                if it == 0:      # same weights on both models: A's gradient must be the mean over ranks of B's local one
                    want = pb.grad.detach().clone()
                    dist.all_reduce(want)
                    want /= world
                    err = float((pa.grad - want).abs().max() / want.abs().max().clamp_min(1e-30))
                    worst = max(worst, err)
                    assert err < 1e-5, (n, k, err)
                assert na.grad_bucket().holds(pa), (n, k)          # gradients are views of the flat bucket
    # replicas in lock-step: every rank holds the same parameters after two steps
    for n in ("G_A", "G_B", "D_A", "D_B"):
        flat = torch.cat([p.detach().reshape(-1) for p in getattr(A, "net" + n).parameters()])
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), n
    report.update(worst_rel_err_vs_manual_mean=worst, overlapped_launches=reducer.launched, pre_hook_launches=reducer.deferred,
                  fallbacks=reducer.fallbacks)
    assert reducer.launched + reducer.deferred == 8 and reducer.fallbacks == 0, (reducer.launched, reducer.deferred,
                                                                                 reducer.fallbacks)   # 4 buckets x 2 steps
    if rank == 0:
        print(json.dumps(report), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
