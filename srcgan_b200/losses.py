"""Drop-in for the reference's ``src/losses.py`` classes that the trainers in scope use
(L1Loss :95-105, MSELoss :123-133, PSNRLoss :136-147, SSIM :20-93, DSSIMLoss :170-180),
running on the fused CUDA reductions of libsrcgan_b200.so (one pass: loss value + d loss/d output)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .metrics import SSIM  # noqa: F401  (losses.SSIM is the same object as metrics.SSIM in the reference)

L1, MSE = 0, 1


class _FusedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, output, target):
        if isinstance(target, torch.Tensor) and target.shape != output.shape:
            target = target.expand_as(output)
        need_o = ctx.needs_input_grad[1]
        need_t = isinstance(target, torch.Tensor) and ctx.needs_input_grad[2]
        loss, grad = ops.loss_fwd_bwd(kind, output, target, need_o or need_t)
        ctx.g = grad.view(output.shape) if grad is not None else None
        ctx.need = (need_o, need_t)
        return loss

    @staticmethod
    def backward(ctx, go):
        g = ctx.g
        need_o, need_t = ctx.need
        d = g * go if g is not None else None
        return None, (d if need_o else None), (-d if need_t else None)


def fused_loss(kind: int, output: torch.Tensor, target) -> torch.Tensor:
    """mean |output-target| (kind 0) or mean (output-target)^2 (kind 1); target may be a python scalar."""
    return _FusedLossFn.apply(kind, output, target)


class L1Loss(nn.Module):
    def __repr__(self):
        return "L1"

    def forward(self, output, target):
        return fused_loss(L1, output, target)


class MSELoss(nn.Module):
    def __repr__(self):
        return "MSE"

    def forward(self, output, target):
        return fused_loss(MSE, output, target)


class PSNRLoss(nn.Module):
    def __repr__(self):
        return "PSNR"

    def forward(self, output, target):
        return 10 * torch.log10(1.0 / fused_loss(MSE, output, target))


class DSSIMLoss(nn.Module):
    """(1 - SSIM) / 2, differentiable through SSIM like the reference's (src/losses.py:170-180)."""

    def __init__(self):
        super().__init__()
        self.criterion = SSIM()

    def __repr__(self):
        return "DSSIM"

    def forward(self, output, target):
        return (1.0 - self.criterion(output, target)) / 2.0
