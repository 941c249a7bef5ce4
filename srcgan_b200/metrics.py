"""Drop-in for the reference's ``src/metrics.py`` (AE :10-33, MSE :36-50, PSNR :53-68, SSIM :71-144)
on the CUDA kernels of libsrcgan_b200.so.  Same call signatures, same ``__repr__`` strings (they
are used as CSV column names by testCas.py:95)."""
from __future__ import annotations

import torch

from . import ops


class AE(object):
    def __init__(self, des="average Angular Error"):
        self.des = des

    def __repr__(self):
        return "AE"

    def __call__(self, y_pred, y_true):
        return ops.angular_error(y_pred, y_true)          # shape (B,), degrees


class MSE(object):
    def __init__(self, des="Mean Square Error"):
        self.des = des

    def __repr__(self):
        return "MSE"

    def __call__(self, y_pred, y_true, dim=1):
        return ops.sq_err_sum(y_pred, y_true) / y_pred.numel()


class PSNR(object):
    def __init__(self, des="Peak Signal to Noise Ratio"):
        self.des = des

    def __repr__(self):
        return "PSNR"

    def __call__(self, y_pred, y_true, dim=1):
        mse = ops.sq_err_sum(y_pred, y_true) / y_pred.numel()
        return 10 * torch.log10(1 / mse)


class _SSIMFn(torch.autograd.Function):
    """mean (or per-image mean) of the SSIM map, differentiable w.r.t. the prediction (losses.py:40-93 builds it from
    conv2d calls, so autograd flows through it in the reference; here: csrc/metrics.cu ssim_bwd_*)."""

    @staticmethod
    def forward(ctx, y_pred, y_true, size_average):
        res = ops.eval_metrics(y_pred, y_true)
        n = y_pred.shape[0]
        ctx.save_for_backward(y_pred.detach(), y_true.detach(), res[4:5])
        ctx.size_average = size_average
        return res[3] if size_average else res[8:8 + n]

    @staticmethod
    def backward(ctx, go):
        y_pred, y_true, L = ctx.saved_tensors
        n, c, h, w = y_pred.shape
        count = c * (h - 10) * (w - 10)
        if ctx.size_average:
            d = ops.ssim_backward(y_pred, y_true, L, go, 1.0 / (n * count))
        else:   # per-image upstream gradients: one launch per image (the reference's scripts never use this form)
            d = torch.cat([ops.ssim_backward(y_pred[i:i + 1], y_true[i:i + 1], L, go[i], 1.0 / count) for i in range(n)])
        return d, None, None


class SSIM(object):
    def __init__(self, des="structural similarity index"):
        self.des = des

    def __repr__(self):
        return "SSIM"

    def __call__(self, y_pred, y_true, w_size=11, size_average=True, full=False):
        if w_size != 11 or full:
            raise NotImplementedError("srcgan_b200.metrics.SSIM: only w_size=11, full=False (what the reference's "
                                      "scripts use) is implemented")
        # the data-range heuristic of metrics.py:102-111 (max > 128 -> 255, min < -0.5 -> +1) runs inside the kernel:
        # no host synchronisation
        return _SSIMFn.apply(y_pred, y_true, bool(size_average))
