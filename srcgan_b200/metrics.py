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


class SSIM(object):
    def __init__(self, des="structural similarity index"):
        self.des = des

    def __repr__(self):
        return "SSIM"

    def __call__(self, y_pred, y_true, w_size=11, size_average=True, full=False):
        if w_size != 11 or full:
            raise NotImplementedError("srcgan_b200.metrics.SSIM: only w_size=11, full=False (what the reference's "
                                      "scripts use) is implemented")
        lo, hi = ops.minmax(y_pred).tolist()              # data-range heuristics, metrics.py:102-111
        L = (255 if hi > 128 else 1) - (-1 if lo < -0.5 else 0)
        n, c, h, w = y_pred.shape
        sums = ops.ssim_sums(y_pred, y_true, float(L))
        count = c * (h - w_size + 1) * (w - w_size + 1)
        if size_average:
            return sums.sum() / (n * count)
        return sums / count
