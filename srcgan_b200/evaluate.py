"""Eval sweep on the device: generator inference + the four metrics of src/metrics.py with ONE
device->host read per tile instead of four ``.item()`` syncs (reference loop: src/testCas.py:65-103,
src/visCas.py:114-141; file I/O and the CSV are the caller's business).

The reference builds an autograd graph for every eval forward (no ``torch.no_grad``); here
inference runs under ``no_grad`` so no activation is retained."""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Tuple

import torch

from . import metrics as M

_EVALUATORS = (M.MSE(), M.PSNR(), M.AE(), M.SSIM())


def metrics_on_device(pred: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:
    """(4,) fp32 device tensor [MSE, PSNR, AE (mean over the batch), SSIM]: ONE kernel launch (csrc/metrics.cu
    eval_metrics_k), no host synchronisation - SSIM's data-range heuristic (metrics.py:102-111) runs on the device."""
    from . import ops
    return ops.eval_metrics(pred, truth)[:4]


def per_image_metrics(pred: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:
    """(n, 4) fp32 device tensor, one row [MSE, PSNR, AE, SSIM] per image of the batch - each image scored exactly as a
    one-tile call would score it (its own SSIM data range), ONE kernel launch for the whole batch."""
    from . import ops
    n = pred.shape[0]
    r = ops.eval_metrics(pred, truth)
    mse = r[8 + 2 * n:8 + 3 * n]
    return torch.stack([mse, 10 * torch.log10(1.0 / mse), r[8 + n:8 + 2 * n], r[8 + 3 * n:8 + 4 * n]], dim=1)


@torch.no_grad()
def evaluate(generator: Callable[[torch.Tensor], torch.Tensor],
             pairs: Iterable[Tuple[torch.Tensor, torch.Tensor]], batch: int = 1) -> Tuple[List[Dict[str, float]], Dict[str, float]]:
    """``pairs`` yields (net_input, ground_truth) device tensors (NCHW fp32).  Returns the per-item metric
    dicts and their mean, keyed by the evaluators' ``repr`` ("MSE", "PSNR", "AE", "SSIM") like the
    reference's CSV columns (testCas.py:95).

    ``batch`` > 1 sends up to that many consecutive same-shaped tiles through the generator in ONE forward (the reference's
    loop runs one tile at a time, which leaves a B200 launch-bound: 57 small launches per 128x128 tile); every tile is still
    scored on its own (per-image MSE / PSNR / AE / SSIM with its own data range), so the rows do not depend on ``batch``
    beyond bf16 rounding of nothing - inference has no cross-sample coupling in eval mode."""
    names = [repr(e) for e in _EVALUATORS]
    rows = []
    group: list = []

    def flush():
        if not group:
            return
        x = torch.cat([g[0] for g in group]) if len(group) > 1 else group[0][0]
        y = torch.cat([g[1] for g in group]) if len(group) > 1 else group[0][1]
        rows.append(per_image_metrics(generator(x), y))
        group.clear()

    for x, y in pairs:
        if group and (len(group) >= batch or x.shape != group[0][0].shape or y.shape != group[0][1].shape
                      or sum(g[0].shape[0] for g in group) + x.shape[0] > max(batch, x.shape[0])):
            flush()
        group.append((x, y))
    flush()
    if not rows:
        return [], {n: float("nan") for n in names}
    table = torch.cat(rows).cpu()                        # one D2H for the whole sweep
    per_item = [dict(zip(names, map(float, r))) for r in table]
    mean = dict(zip(names, map(float, table.mean(0))))
    return per_item, mean


@torch.no_grad()
def evaluate_cascade(net_a2c, net_c2b, pairs: Iterable[Tuple[torch.Tensor, torch.Tensor]], up: int,
                     const: bool = False):
    """The testCas.py:65-85 loop for a checkpoint pair: for each (realA gray, realB RGB) tile, the luma of realB
    is resized by 1/up (nearest, the ``interpolate`` default; the Const variants go down and back up bilinearly,
    testCasConst.py:75-76) and sent through SR -> colouriser; ``fake_BB`` is scored against realB.  Returns ``evaluate``'s result."""
    import torch.nn.functional as F

    def gen(real_b):
        luma = 0.2125 * real_b[:, :1] + 0.7154 * real_b[:, 1:2] + 0.0721 * real_b[:, 2:3]
        if const:
            lr = F.interpolate(F.interpolate(luma, scale_factor=1.0 / up, mode="bilinear"), scale_factor=up,
                               mode="bilinear")
        else:
            lr = F.interpolate(luma, scale_factor=1.0 / up)
        return net_c2b(net_a2c(lr))

    return evaluate(gen, ((b, b) for _a, b in pairs))
