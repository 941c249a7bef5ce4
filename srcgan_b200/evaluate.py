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


@torch.no_grad()
def evaluate(generator: Callable[[torch.Tensor], torch.Tensor],
             pairs: Iterable[Tuple[torch.Tensor, torch.Tensor]]) -> Tuple[List[Dict[str, float]], Dict[str, float]]:
    """``pairs`` yields (net_input, ground_truth) device tensors (NCHW fp32).  Returns the per-item metric
    dicts and their mean, keyed by the evaluators' ``repr`` ("MSE", "PSNR", "AE", "SSIM") like the
    reference's CSV columns (testCas.py:95)."""
    names = [repr(e) for e in _EVALUATORS]
    rows = []
    for x, y in pairs:
        out = generator(x)
        rows.append(metrics_on_device(out, y))
    if not rows:
        return [], {n: float("nan") for n in names}
    table = torch.stack(rows).cpu()                      # one D2H for the whole sweep
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
