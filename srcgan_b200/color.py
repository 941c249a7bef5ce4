"""Device-side RGB <-> LAB with the reference's normalisation (dataset.py:154-157, utils.py:22-26)."""
from __future__ import annotations

import torch

from . import ops


def rgb2lab(rgb: torch.Tensor, normalised: bool = True) -> torch.Tensor:
    """(B,3,H,W) RGB in [0,1] -> LAB; normalised: L/100, (a,b+128)/255 (what G2LAB feeds the nets)."""
    return ops.rgb2lab(rgb, normalised)


def lab2rgb(lab: torch.Tensor, normalised: bool = True) -> torch.Tensor:
    return ops.lab2rgb(lab, normalised)


def tensor2img(t: torch.Tensor, mode: str = "RGB") -> torch.Tensor:
    """uint8 (3,H,W) image of the first batch element, like utils.tensor2img before its cv2.resize."""
    x = t[:1].detach()
    if mode != "RGB":
        x = lab2rgb(x, True)
    if x.shape[1] == 1:
        x = x.expand(-1, 3, -1, -1)
    return (x[0] * 255).to(torch.uint8)      # truncation, as .astype(uint8)
