"""Device-side RGB <-> LAB with the reference's normalisation (dataset.py:154-157, utils.py:22-26)."""
from __future__ import annotations

import torch

from . import ops


def rgb2lab(rgb: torch.Tensor, normalised: bool = True) -> torch.Tensor:
    """(B,3,H,W) RGB in [0,1] -> LAB; normalised: L/100, (a,b+128)/255 (what G2LAB feeds the nets)."""
    return ops.rgb2lab(rgb, normalised)


def lab2rgb(lab: torch.Tensor, normalised: bool = True) -> torch.Tensor:
    return ops.lab2rgb(lab, normalised)


def image_to_lab(img_u8: torch.Tensor) -> torch.Tensor:
    """``Basic._arr2lab`` (dataset.py:148-159) on the device: uint8 (N,H,W,3) or (H,W,3) image -> normalised LAB
    float32 (N,3,H,W).  float64 arithmetic like scikit-image, bit-compatible with the reference's example tiles."""
    return ops.rgb2lab_u8(img_u8 if img_u8.dim() == 4 else img_u8.unsqueeze(0))


def lab_to_image(lab: torch.Tensor) -> torch.Tensor:
    """``Basic._lab2img`` (dataset.py:94-104): normalised LAB float32 (N,3,H,W) -> uint8 (N,H,W,3), truncating."""
    return ops.lab2rgb_u8(lab)


def tensor2img(t: torch.Tensor, mode: str = "RGB") -> torch.Tensor:
    """uint8 (3,H,W) image of the first batch element, like utils.tensor2img (utils.py:15-29) before its cv2.resize."""
    x = t[:1].detach()
    if mode != "RGB":
        return ops.lab2rgb_u8(x)[0].permute(2, 0, 1)       # exact float64 path + truncation
    if x.shape[1] == 1:
        x = x.expand(-1, 3, -1, -1)
    return (x[0] * 255).to(torch.uint8)      # truncation, as .astype(uint8)
