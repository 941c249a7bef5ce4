"""Engine selection: which kernel family runs a given convolution.

SIMT  - FFMA implicit GEMM, fp32 accumulate; the fp32 parity mode and the thin / strided layers.
TC    - TMA + tcgen05/TMEM implicit GEMM (bf16); 3x3 stride-1 pad-1 convs with wide channels.
"""
from __future__ import annotations

import os

import torch

from ._lib import ENGINE_SIMT, ENGINE_TC, WL_RSCK, WL_TC, WL_TC_S2

_force_simt = os.environ.get("SRCGAN_B200_ENGINE", "auto").lower() == "simt"


def force_simt(flag: bool) -> None:
    global _force_simt
    _force_simt = bool(flag)


def tc_fprop_supported(cin: int, cout: int, k: int, stride: int, upsample: bool, dtype, h: int, w: int) -> bool:
    """Mirror of conv_tc_supported() in csrc/conv_tc.cu (the C side re-checks alignment and refuses loudly)."""
    if dtype != torch.bfloat16 or upsample or k not in (1, 3, 4) or stride not in (1, 2):
        return False
    if cout <= 16:                     # thin outputs (64->3 image conv): 3x3 stride-1 halo kernel, plain epilogue
        return k == 3 and stride == 1
    return cout in (32, 64, 128, 256)  # any cin: TMA zero-fills the K padding (thin buffers have ld = 8)


def tc_dgrad_s2_supported(cin: int, cout: int, k: int, stride: int, pad: int, dtype) -> bool:
    """Mirror of conv_dgrad_tc_supported(): stride-2 dgrad as four output-phase convolutions."""
    return (not _force_simt and dtype == torch.bfloat16 and stride == 2 and pad == 1 and k in (3, 4) and cout >= 64
            and cout % 16 == 0 and cin in (32, 64, 128, 256))


def tc_wgrad_supported(cin: int, cout: int, k: int, stride: int, upsample: bool, dtype, h: int, w: int) -> bool:
    """Mirror of conv_wgrad_tc_supported() in csrc/conv_tc.cu."""
    if dtype != torch.bfloat16 or upsample or k not in (1, 3, 4) or stride not in (1, 2):
        return False
    if k == 3 and stride == 1:         # halo wgrad kernel: thin sides allowed (buffers have ld = 8)
        return cout <= 32 or cout % 64 == 0
    return cin >= 16 and cin % 8 == 0 and cout >= 16 and cout % 8 == 0


def sweep_bits_supported(cin: int, cout: int, k: int, stride: int, pad: int, dtype, h: int, w: int) -> bool:
    """True if conv_fprop_tc() will run this layer on the paired-sweep kernel (mirror of sweep2_cg() in csrc/conv_tc.cu),
    the only kernel that reads / writes the packed LeakyReLU masks; the C side refuses loudly if this mirror is wrong."""
    if _force_simt or dtype != torch.bfloat16 or k != 3 or stride != 1 or pad != 1 or cout not in (32, 64):
        return False
    if h < 96 and w < 96:
        return False
    if any(os.environ.get(v) for v in ("SRCGAN_B200_NO_SWEEP2", "SRCGAN_B200_SWEEP2_CG", "SRCGAN_B200_NO_MASKBITS")):
        return False
    nchunks = (cin + 63) // 64
    fixed = nchunks * (3 * (3 * cout * 128 // 2)) + 2048 + 1024          # the pair's half of the stacked weights + aux
    ring = min(10, (227 * 1024 - fixed) // 17408)
    return ring >= nchunks + 2


def select(cin, cout, k, stride, upsample, dtype, h, w):
    """-> (engine, fprop weight layout); (h, w) are the OUTPUT spatial dims."""
    if not _force_simt and dtype == torch.bfloat16 and tc_fprop_supported(cin, cout, k, stride, upsample, dtype, h, w):
        if cin <= 4 and cout % 64 == 0 and not (k == 3 and stride == 1) and not os.environ.get("SRCGAN_B200_NO_THIN_TILED"):
            # image -> features through a 4x4 (stride-2) filter, gradient of a 256 -> 1 patch-logit layer: the FFMA kernel
            # thin_in_tiled (csrc/conv_simt.cu) beats the tcgen05 GEMM with K zero-padded from <= 4 to 64 channels per tap
            return ENGINE_SIMT, WL_RSCK
        return ENGINE_TC, (WL_TC if stride == 1 else WL_TC_S2)
    return ENGINE_SIMT, WL_RSCK


def select_wgrad(cin, cout, k, stride, upsample, dtype, h, w):
    if not _force_simt and dtype == torch.bfloat16 and tc_wgrad_supported(cin, cout, k, stride, upsample, dtype, h, w):
        return ENGINE_TC
    return ENGINE_SIMT
