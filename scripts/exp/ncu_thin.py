"""The FFMA thin-layer kernels on their hot shapes (batch 64), two warm-up launches and one measured launch each - meant to run
under `ncu --set full -k regex:'thin_in_tiled|thin_wgrad7'` (scripts/gpu/r2_ba.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from srcgan_b200 import ops

DEV = "cuda:0"
bf = torch.bfloat16
for (n, h, w, cin, cout, k, s, p) in ((64, 256, 256, 3, 64, 7, 2, 3), (64, 256, 256, 3, 64, 4, 2, 1)):
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    x = ops.Slice(torch.randn((n, h, w, 8), dtype=bf, device=DEV), 0, cin)
    y = ops.Slice(torch.empty((n, ho, wo, cout), dtype=bf, device=DEV))
    wp = ops.pack_weights(torch.randn(cout, cin, k, k, device=DEV) * 0.05, ops.WL_RSCK, bf)
    b = torch.randn(cout, device=DEV)
    for _ in range(3):
        ops.conv_fprop(x, wp, b, y, k, s, p, act=0.2)
    if k == 7:
        dw = torch.empty(cout, cin, k, k, device=DEV)
        gy = ops.Slice(torch.randn((n, ho, wo, cout), dtype=bf, device=DEV))
        for _ in range(3):
            ops.conv_wgrad(x, gy, dw, None, k, s, p)
    torch.cuda.synchronize()
    print("done", n, h, w, cin, cout, k, s, flush=True)
