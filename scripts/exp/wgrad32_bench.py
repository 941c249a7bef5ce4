"""The 32 -> 32 remainder launches of the paired dense-block weight gradients (x_k slice x dZ_(k+1) slice of the 192-channel
concat buffers) at 64 x 256 x 256: HBM-bound, 64 useful bytes per pixel and operand."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from srcgan_b200 import ops  # noqa: E402

DEV = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
X = torch.randn((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV)
D = torch.randn((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV)
for c0, d0 in ((64, 96), (128, 160)):
    x, dy = ops.Slice(X, c0, 32), ops.Slice(D, d0, 32)
    dw = torch.zeros(32, 160, 3, 3, device=DEV)
    for _ in range(3):
        ops.conv_wgrad_split(x, dy, (dw, c0, None), None, 32)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ops.conv_wgrad_split(x, dy, (dw, c0, None), None, 32)
    e.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 20
    print("wgrad 32->32 (x at channel %d, dY at %d): %.4f ms  %.0f GB/s of useful bytes" % (c0, d0, ms, n * 65536 * 128 / ms / 1e6))
