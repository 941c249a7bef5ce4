"""Paired-sweep kernel with 2 / 3 / 4 epilogue groups (SRCGAN_B200_SWEEP_GROUPS) on the step's shapes at 64 x 256 x 256:
the kernel is bound by its epilogue on the thin layers (one warp needs ~2 000 clocks per [32 lanes][32 channels] block)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from srcgan_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
DEV = "cuda:0"
g = torch.Generator(device=DEV).manual_seed(1)
xb = torch.randn((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV, generator=g)
yb = torch.empty((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV)
r2b = torch.randn((n, 256, 256, 64), dtype=torch.bfloat16, device=DEV, generator=g)
bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, 256, 256, 1), dtype=torch.int64, device=DEV, generator=g).to(torch.int32)
cases = [
    ("64->64 lrelu (HRconv)", 64, 64, dict(act=0.2)),
    ("64->64 + r1 (trunk_conv + fea)", 64, 64, dict(r1=ops.Slice(r2b), beta1=1.0)),
    ("192->64 *0.2 + x (conv5)", 192, 64, dict(alpha=0.2, r1=ops.Slice(xb, 0, 64), beta1=1.0)),
    ("192->64 conv5 of RDB3 (two residuals)", 192, 64, dict(alpha=0.04, r1=ops.Slice(xb, 0, 64), beta1=0.2, r2=ops.Slice(r2b), beta2=1.0)),
    ("64->32 lrelu + signbits", 64, 32, dict(act=0.2, signbits=bits)),
    ("160->32 maskbits (dgrad)", 160, 32, dict(maskbits=bits, mask_slope=0.2)),
]
for name, cin, cout, ep in cases:
    x = ops.Slice(xb, 0, cin)
    y = ops.Slice(yb, 64 if cout == 32 else 0, cout)
    wp = ops.pack_weights(torch.randn(cout, cin, 3, 3, device=DEV, generator=g) * 0.05, ops.WL_TC, torch.bfloat16)
    b = torch.randn(cout, device=DEV, generator=g)
    for _ in range(3):
        ops.conv_fprop(x, wp, b, y, 3, 1, 1, engine=ops.ENGINE_TC, **ep)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ops.conv_fprop(x, wp, b, y, 3, 1, 1, engine=ops.ENGINE_TC, **ep)
    e.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 20
    print("%-40s %.4f ms  %5.0f TFLOP/s" % (name, ms, 2.0 * n * 65536 * cin * cout * 9 / ms / 1e9), flush=True)
