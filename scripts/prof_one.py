"""One fprop + one wgrad launch of the dominant dense-block shape, for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srcgan_b200 import ops
DEV = "cuda:0"
n, h, w = 16, 256, 256
cases = [(192, 64), (64, 32), (64, 64)]
for cin, cout in cases:
    x = ops.Slice(torch.randn((n, h, w, 192), dtype=torch.bfloat16, device=DEV), 0, cin)
    y = ops.Slice(torch.empty((n, h, w, 192), dtype=torch.bfloat16, device=DEV), 0, cout)
    wt = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05
    b = torch.randn(cout, device=DEV)
    wp = ops.pack_weights(wt, ops.WL_TC, torch.bfloat16)
    dw = torch.empty(cout, cin, 3, 3, device=DEV)
    for _ in range(3):
        ops.conv_fprop(x, wp, b, y, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC)
        ops.conv_wgrad(x, y, dw, None, 3, 1, 1, engine=ops.ENGINE_TC)
torch.cuda.synchronize()
print("ok")
