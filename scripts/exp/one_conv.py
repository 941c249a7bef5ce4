"""One TC fprop launch of a chosen shape (for in-kernel instrumentation experiments): one_conv.py cin cout [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from srcgan_b200 import ops
cin, cout = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16
DEV = "cuda:0"
x = ops.Slice(torch.randn((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV), 0, cin)
y = ops.Slice(torch.empty((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV), 0, cout)
wp = ops.pack_weights(torch.randn(cout, cin, 3, 3, device=DEV) * 0.05, ops.WL_TC, torch.bfloat16)
b = torch.randn(cout, device=DEV)
for _ in range(2):
    ops.conv_fprop(x, wp, b, y, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC)
torch.cuda.synchronize()
