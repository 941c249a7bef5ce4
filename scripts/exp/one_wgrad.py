"""One TC wgrad call of a chosen shape (for ncu): one_wgrad.py cin cout [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from srcgan_b200 import ops
cin, cout = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16
DEV = "cuda:0"
x = ops.Slice(torch.randn((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV), 0, cin)
g = ops.Slice(torch.randn((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV), 0, cout)
dw = torch.empty(cout, cin, 3, 3, device=DEV); db = torch.empty(cout, device=DEV)
for _ in range(2):
    ops.conv_wgrad(x, g, dw, db, 3, 1, 1, engine=ops.ENGINE_TC)
torch.cuda.synchronize()
