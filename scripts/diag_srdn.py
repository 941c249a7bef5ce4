import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import srcgan_oracle as O
from srcgan_b200 import nn as snn
snn.set_precision("fp32")
DEV = "cuda:0"
def l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
sd = O.init_srdn(32)
x = torch.rand(2, 1, 16, 12, generator=torch.Generator().manual_seed(302))
net = snn.SRDN(1, 1, 2); net.load_state_dict(sd); net.to(DEV)
xg = x.to(DEV).requires_grad_(True)
y = net(xg)
pr = torch.randn(y.shape, generator=torch.Generator().manual_seed(18))
(y * pr.to(DEV)).sum().backward()
s = O.as_leaf_params({k: v.double() for k, v in sd.items()})
xr = x.double().requires_grad_(True)
yr = O.srdn(s, xr); (yr * pr.double()).sum().backward()
print("out", l2(y, yr), "dx", l2(xg.grad, xr.grad))
named = dict(net.named_parameters())
for k, v in s.items():
    if v.grad is None: continue
    e = l2(named[k].grad, v.grad)
    if e > 1e-3 or k.endswith("conv5.weight") or "conv_" in k: print("%.3e %s" % (e, k))
