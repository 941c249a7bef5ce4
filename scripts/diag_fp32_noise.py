"""Diagnostic: is the fp32-mode deviation from the CPU oracle rounding noise?  Compares the CUDA fp32
path and the CPU fp32 oracle against an fp64 evaluation of the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import srcgan_oracle as O
from srcgan_b200 import nn as snn

snn.set_precision("fp32")
DEV = "cuda:0"


def relerr(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def run(mode):
    sd = O.init_rddbnet_b(11)
    x = torch.rand(2, 3, 16, 16, generator=torch.Generator().manual_seed(101))
    net = snn.RDDBNetB(3, 3, 64, nb=3, mode=mode)
    net.load_state_dict(sd); net.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    y = net(xg)
    pr = torch.randn(y.shape, generator=torch.Generator().manual_seed(7))
    (y * pr.to(DEV)).sum().backward()
    res = {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        s = O.as_leaf_params({k: v.to(dt) for k, v in sd.items()})
        xr = x.detach().clone().to(dt).requires_grad_(True)
        yr = O.rddbnet_b(s, xr, mode)
        (yr * pr.to(dt)).sum().backward()
        res[tag] = (yr.detach(), s, xr.grad)
    y64, s64, dx64 = res["f64"]
    y32, s32, dx32 = res["f32"]
    print(mode, "out   gpu-vs-64 %.2e  cpu32-vs-64 %.2e  gpu-vs-cpu32 %.2e" % (relerr(y, y64), relerr(y32, y64), relerr(y, y32)))
    print(mode, "dx    gpu-vs-64 %.2e  cpu32-vs-64 %.2e  gpu-vs-cpu32 %.2e" % (relerr(xg.grad, dx64), relerr(dx32, dx64), relerr(xg.grad, dx32)))
    worst = []
    named = dict(net.named_parameters())
    for k, v in s64.items():
        if v.grad is None:
            continue
        worst.append((relerr(named[k].grad, v.grad), relerr(s32[k].grad, v.grad), relerr(named[k].grad, s32[k].grad), k))
    worst.sort(reverse=True)
    for w in worst[:6]:
        print(mode, "grad  gpu-vs-64 %.2e  cpu32-vs-64 %.2e  gpu-vs-cpu32 %.2e  %s" % w)

for mode in ("x4", "x2"):
    run(mode)
