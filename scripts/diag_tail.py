"""Diagnostic: per-stage gradient error of the RDDBNetB tail (fp32 CUDA path vs fp64 oracle)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import srcgan_oracle as O
from srcgan_b200 import nn as snn, ops
from srcgan_b200.nn import _GradSink

snn.set_precision("fp32")
DEV = "cuda:0"


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


mode = sys.argv[1] if len(sys.argv) > 1 else "x4"
sd = O.init_rddbnet_b(11)
x = torch.rand(2, 3, 16, 16, generator=torch.Generator().manual_seed(101))
net = snn.RDDBNetB(3, 3, 64, nb=3, mode=mode)
net.load_state_dict(sd); net.to(DEV)
st = {"_debug": {}}
with torch.no_grad():
    y = net._forward_impl(x.to(DEV), st)
pr = torch.randn(y.shape, generator=torch.Generator().manual_seed(7))
want = {id(p): True for p in net.parameters()}
sink = _GradSink()
with torch.no_grad():
    dx = net._backward_impl(st, pr.to(DEV), sink, want, True)

for dt in (torch.float64, torch.float32):
    s = {k: v.to(dt) for k, v in sd.items()}
    xr = x.to(dt)
    fea = F.conv2d(xr, s["conv_first.weight"], s["conv_first.bias"], padding=1)
    fea2 = (fea + O._trunk(s, fea, 3)).detach().requires_grad_(True)
    acts = [fea2]
    cur = fea2
    stages = [("upconv1", True), ("upconv2", True)] if mode == "x4" else [("upconv1", True), ("upconv1", False)]
    stages += [("HRconv", False)] * 8
    pre = []
    for name, up in stages:
        z = F.conv2d(F.interpolate(cur, scale_factor=2, mode="nearest") if up else cur, s[name + ".weight"], s[name + ".bias"], padding=1)
        z.retain_grad()
        pre.append(z)
        cur = F.leaky_relu(z, 0.2)
        acts.append(cur)
    out = F.conv2d(cur, s["conv_last.weight"], s["conv_last.bias"], padding=1)
    (out * pr.to(dt)).sum().backward()
    print("dtype", dt, "out err", relerr(y, out))
    gpu_acts = st["acts"]
    for i in range(len(stages) - 1, -1, -1):
        # gz{i} on the GPU = gradient w.r.t. pre-activation of stage i-1 (or fea2 for i=0)
        ref = pre[i - 1].grad if i > 0 else fea2.grad
        a_gpu = gpu_acts[i + 1].view().float().permute(0, 3, 1, 2)
        nflip = int(((a_gpu.cpu() > 0) != (acts[i + 1].detach().float() > 0)).sum())
        print("  stage %2d  act err %.2e  sign flips %d   grad(gz%d) err %.2e" % (
            i, relerr(a_gpu, acts[i + 1]), nflip, i, relerr(st["_debug"]["gz%d" % i], ref)))
