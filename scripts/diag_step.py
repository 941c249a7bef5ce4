"""Diagnostic: gradients of one backward_G / backward_D of the CUDA fp32 path vs the CPU oracle (fp32 and fp64)."""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import srcgan_oracle as O
from srcgan_b200 import nn as snn, trainer

snn.set_precision("fp32")
DEV = "cuda:0"
B, LR = int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 16


def errs(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30)), float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


states = O.default_states(0)
opt = trainer.params(); opt.device = torch.device(DEV); opt.mode = "x4"; opt.net = "1"
m = trainer.SRCycleGAN(opt)
for name in ("G_A", "G_B", "D_A", "D_B"):
    getattr(m, "net" + name).load_state_dict(states[name], strict=True)
real_A, real_B = O.synthetic_batch(B, lr=LR, scale=4, seed=1234)

def gpu_phase():
    m.forward(real_A.to(DEV), real_B.to(DEV))
    m.set_requires_grad([m.netD_A, m.netD_B], False)
    m.optimizer_G.zero_grad()
    m.backward_G()
    m.set_requires_grad([m.netD_A, m.netD_B], True)
    m.optimizer_D.zero_grad()
    random.seed(5); m.backward_D_A(); m.backward_D_B()

def ref_phase(dt):
    st = {n: {k: (v.to(dt) if v.is_floating_point() else v.clone()) for k, v in s.items()} for n, s in states.items()}
    r = O.CycleGANStepOracle(st)
    r.forward(real_A.to(dt), real_B.to(dt))
    r._set_requires_grad(r.D_A, False); r._set_requires_grad(r.D_B, False)
    r.backward_G()
    r._set_requires_grad(r.D_A, True); r._set_requires_grad(r.D_B, True)
    random.seed(5)
    r._backward_D(r.netD_A, r.real_B, r.fake_B_pool.query(r.fake_B))
    r._backward_D(r.netD_B, r.real_A, r.fake_A_pool.query(r.fake_A))
    return r

gpu_phase()
r64, r32 = ref_phase(torch.float64), ref_phase(torch.float32)
for name in ("G_A", "G_B", "D_A", "D_B"):
    net = getattr(m, "net" + name)
    rows = []
    for k, p in net.named_parameters():
        g64 = getattr(r64, name)[k].grad
        if g64 is None:
            assert p.grad is None, k
            continue
        rows.append((errs(p.grad, g64), errs(getattr(r32, name)[k].grad, g64), k))
    rows.sort(reverse=True)
    print(name, "worst by L2 (gpu-vs-64 [l2,max] | cpu32-vs-64 [l2,max])")
    for (g, c, k) in rows[:4]:
        print("   %.2e %.2e | %.2e %.2e  %s" % (g[0], g[1], c[0], c[1], k))
    import statistics
    print("   median gpu l2 %.2e   median cpu32 l2 %.2e" % (statistics.median(r[0][0] for r in rows), statistics.median(r[1][0] for r in rows)))
for n in ("fake_B", "fake_A", "recl_A", "recl_B", "iden_A", "iden_B"):
    print(n, "gpu-vs-64 max %.2e   cpu32-vs-64 max %.2e" % (errs(getattr(m, n), getattr(r64, n))[1], errs(getattr(r32, n), getattr(r64, n))[1]))
