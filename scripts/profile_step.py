"""torch.profiler view of one batch-64 step: top CUDA kernels by device time."""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
from oracle import srcgan_oracle as O
from srcgan_b200 import nn as snn, trainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
snn.set_precision("bf16")
opt = trainer.params(); opt.device = torch.device("cuda:0"); opt.mode, opt.net = "x4", "1"
m = trainer.SRCycleGAN(opt)
st = O.default_states(0)
for n in ("G_A", "G_B", "D_A", "D_B"):
    getattr(m, "net" + n).load_state_dict(st[n])
rb = torch.rand(B, 3, 256, 256, device="cuda"); ra = F.interpolate(rb, scale_factor=0.25)
for _ in range(2):
    m.optimize_parameters(ra, rb)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    m.optimize_parameters(ra, rb)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total) for e in prof.key_averages()]
rows = [r for r in rows if r[2] > 0]
tot = sum(r[2] for r in rows)
rows.sort(key=lambda r: -r[2])
print("total device time (us):", tot)
for k, c, t in rows[:40]:
    print("%-80s %6d %10.0f %6.2f%%" % (k[:80], c, t, 100 * t / tot))
