"""torch.profiler (CUPTI) view of ONE batch-64 G+D step: every CUDA kernel with its launch count and device time, the sum of
device time against the step's wall time (idle share), and the ATen kernels that are still on the path.
python scripts/profile_step.py [batch] > gpurun_out/profile_step.txt"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.profiler import ProfilerActivity, profile

from srcgan_b200 import nn as snn, trainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
snn.set_precision("bf16")
opt = trainer.params()
opt.device = torch.device("cuda:0")
opt.mode, opt.net = "x4", "1"
torch.manual_seed(0)
m = trainer.SRCycleGAN(opt)
rb = torch.rand(B, 3, 256, 256, device="cuda")
ra = F.interpolate(rb, scale_factor=0.25)
for _ in range(3):
    m.optimize_parameters(ra, rb)
torch.cuda.synchronize()
t0 = time.perf_counter()
m.optimize_parameters(ra, rb)
t_enq = (time.perf_counter() - t0) * 1e3          # host time to ENQUEUE the step (returns before the GPU is done)
torch.cuda.synchronize()
wall_plain = (time.perf_counter() - t0) * 1e3
print("host enqueue time of one step: %.1f ms of %.1f ms wall (the GPU is starved whenever the host is the later one)" % (t_enq, wall_plain))
if os.environ.get("PROFILE_STEP_HOST"):
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    m.optimize_parameters(ra, rb)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(30)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    m.optimize_parameters(ra, rb)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
rows = [(e.key, e.count, getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)) for e in prof.key_averages()]
rows = [r for r in rows if r[2] > 0]
tot = sum(r[2] for r in rows)
rows.sort(key=lambda r: -r[2])
ours = ("conv", "bn_", "gn_", "loss_", "ssim", "eval_metrics", "colsum", "wgrad", "pack_weights", "nchw", "nhwc", "upsample", "add_", "act_bwd",
        "rgb2lab", "lab2rgb", "thin_", "d2s", "pixel_shuffle", "minmax", "ae_kernel", "tc::", "tcw")
mine = [r for r in rows if any(k in r[0] for k in ours)]
aten = [r for r in rows if r not in mine]
print("batch %d: wall %.1f ms (%.1f ms under the profiler), kernels %d launches, device time %.1f ms = %.1f %% of the wall time" % (
    B, wall_plain, wall, sum(r[1] for r in rows), tot / 1e3, tot / 1e3 / wall_plain * 100))
print("this library: %d launches, %.1f ms;  ATen / NCCL / memcpy: %d launches, %.1f ms" % (
    sum(r[1] for r in mine), sum(r[2] for r in mine) / 1e3, sum(r[1] for r in aten), sum(r[2] for r in aten) / 1e3))
print("\n%-90s %7s %10s %7s" % ("kernel", "count", "us", "share"))
for k, c, t in rows[:45]:
    print("%-90s %7d %10.0f %6.2f%%" % (k[:90], c, t, 100 * t / tot))
print("\nATen / other kernels still on the path:")
for k, c, t in aten[:25]:
    print("%-90s %7d %10.0f %6.2f%%" % (k[:90], c, t, 100 * t / tot))

# ---- where the device idles: gaps between consecutive kernels on the timeline (same profiler run)
evs = []
for e in prof.events():
    dt = getattr(e, "device_type", None)
    if dt is not None and "CUDA" in str(dt) and e.time_range.end > e.time_range.start:
        evs.append((e.time_range.start, e.time_range.end, e.name))
evs.sort()
if evs:
    gaps = []
    end = evs[0][1]
    prev = evs[0][2]
    for s, t, name in evs[1:]:
        if s > end:
            gaps.append((s - end, prev, name))
        if t > end:
            end, prev = t, name
    total_gap = sum(g[0] for g in gaps)
    print("\ntimeline: first kernel start -> last kernel end %.1f ms; %d gaps, %.2f ms idle in total" % (
        (evs[-1][1] - evs[0][0]) / 1e3, len(gaps), total_gap / 1e3))
    for lo, hi in ((0, 2), (2, 5), (5, 20), (20, 100), (100, 1e9)):
        sel = [g[0] for g in gaps if lo <= g[0] < hi]
        print("  gaps of %5g - %5g us: %5d, %.2f ms" % (lo, hi, len(sel), sum(sel) / 1e3))
    print("largest gaps (us, after kernel -> before kernel):")
    for g in sorted(gaps, key=lambda g: -g[0])[:25]:
        print("  %8.1f  %-60s -> %s" % (g[0], g[1][:60], g[2][:60]))
    by_prev = {}
    for g in gaps:
        d = by_prev.setdefault(g[1][:70], [0, 0.0])
        d[0] += 1
        d[1] += g[0]
    print("idle time by the kernel that ran before the gap:")
    for k, (c, t) in sorted(by_prev.items(), key=lambda kv: -kv[1][1])[:15]:
        print("  %8.1f us in %4d gaps after %s" % (t, c, k))
