"""Per-layer-shape view of ONE batch-64 G+D step: every convolution call timed with CUDA events on the launching stream and
keyed by (kernel, fprop / dgrad / wgrad, shape), sorted by time, with TFLOP/s against the sustained bf16 peak.  The wgrad rows
include the split-K reduce that follows the tcgen05 kernel.
python scripts/profile_shapes.py [batch] > gpurun_out/profile_shapes.txt"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from srcgan_b200 import nn as snn, ops, trainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
PEAK = 1418.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
except Exception:
    pass
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
ops.timer.enabled = ops.timer.by_shape = True
ops.timer.reset()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
m.optimize_parameters(ra, rb)
b.record()
torch.cuda.synchronize()
step_ms = a.elapsed_time(b)
rows = sorted(ops.timer.summary().items(), key=lambda kv: -kv[1]["ms"])
tot = sum(d["ms"] for _, d in rows)
print("batch %d: instrumented step %.1f ms, convolution calls %.1f ms in %d calls; peak = %.0f TFLOP/s (sustained bf16)" % (
    B, step_ms, tot, sum(d["launches"] for _, d in rows), PEAK))
print("%-36s %-44s %5s %9s %8s %8s %6s %7s" % ("kernel", "call", "calls", "ms/step", "us/call", "TFLOP/s", "frac", "GB/s"))
for key, d in rows:
    kern, shape = key.split(" | ")
    tf = d["flops"] / d["ms"] / 1e9
    print("%-36s %-44s %5d %9.3f %8.1f %8.1f %6.2f %7.0f" % (kern[:36], shape, d["launches"], d["ms"], d["ms"] / d["launches"] * 1e3,
                                                          tf, tf / PEAK, d["bytes"] / d["ms"] / 1e6))
