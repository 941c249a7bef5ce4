"""Dense block (five 3x3 layers over one 192-channel concat buffer, 64 x 256 x 256) with the work units of every second layer
processed from the last image to the first (SRCGAN_CONV_FLAG_REVERSE): does the next layer then find the previous layer's
tail in L2?   python scripts/exp/rdb_order.py [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from srcgan_b200 import ops

DEV = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
buf = torch.randn((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV)
out = torch.empty((n, 256, 256, 64), dtype=torch.bfloat16, device=DEV)
layers = []
for cin, cout in [(64, 32), (96, 32), (128, 32), (160, 32), (192, 64)]:
    wp = ops.pack_weights(torch.randn(cout, cin, 3, 3, device=DEV) * 0.02, ops.WL_TC, torch.bfloat16)
    layers.append((cin, cout, wp, torch.zeros(cout, device=DEV)))
flop = sum(2.0 * n * 65536 * ci * co * 9 for ci, co, _, _ in layers)


def run(pattern):
    for i, (cin, cout, wp, bias) in enumerate(layers):
        dst = ops.Slice(buf, cin, 32) if cout == 32 else ops.Slice(out, 0, 64)
        ops.conv_fprop(ops.Slice(buf, 0, cin), wp, bias, dst, 3, 1, 1, act=0.2 if cout == 32 else None, engine=ops.ENGINE_TC,
                       reverse=pattern[i])


for name, pattern in (("same order", [False] * 5), ("alternating", [False, True, False, True, False]),
                      ("same order", [False] * 5), ("alternating", [False, True, False, True, False])):
    for _ in range(2):
        run(pattern)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        run(pattern)
    e.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 5
    print("n=%d %-12s: %.3f ms per dense block  %.0f TFLOP/s" % (n, name, ms, flop / ms / 1e9), flush=True)
