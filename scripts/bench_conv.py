"""Micro-benchmark of the convolution engines on the hot shapes (CUDA events, L2-exceeding tensors)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srcgan_b200 import ops

DEV = "cuda:0"
SHAPES = [  # n, h, w, cin, cout
    (16, 256, 256, 64, 32), (16, 256, 256, 96, 32), (16, 256, 256, 128, 32), (16, 256, 256, 160, 32),
    (16, 256, 256, 192, 64), (16, 256, 256, 64, 64), (64, 64, 64, 192, 64), (16, 256, 256, 64, 128),
    (16, 128, 128, 256, 128),
]
which = sys.argv[1] if len(sys.argv) > 1 else "tc"
eng, layout = (ops.ENGINE_TC, ops.WL_TC) if which == "tc" else (ops.ENGINE_SIMT, ops.WL_RSCK)
rows = []
for (n, h, w, cin, cout) in SHAPES:
    x = ops.Slice(torch.randn((n, h, w, 192 if cin <= 192 else cin), dtype=torch.bfloat16, device=DEV), 0, cin)
    y = ops.Slice(torch.empty((n, h, w, 192 if cout <= 64 else cout), dtype=torch.bfloat16, device=DEV), 0, cout)
    wt = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05
    b = torch.randn(cout, device=DEV)
    wp = ops.pack_weights(wt, layout, torch.bfloat16)
    for _ in range(3):
        ops.conv_fprop(x, wp, b, y, 3, 1, 1, act=0.2, engine=eng)
    torch.cuda.synchronize()
    reps = 10
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.conv_fprop(x, wp, b, y, 3, 1, 1, act=0.2, engine=eng)
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / reps
    fl = 2.0 * n * h * w * cin * cout * 9
    rows.append({"shape": [n, h, w, cin, cout], "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)})
    print(rows[-1], flush=True)
json.dump(rows, open("gpurun_out/bench_conv_%s.json" % which, "w"))
