"""Micro-benchmark of the wgrad engines on the hot shapes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srcgan_b200 import ops

DEV = "cuda:0"
SHAPES = [(16, 256, 256, 64, 32), (16, 256, 256, 96, 32), (16, 256, 256, 128, 32), (16, 256, 256, 160, 32),
          (16, 256, 256, 192, 64), (16, 256, 256, 64, 64), (64, 64, 64, 192, 64), (16, 128, 128, 256, 128)]
which = sys.argv[1] if len(sys.argv) > 1 else "tc"
eng = ops.ENGINE_TC if which == "tc" else ops.ENGINE_SIMT
rows = []
for (n, h, w, cin, cout) in SHAPES:
    x = ops.Slice(torch.randn((n, h, w, 192 if cin <= 192 else cin), dtype=torch.bfloat16, device=DEV), 0, cin)
    g = ops.Slice(torch.randn((n, h, w, 192 if cout <= 64 else cout), dtype=torch.bfloat16, device=DEV), 0, cout)
    dw = torch.empty(cout, cin, 3, 3, device=DEV); db = torch.empty(cout, device=DEV)
    for _ in range(2):
        ops.conv_wgrad(x, g, dw, db, 3, 1, 1, engine=eng)
    torch.cuda.synchronize()
    reps = 5
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.conv_wgrad(x, g, dw, db, 3, 1, 1, engine=eng)
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / reps
    fl = 2.0 * n * h * w * cin * cout * 9
    rows.append({"shape": [n, h, w, cin, cout], "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)})
    print(rows[-1], flush=True)
json.dump(rows, open("gpurun_out/bench_wgrad_%s.json" % which, "w"))
