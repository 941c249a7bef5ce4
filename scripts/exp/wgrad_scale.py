"""Time TC wgrad (kernel + split reduce + bias gradient) of the dense-block shapes at batch n: wgrad_scale.py n"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from srcgan_b200 import ops
DEV = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for cin, cout in [(64, 32), (96, 32), (128, 32), (160, 32), (192, 64), (64, 64)]:
    x = ops.Slice(torch.randn((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV), 0, cin)
    g = ops.Slice(torch.randn((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV), 0, cout)
    dw = torch.empty(cout, cin, 3, 3, device=DEV); db = torch.empty(cout, device=DEV)
    for _ in range(2):
        ops.conv_wgrad(x, g, dw, db, 3, 1, 1, engine=ops.ENGINE_TC)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ops.conv_wgrad(x, g, dw, db, 3, 1, 1, engine=ops.ENGINE_TC)
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 10
    px = n * 65536
    print("n=%d wgrad %d->%d: %.4f ms  %.0f TFLOP/s  %.2f TB/s algorithmic" % (n, cin, cout, ms, 2.0 * px * cin * cout * 9 / ms / 1e9,
                                                                               px * (cin + cout) * 2 / ms / 1e9))
