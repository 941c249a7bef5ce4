"""Time one TC fprop shape at several batch sizes (is the kernel bound by DRAM or by the SM-side feed?):
conv_scale.py cin cout n1 n2 ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from srcgan_b200 import ops
cin, cout = int(sys.argv[1]), int(sys.argv[2])
DEV = "cuda:0"
for n in [int(v) for v in sys.argv[3:]]:
    x = ops.Slice(torch.randn((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV), 0, cin)
    y = ops.Slice(torch.empty((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV), 0, cout)
    wp = ops.pack_weights(torch.randn(cout, cin, 3, 3, device=DEV) * 0.05, ops.WL_TC, torch.bfloat16)
    b = torch.randn(cout, device=DEV)
    for _ in range(3):
        ops.conv_fprop(x, wp, b, y, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ops.conv_fprop(x, wp, b, y, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC)
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 20
    px = n * 65536
    print("n=%d cin=%d cout=%d: %.4f ms  %.0f TFLOP/s  %.2f TB/s algorithmic (%.0f MB touched)  %.2f ns/pixel" % (
        n, cin, cout, ms, 2.0 * px * cin * cout * 9 / ms / 1e9, px * (cin + cout) * 2 / ms / 1e9, px * (cin + cout) * 2 / 1e6, ms * 1e6 / px))
