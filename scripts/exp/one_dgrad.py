"""One TC fprop launch with a LeakyReLU-mask epilogue (the form every dense-block dgrad step takes): one_dgrad.py cin cout [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from srcgan_b200 import ops
cin, cout = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16
DEV = "cuda:0"
D = torch.randn((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV)
C = torch.randn((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV)
wp = ops.pack_weights(torch.randn(cout, cin, 3, 3, device=DEV) * 0.05, ops.WL_TC, torch.bfloat16)
x, y, m = ops.Slice(D, 0, cin), ops.Slice(D, cin, cout), ops.Slice(C, 64, cout)
for _ in range(3):
    ops.conv_fprop(x, wp, None, y, 3, 1, 1, mask=m, mask_slope=0.2, engine=ops.ENGINE_TC)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.conv_fprop(x, wp, None, y, 3, 1, 1, mask=m, mask_slope=0.2, engine=ops.ENGINE_TC)
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 10
px = n * 65536
print("n=%d dgrad-form %d->%d (+mask): %.4f ms  %.0f TFLOP/s  %.2f TB/s incl. mask" % (n, cin, cout, ms, 2.0 * px * cin * cout * 9 / ms / 1e9,
                                                                              px * (cin + 2 * cout) * 2 / ms / 1e9))
