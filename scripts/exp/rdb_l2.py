"""Does running a dense block's five convolutions over a few images at a time keep the concat buffer in L2?
rdb_l2.py n mb1 mb2 ...   (mb = images per micro-batch; mb = n is the plain layer-by-layer order)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from srcgan_b200 import ops
DEV = "cuda:0"
n = int(sys.argv[1])
buf = torch.randn((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV)
out = torch.empty((n, 256, 256, 192), dtype=torch.bfloat16, device=DEV)
layers = []
for i, (cin, cout) in enumerate([(64, 32), (96, 32), (128, 32), (160, 32), (192, 64)]):
    wp = ops.pack_weights(torch.randn(cout, cin, 3, 3, device=DEV) * 0.02, ops.WL_TC, torch.bfloat16)
    layers.append((cin, cout, wp, torch.zeros(cout, device=DEV)))
flop = sum(2.0 * n * 65536 * ci * co * 9 for ci, co, _, _ in layers)

def run(mb):
    for i0 in range(0, n, mb):
        b, o = buf[i0:i0 + mb], out[i0:i0 + mb]
        for cin, cout, wp, bias in layers:
            dst = ops.Slice(b, cin, 32) if cout == 32 else ops.Slice(o, 0, 64)
            ops.conv_fprop(ops.Slice(b, 0, cin), wp, bias, dst, 3, 1, 1, act=0.2 if cout == 32 else 0.0, engine=ops.ENGINE_TC)

for mb in [int(v) for v in sys.argv[2:]]:
    run(mb); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        run(mb)
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 3
    print("n=%d micro-batch %d: %.3f ms per dense block  %.0f TFLOP/s" % (n, mb, ms, flop / ms / 1e9))
