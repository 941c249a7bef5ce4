"""Fused dense-block layer pair (csrc/conv_pair.cuh) against the two paired-sweep launches it replaces, at the benchmark's
shape (64 x 256 x 256, bf16, 192-channel concat buffer).  CUDA events on the launching stream; the buffers (1.6 GB each) are
far larger than L2.  Usage: python scripts/exp/pair_bench.py [batch]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from srcgan_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
h = w = 256
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
buf = torch.randn((n, h, w, 192), dtype=torch.bfloat16, device=dev, generator=g)
out = {}
for cin in (64, 128):
    wa = ops.pack_weights(torch.randn((32, cin, 3, 3), device=dev, generator=g) * 0.05, ops.WL_TC, torch.bfloat16)
    wb = ops.pack_weights(torch.randn((32, cin + 32, 3, 3), device=dev, generator=g) * 0.05, ops.WL_TC, torch.bfloat16)
    ba = torch.randn(32, device=dev, generator=g)
    bb = torch.randn(32, device=dev, generator=g)
    bits = [torch.empty((n, h, w, 1), dtype=torch.int32, device=dev) for _ in range(2)]
    xa, ya = ops.Slice(buf, 0, cin), ops.Slice(buf, cin, 32)
    xb, yb = ops.Slice(buf, 0, cin + 32), ops.Slice(buf, cin + 32, 32)

    def fused():
        assert ops.conv_fprop_pair(xa, wa, ba, ya, xb, wb, bb, yb, act=0.2, signbits=bits)

    def separate():
        ops.conv_fprop(xa, wa, ba, ya, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC, signbits=bits[0])
        ops.conv_fprop(xb, wb, bb, yb, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC, signbits=bits[1])

    def fused_bwd():
        assert ops.conv_fprop_pair(xa, wa, None, ya, xb, wb, None, yb, maskbits=bits, mask_slope=0.2)

    def separate_bwd():
        ops.conv_fprop(xa, wa, None, ya, 3, 1, 1, engine=ops.ENGINE_TC, maskbits=bits[0], mask_slope=0.2)
        ops.conv_fprop(xb, wb, None, yb, 3, 1, 1, engine=ops.ENGINE_TC, maskbits=bits[1], mask_slope=0.2)

    flops = 2.0 * n * h * w * 9 * 32 * (2 * cin + 32)
    for name, fn in (("separate", separate), ("fused", fused), ("separate_bwd", separate_bwd), ("fused_bwd", fused_bwd)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        out["%d+%d %s" % (cin, cin + 32, name)] = {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1)}
        print("%3d->32 + %3d->32  %-13s %.4f ms  %.0f TFLOP/s" % (cin, cin + 32, name, ms, flops / ms / 1e9), flush=True)
print(json.dumps(out))
