"""The dominant convolution kernels on the benchmark's layer shapes (64 x 256 x 256 unless noted), two warm-up launches and
one measured launch each - meant to run under `ncu --set full -k regex:'sweep|wgrad_stack|wgrad_r32'`; the launch order printed
here is the order of the rows in the exported CSV (scripts/summarise_ncu_conv.py turns it into profiles/traffic.json)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from srcgan_b200 import _lib, ops

DEV = "cuda:0"
REPS = int(os.environ.get("NCU_SHAPES_REPS", "3"))


def fprop(cin, cout, n, hw_h, hw_w, tall=0):
    X = ops.Slice(torch.randn((n, hw_h, hw_w, 192), dtype=torch.bfloat16, device=DEV), 0, cin)
    Y = ops.Slice(torch.empty((n, hw_h, hw_w, 192), dtype=torch.bfloat16, device=DEV), 192 - cout, cout)
    wp = ops.pack_weights(torch.randn(cout, cin, 3, 3, device=DEV) * 0.05, ops.WL_TC, torch.bfloat16)
    b = torch.randn(cout, device=DEV)
    for _ in range(REPS):
        ops.conv_fprop(X, wp, b, Y, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC, zero_rows=tall)
    torch.cuda.synchronize()
    px = n * hw_h * hw_w
    return {"kernel": _lib.last_kernel(), "what": "fprop %d->%d @ %dx%dx%d%s" % (cin, cout, n, hw_h, hw_w, " (tall image)" if tall else ""),
            "launches": REPS, "algorithmic_bytes": px * (cin + cout) * 2, "flops": 2.0 * px * cin * cout * 9}


def wgrad(cin, cout, n, hw):
    X = ops.Slice(torch.randn((n, hw, hw, 192), dtype=torch.bfloat16, device=DEV), 0, cin)
    DY = ops.Slice(torch.randn((n, hw, hw, 192), dtype=torch.bfloat16, device=DEV), 192 - cout, cout)
    dw, db = torch.empty(cout, cin, 3, 3, device=DEV), torch.empty(cout, device=DEV)
    if cin == 32:        # the 32 -> 32 remainder launches of the paired dense-block wgrads carry no bias gradient (conv3x3_wgrad_r32_tc)
        db = None
    for _ in range(REPS):
        ops.conv_wgrad(X, DY, dw, db, 3, 1, 1, engine=ops.ENGINE_TC)
    torch.cuda.synchronize()
    px = n * hw * hw
    return {"kernel": "conv3x3_wgrad_stack_tc", "what": "wgrad %d->%d @ %dx%dx%d" % (cin, cout, n, hw, hw), "launches": REPS,
            "algorithmic_bytes": px * (cin + cout) * 2, "flops": 2.0 * px * cin * cout * 9}


def pair(cin, n, hw):
    """conv_k + conv_(k+1) of a dense block in one launch (csrc/conv_pair.cuh)"""
    buf = torch.randn((n, hw, hw, 192), dtype=torch.bfloat16, device=DEV)
    wa = ops.pack_weights(torch.randn(32, cin, 3, 3, device=DEV) * 0.05, ops.WL_TC, torch.bfloat16)
    wb = ops.pack_weights(torch.randn(32, cin + 32, 3, 3, device=DEV) * 0.05, ops.WL_TC, torch.bfloat16)
    ba, bb = torch.randn(32, device=DEV), torch.randn(32, device=DEV)
    bits = [torch.empty((n, hw, hw, 1), dtype=torch.int32, device=DEV) for _ in range(2)]
    for _ in range(REPS):
        assert ops.conv_fprop_pair(ops.Slice(buf, 0, cin), wa, ba, ops.Slice(buf, cin, 32), ops.Slice(buf, 0, cin + 32), wb, bb,
                                   ops.Slice(buf, cin + 32, 32), act=0.2, signbits=bits)
    torch.cuda.synchronize()
    px = n * hw * hw
    return {"kernel": "conv3x3_pair_sweep_tc", "what": "fprop pair %d->32 + %d->32 @ %dx%dx%d" % (cin, cin + 32, n, hw, hw),
            "launches": REPS, "algorithmic_bytes": px * (cin + 64) * 2 + px * 8, "flops": 2.0 * px * (2 * cin + 32) * 32 * 9}


rows = [wgrad(192, 64, 64, 256), wgrad(128, 64, 64, 256), wgrad(64, 64, 64, 256), wgrad(32, 32, 64, 256),
        pair(64, 64, 256), pair(128, 64, 256),
        fprop(192, 64, 64, 256, 256), fprop(64, 64, 64, 256, 256), fprop(64, 32, 64, 256, 256), fprop(160, 32, 64, 256, 256),
        fprop(64, 32, 1, 64 * 65 + 1, 64, tall=65)]
for r in rows:
    print(json.dumps(r), flush=True)
