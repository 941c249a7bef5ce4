#!/usr/bin/env python
"""This repo's tcgen05 kernels vs the incumbent (stock PyTorch / cuDNN, bf16 channels_last, cudnn.benchmark) on the same
B200, layer by layer: fprop, dgrad and wgrad of the dense-block / HRconv shapes at the benchmark's size (64 x 256 x 256).

    python scripts/bench_incumbent_layers.py [--n 64] [--hw 256] > gpurun_out/incumbent_layers.json

Both sides are timed with CUDA events over 10 back-to-back launches after 3 warm-ups.  cuDNN gets a dense NHWC tensor of
exactly cin / cout channels (its best case: no channel-sliced concat buffer, no fused LeakyReLU / residual epilogue - the
reference pays for torch.cat and the activation separately); ours reads / writes channel slices of the 192-channel concat
buffer with the LeakyReLU epilogue, as in the step."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from srcgan_b200 import ops

DEV = "cuda:0"


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--layers", default="64:32,96:32,128:32,160:32,192:64,64:64")
    args = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    n, hw = args.n, args.hw
    rows = []
    for spec in args.layers.split(","):
        cin, cout = (int(v) for v in spec.split(":"))
        flops = 2.0 * n * hw * hw * cin * cout * 9
        row = {"layer": "%d->%d 3x3 @ %dx%dx%d" % (cin, cout, n, hw, hw), "gflop": flops / 1e9}
        # ---- incumbent: cuDNN through torch, bf16 NHWC (channels_last)
        x = torch.randn(n, cin, hw, hw, device=DEV, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = (torch.randn(cout, cin, 3, 3, device=DEV, dtype=torch.bfloat16) * 0.05).contiguous(memory_format=torch.channels_last)
        gy = torch.randn(n, cout, hw, hw, device=DEV, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        bw = lambda mask: torch.ops.aten.convolution_backward(gy, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, mask)
        t = {"fprop": timed(lambda: F.conv2d(x, w, None, 1, 1)),
             "dgrad": timed(lambda: bw([True, False, False])),
             "wgrad": timed(lambda: bw([False, True, False]))}
        row["cudnn_ms"] = t
        row["cudnn_tflops"] = {k: flops / v / 1e9 for k, v in t.items()}
        del x, w, gy
        # ---- this repo: channel slices of a 192-channel concat buffer, LeakyReLU epilogue, fp32 bias
        ctot = 192
        X = ops.Slice(torch.randn((n, hw, hw, ctot), dtype=torch.bfloat16, device=DEV), 0, cin)
        Y = ops.Slice(torch.empty((n, hw, hw, ctot), dtype=torch.bfloat16, device=DEV), ctot - cout, cout)
        wf = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05
        wp = ops.pack_weights(wf, ops.WL_TC, torch.bfloat16)
        wt = ops.pack_weights(wf.transpose(0, 1).flip(2, 3).contiguous(), ops.WL_TC, torch.bfloat16) if cin in (32, 64) else None
        bias = torch.randn(cout, device=DEV)
        dw = torch.empty(cout, cin, 3, 3, device=DEV)
        db = torch.empty(cout, device=DEV)
        o = {"fprop": timed(lambda: ops.conv_fprop(X, wp, bias, Y, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC))}
        row["fprop_kernel"] = __import__("srcgan_b200")._lib.last_kernel()
        # the same kernel on cuDNN's layout (dense cin / cout tensors) - separates the kernel from the concat-buffer layout
        Xd = ops.Slice(torch.randn((n, hw, hw, cin), dtype=torch.bfloat16, device=DEV))
        Yd = ops.Slice(torch.empty((n, hw, hw, cout), dtype=torch.bfloat16, device=DEV))
        o["fprop_dense_layout"] = timed(lambda: ops.conv_fprop(Xd, wp, bias, Yd, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC))
        o["wgrad_dense_layout"] = timed(lambda: ops.conv_wgrad(Xd, Yd, dw, db, 3, 1, 1, engine=ops.ENGINE_TC))
        del Xd, Yd
        # dgrad = fprop over the transposed / rotated weights (cout -> cin channels), with the LeakyReLU mask epilogue
        DY = ops.Slice(Y.buf, ctot - cout, cout)
        DX = ops.Slice(torch.empty((n, hw, hw, ctot), dtype=torch.bfloat16, device=DEV), 0, cin)
        if cin in (32, 64):
            o["dgrad"] = timed(lambda: ops.conv_fprop(DY, wt, None, DX, 3, 1, 1, mask=X, mask_slope=0.2, engine=ops.ENGINE_TC))
            row["dgrad_kernel"] = __import__("srcgan_b200")._lib.last_kernel()
        # (the dgrad of the wider dense-block layers never runs as a single cout -> cin convolution here: the backward of a
        #  dense block is the mirrored dense block, whose steps are again (64 + 32k) -> 32 forward convolutions)
        o["wgrad"] = timed(lambda: ops.conv_wgrad(X, DY, dw, db, 3, 1, 1, engine=ops.ENGINE_TC))
        row["ours_ms"] = o
        row["ours_tflops"] = {k: flops / v / 1e9 for k, v in o.items()}
        row["speedup_vs_cudnn"] = {k: t[k.replace("_dense_layout", "")] / o[k] for k in o}
        rows.append(row)
        print(json.dumps(row), flush=True)
        del X, Y, DX
        torch.cuda.empty_cache()
    print(json.dumps({"summary": {r["layer"]: r["speedup_vs_cudnn"] for r in rows},
                      "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}))


if __name__ == "__main__":
    main()
