"""Is the paired-sweep kernel on the narrow dense-block layers limited by the concat-buffer layout (128 B read out of every
384 B pixel, 64 B written) rather than by the kernel?  One 3x3 layer at 64 x 256 x 256 with the input / output slices living in
buffers of different pixel pitch.  python scripts/exp/layout_probe.py [cin cout]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from srcgan_b200 import _lib, ops

DEV = "cuda:0"
N, HW = 64, 256


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
    shapes = [(64, 32), (96, 32), (128, 32), (160, 32), (192, 64), (64, 64)]
    if len(sys.argv) >= 3:
        shapes = [(int(sys.argv[1]), int(sys.argv[2]))]
    for cin, cout in shapes:
        wp = ops.pack_weights(torch.randn(cout, cin, 3, 3, device=DEV) * 0.05, ops.WL_TC, torch.bfloat16)
        bias = torch.randn(cout, device=DEV)
        dw, db = torch.empty(cout, cin, 3, 3, device=DEV), torch.empty(cout, device=DEV)
        flops = 2.0 * N * HW * HW * cin * cout * 9
        pitches = [("concat 192 / 192", 192, 192), ("dense", cin, cout), ("x dense, y 192", cin, 192), ("x 192, y dense", 192, cout),
                   ("x pitch %d, y pitch 64" % (((cin + 63) // 64) * 64), ((cin + 63) // 64) * 64, 64)]
        for name, xl, yl in pitches:
            if xl < cin or yl < cout:
                continue
            X = ops.Slice(torch.randn((N, HW, HW, xl), dtype=torch.bfloat16, device=DEV), 0, cin)
            Y = ops.Slice(torch.zeros((N, HW, HW, yl), dtype=torch.bfloat16, device=DEV), yl - cout, cout)
            f = timed(lambda: ops.conv_fprop(X, wp, bias, Y, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC))
            k = _lib.last_kernel()
            w = timed(lambda: ops.conv_wgrad(X, Y, dw, db, 3, 1, 1, engine=ops.ENGINE_TC))
            px = N * HW * HW
            print(json.dumps({"layer": "%d->%d" % (cin, cout), "layout": name, "kernel": k, "fprop_ms": f, "fprop_tflops": flops / f / 1e9,
                              "fprop_alg_tbs": px * (cin + cout) * 2 / f / 1e9, "wgrad_ms": w, "wgrad_tflops": flops / w / 1e9}), flush=True)
            del X, Y
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
