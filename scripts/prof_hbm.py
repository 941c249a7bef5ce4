#!/usr/bin/env python
"""The bandwidth-bound kernels of the step, one call each at the benchmark's shapes (batch 64, 64x64 -> 256x256; metrics at
512x512), timed with CUDA events and printed as achieved GB/s of ALGORITHMIC bytes against the measured copy bandwidth.

    python scripts/prof_hbm.py                       # event timing, JSON lines
    ncu --set full --clock-control none -k regex:'bn_|loss_|ssim|ae_k|minmax|rgb2lab|lab2rgb|upsample|nchw|nhwc|colsum|reduce|add_' \\
        -o gpurun_out/r2_hbm python scripts/prof_hbm.py --once

Algorithmic bytes = every input element read once + every output element written once."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from srcgan_b200 import _lib, color, losses, metrics, ops

DEV = "cuda:0"
BF = torch.bfloat16


def peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                 "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--once", action="store_true", help="one warm-up + one launch per op (for ncu)")
    ap.add_argument("--n", type=int, default=64)
    args = ap.parse_args()
    n = args.n
    reps = 1 if args.once else 10
    pk = peak()
    ops_list = []

    def op(name, nbytes, fn):
        ops_list.append((name, nbytes, fn))

    # BatchNorm (+LeakyReLU) forward / backward: Decoder bn1 (64 ch @ 256^2) and bn2 (128 ch @ 256^2), D bn (128 ch @ 64^2)
    for c, hw in ((64, 256), (128, 256), (128, 64)):
        x = ops.Slice(torch.randn((n, hw, hw, c), dtype=BF, device=DEV))
        y = ops.Slice(torch.empty((n, hw, hw, c), dtype=BF, device=DEV))
        g = ops.Slice(torch.randn((n, hw, hw, c), dtype=BF, device=DEV))
        gamma, beta = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
        rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
        dg, db = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
        e = x.npix * c * 2
        st = {}

        def fwd(x=x, y=y, gamma=gamma, beta=beta, rm=rm, rv=rv, st=st):
            st["s"] = ops.bn_forward(x, y, gamma, beta, rm, rv, True, 0.1)

        def bwd(x=x, y=y, g=g, gamma=gamma, st=st, dg=dg, db=db):
            ops.bn_backward(g, y, x, g, gamma, st["s"][0], st["s"][1], 0.1, True, dg, db)
        # forward: stats pass reads x, apply pass reads x and writes y (3 e); an ideal fused-with-producer form would be 2 e
        op("bn_forward %dch@%d^2 (2 passes: 3 tensor sweeps)" % (c, hw), 3 * e, fwd)
        # backward: partial pass reads dy, y, x; apply pass reads dy, y, x, writes dx (7 e)
        op("bn_backward %dch@%d^2 (2 passes: 7 tensor sweeps)" % (c, hw), 7 * e, bwd)
    # nearest x2 upsample + adjoint (64 ch, 128^2 -> 256^2)
    lo = ops.Slice(torch.randn((n, 128, 128, 64), dtype=BF, device=DEV))
    hi = ops.Slice(torch.empty((n, 256, 256, 64), dtype=BF, device=DEV))
    op("upsample2x 64ch 128^2->256^2", (lo.npix + hi.npix) * 64 * 2, lambda: ops.upsample2x(lo, hi))
    lo2 = ops.Slice(torch.empty((n, 128, 128, 64), dtype=BF, device=DEV))
    op("upsample2x_adjoint 64ch 256^2->128^2 (+mask)", (2 * lo.npix + hi.npix) * 64 * 2,
       lambda: ops.upsample2x_adjoint(hi, lo2, lo, 0.2))
    # residual add (64 ch @ 256^2... the trunk's skip join runs at 256^2 in G_B)
    a = ops.Slice(torch.randn((n, 256, 256, 64), dtype=BF, device=DEV))
    b = ops.Slice(torch.randn((n, 256, 256, 64), dtype=BF, device=DEV))
    d = ops.Slice(torch.empty((n, 256, 256, 64), dtype=BF, device=DEV))
    op("add 64ch@256^2", 3 * a.npix * 64 * 2, lambda: ops.add(a, b, d))
    # column sums of a 192-channel gradient concat buffer (bias gradients of a dense block)
    cb = ops.Slice(torch.randn((n, 256, 256, 192), dtype=BF, device=DEV))
    cs = torch.empty(192, device=DEV)
    op("colsum 192ch@256^2", cb.npix * 192 * 2, lambda: ops.colsum(cb, cs))
    # layout glue at the module boundary
    img = torch.rand(n, 3, 256, 256, device=DEV)
    thin = ops.Slice(torch.empty((n, 256, 256, 8), dtype=BF, device=DEV), 0, 3)
    op("nchw_to_nhwc 3ch@256^2 (fp32 -> bf16, pitch 8)", img.numel() * 4 + img.numel() * 2, lambda: ops.nchw_to_nhwc(img, thin))
    op("nhwc_to_nchw 3ch@256^2 (bf16 -> fp32)", img.numel() * 2 + img.numel() * 4, lambda: ops.nhwc_to_nchw(thin))
    # fused losses: value + gradient in one pass
    tgt = torch.rand(n, 3, 256, 256, device=DEV)
    op("loss L1 fwd+bwd 3x256^2 fp32", 3 * img.numel() * 4, lambda: ops.loss_fwd_bwd(losses.L1, img, tgt, True))
    op("loss MSE fwd+bwd 3x256^2 fp32", 3 * img.numel() * 4, lambda: ops.loss_fwd_bwd(losses.MSE, img, tgt, True))
    # metrics at the eval size
    p5, t5 = torch.rand(8, 3, 512, 512, device=DEV), torch.rand(8, 3, 512, 512, device=DEV)
    op("ssim 8x3x512^2", 2 * p5.numel() * 4, lambda: ops.ssim_sums(p5, t5, 1.0))
    op("angular error 8x3x512^2", 2 * p5.numel() * 4, lambda: ops.angular_error(p5, t5))
    op("squared error 8x3x512^2", 2 * p5.numel() * 4, lambda: ops.sq_err_sum(p5, t5))
    op("minmax 8x3x512^2", p5.numel() * 4, lambda: ops.minmax(p5))
    if hasattr(ops, "eval_metrics"):
        op("fused eval metrics 8x3x512^2 (MSE+PSNR+AE+SSIM)", 2 * p5.numel() * 4, lambda: ops.eval_metrics(p5, t5))
    # colour
    op("rgb2lab fp32 3x256^2", 2 * img.numel() * 4, lambda: color.rgb2lab(img))
    op("lab2rgb fp32 3x256^2", 2 * img.numel() * 4, lambda: color.lab2rgb(img))
    u8 = (torch.rand(n, 256, 256, 3, device=DEV) * 255).to(torch.uint8)
    op("rgb2lab_u8 exact (fp64 math)", u8.numel() * 1 + u8.numel() * 4, lambda: ops.rgb2lab_u8(u8))
    op("lab2rgb_u8 exact (fp64 math)", u8.numel() * 4 + u8.numel() * 1, lambda: ops.lab2rgb_u8(img))
    # wgrad split reduce + bias-gradient finalise ride on conv_wgrad (160 -> 32 @ 256^2)
    X = ops.Slice(torch.randn((n, 256, 256, 192), dtype=BF, device=DEV), 0, 160)
    DY = ops.Slice(torch.randn((n, 256, 256, 192), dtype=BF, device=DEV), 160, 32)
    dw, dbias = torch.empty(32, 160, 3, 3, device=DEV), torch.empty(32, device=DEV)
    op("conv_wgrad 160->32 (kernel + split reduce; for the reduce's share)", X.npix * 192 * 2,
       lambda: ops.conv_wgrad(X, DY, dw, dbias, 3, 1, 1, engine=ops.ENGINE_TC))

    rows = []
    for name, nbytes, fn in ops_list:
        fn()
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / reps
        row = {"op": name, "ms": ms, "algorithmic_mb": nbytes / 1e6, "gbs": nbytes / ms / 1e6, "frac_of_copy_bw": nbytes / ms / 1e6 / pk,
               "launches": (_lib.launch_count() - l0) // reps, "last_kernel": _lib.last_kernel()}
        rows.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
