"""Kernel-level parity (through the C ABI) against torch fp32 references computed on the CPU."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def rand(shape, seed, scale=1.0):
    return (torch.rand(*shape, generator=torch.Generator().manual_seed(seed)) - 0.5) * 2 * scale


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _lib_last_kernel():
    from srcgan_b200 import _lib
    return _lib.last_kernel()


def _lib_launch_count():
    from srcgan_b200 import _lib
    return _lib.launch_count()


def to_nhwc(x, dtype, ctot=None, c0=0):
    from srcgan_b200 import ops
    n, c, h, w = x.shape
    buf = torch.zeros((n, h, w, ctot or c), dtype=dtype, device=DEV)
    buf[..., c0:c0 + c] = x.permute(0, 2, 3, 1).to(DEV, dtype)
    return ops.Slice(buf, c0, c)


def from_nhwc(s):
    return s.view().float().permute(0, 3, 1, 2).cpu()


CONV_CASES = [
    # n, h, w, cin, cout, k, stride, pad, upsample
    (2, 12, 10, 64, 32, 3, 1, 1, False),
    (1, 9, 11, 96, 32, 3, 1, 1, False),
    (1, 8, 8, 192, 64, 3, 1, 1, False),
    (2, 7, 5, 3, 64, 3, 1, 1, False),
    (2, 7, 5, 64, 3, 3, 1, 1, False),
    (1, 6, 6, 64, 64, 3, 1, 1, True),
    (2, 16, 16, 3, 64, 4, 2, 1, False),
    (2, 16, 14, 64, 128, 4, 2, 1, False),
    (1, 9, 9, 128, 256, 4, 1, 1, False),
    (2, 8, 8, 256, 1, 4, 1, 1, False),
    (1, 16, 16, 128, 128, 3, 2, 1, False),
    (1, 15, 13, 128, 256, 3, 2, 1, False),
    (2, 20, 18, 3, 64, 7, 2, 3, False),      # ResDeconv's stem (resdeconv.py): thin 7x7 wgrad in seven filter-row launches
    (1, 40, 70, 1, 64, 7, 2, 3, False),
]


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fprop_dgrad_wgrad(case, dtype, tol):
    from srcgan_b200 import ops
    n, h, w, cin, cout, k, s, p, up = case
    x = rand((n, cin, h, w), 1)
    wt = rand((cout, cin, k, k), 2, 0.1)
    b = rand((cout,), 3)
    if dtype == torch.bfloat16:
        x, wt = x.bfloat16().float(), wt.bfloat16().float()
    xin = F.interpolate(x, scale_factor=2, mode="nearest") if up else x
    xin = xin.clone().requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    y_ref = F.conv2d(xin, wr, br, stride=s, padding=p)
    ho, wo = y_ref.shape[2:]
    gy = rand(tuple(y_ref.shape), 4)
    if dtype == torch.bfloat16:
        gy = gy.bfloat16().float()
    y_ref.backward(gy)

    xs = to_nhwc(x, dtype, ctot=cin + 8, c0=8)           # exercise channel-slice addressing
    ys = ops.Slice(torch.zeros((n, ho, wo, cout + 4), dtype=dtype, device=DEV), 4, cout)
    wp = ops.pack_weights(wt.to(DEV), ops.WL_RSCK, dtype)
    ops.conv_fprop(xs, wp, b.to(DEV), ys, k, s, p, upsample=up)
    assert relerr(from_nhwc(ys), y_ref.detach()) < tol

    gys = to_nhwc(gy, dtype)
    dw = torch.empty((cout, cin, k, k), device=DEV)
    db = torch.empty((cout,), device=DEV)
    ops.conv_wgrad(xs, gys, dw, db, k, s, p, upsample=up)
    assert relerr(dw.cpu(), wr.grad) < tol
    assert relerr(db.cpu(), br.grad) < tol
    ops.conv_wgrad(xs, gys, dw, db, k, s, p, upsample=up, accumulate=True, alpha=0.5)
    assert relerr(dw.cpu(), 1.5 * wr.grad) < tol

    if not up:
        dxs = ops.Slice(torch.zeros((n, h, w, cin), dtype=dtype, device=DEV))
        wd = ops.pack_weights(wt.to(DEV), ops.WL_RSKC, dtype)
        ops.conv_dgrad(gys, wd, dxs, k, s, p)
        assert relerr(from_nhwc(dxs), xin.grad) < tol
        if s == 1:   # dgrad as an fprop over transposed+rotated weights
            wtp = ops.pack_weights(wt.transpose(0, 1).flip(2, 3).contiguous().to(DEV), ops.WL_RSCK, dtype)
            dxs2 = ops.Slice(torch.zeros((n, h, w, cin), dtype=dtype, device=DEV))
            ops.conv_fprop(gys, wtp, None, dxs2, k, 1, k - 1 - p)
            assert relerr(from_nhwc(dxs2), xin.grad) < tol


THIN_IN_CASES = [
    # n, h, w, cin, cout, k, stride, pad
    (2, 40, 70, 3, 64, 7, 2, 3),       # ResDeconv's stem, ragged tile grid
    (3, 64, 64, 3, 64, 4, 2, 1),       # first discriminator layer
    (1, 33, 17, 1, 128, 4, 1, 1),      # two 64-channel groups
    (2, 19, 45, 3, 64, 3, 1, 1),
    (1, 30, 30, 1, 64, 9, 1, 4),       # SRCNN's first layer
]


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("case", THIN_IN_CASES)
def test_thin_in_tiled_kernel(case, dtype, tol, monkeypatch):
    """thin_in_tiled (<= 4 input channels, tile of 8 x 32 pixels x 64 channels per block, patch and weights in shared memory):
    against torch, bit for bit against thin_in_conv (same accumulation order), with the whole epilogue; and the gather-form
    gradient of a thin-output layer (256 -> 1 patch logits) through the same kernel."""
    from srcgan_b200 import ops
    n, h, w, cin, cout, k, s, p = case
    x = rand((n, cin, h, w), 11)
    wt = rand((cout, cin, k, k), 12, 0.1)
    b = rand((cout,), 13)
    if dtype == torch.bfloat16:
        x, wt = x.bfloat16().float(), wt.bfloat16().float()
    y_ref = F.conv2d(x, wt, b, stride=s, padding=p)
    ho, wo = y_ref.shape[2:]
    r1 = rand(tuple(y_ref.shape), 14)
    mk = rand(tuple(y_ref.shape), 15)
    if dtype == torch.bfloat16:
        r1, mk = r1.bfloat16().float(), mk.bfloat16().float()
    lre = lambda t: torch.where(t > 0, t, 0.2 * t)
    full_ref = (lre(y_ref) * 0.5 + 0.25 * r1) * torch.where(mk > 0, torch.ones_like(mk), torch.full_like(mk, 0.1))
    xs = to_nhwc(x, dtype, ctot=8)                         # thin buffers have a pitch of 8 channels
    wp = ops.pack_weights(wt.to(DEV), ops.WL_RSCK, dtype)
    r1s, mks = to_nhwc(r1, dtype), to_nhwc(mk, dtype)

    def run():
        y0 = ops.Slice(torch.zeros((n, ho, wo, cout), dtype=dtype, device=DEV))
        ops.conv_fprop(xs, wp, b.to(DEV), y0, k, s, p)
        y1 = ops.Slice(torch.zeros((n, ho, wo, cout), dtype=dtype, device=DEV))
        ops.conv_fprop(xs, wp, b.to(DEV), y1, k, s, p, act=0.2, alpha=0.5, r1=r1s, beta1=0.25, mask=mks, mask_slope=0.1)
        torch.cuda.synchronize()
        return y0, y1

    y0, y1 = run()
    assert relerr(from_nhwc(y0), y_ref) < tol
    assert relerr(from_nhwc(y1), full_ref) < tol
    monkeypatch.setenv("SRCGAN_B200_NO_THIN_TILED", "1")
    o0, o1 = run()
    assert torch.equal(y0.buf, o0.buf) and torch.equal(y1.buf, o1.buf)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_thin_in_tiled_dgrad(dtype, tol, monkeypatch):
    from srcgan_b200 import ops
    n, h, w, cin, cout, k, p = 2, 21, 37, 256, 1, 4, 1
    x = rand((n, cin, h, w), 21).requires_grad_(True)
    wt = rand((cout, cin, k, k), 22, 0.1)
    if dtype == torch.bfloat16:
        wt = wt.bfloat16().float()
    y = F.conv2d(x, wt, None, stride=1, padding=p)
    gy = rand(tuple(y.shape), 23)
    if dtype == torch.bfloat16:
        gy = gy.bfloat16().float()
    y.backward(gy)
    gys = to_nhwc(gy, dtype, ctot=8)
    wd = ops.pack_weights(wt.to(DEV), ops.WL_RSKC, dtype)

    def run():
        dxs = ops.Slice(torch.zeros((n, h, w, cin), dtype=dtype, device=DEV))
        ops.conv_dgrad(gys, wd, dxs, k, 1, p)
        torch.cuda.synchronize()
        return dxs

    new = run()
    assert relerr(from_nhwc(new), x.grad) < tol
    monkeypatch.setenv("SRCGAN_B200_NO_THIN_TILED", "1")
    assert torch.equal(new.buf, run().buf)


def test_conv_epilogue():
    from srcgan_b200 import ops
    n, h, w, cin, cout = 1, 6, 7, 64, 32
    x, wt, b = rand((n, cin, h, w), 1), rand((cout, cin, 3, 3), 2, 0.1), rand((cout,), 3)
    r1, r2, mk = rand((n, cout, h, w), 5), rand((n, cout, h, w), 6), rand((n, cout, h, w), 7)
    y = F.leaky_relu(F.conv2d(x, wt, b, padding=1), 0.2) * 0.3 + 0.7 * r1 - 1.5 * r2
    y = y * torch.where(mk > 0, torch.ones_like(mk), torch.full_like(mk, 0.2))
    ys = ops.Slice(torch.zeros((n, h, w, cout), device=DEV))
    ops.conv_fprop(to_nhwc(x, torch.float32), ops.pack_weights(wt.to(DEV), ops.WL_RSCK, torch.float32), b.to(DEV), ys,
                   3, 1, 1, act=0.2, alpha=0.3, r1=to_nhwc(r1, torch.float32), beta1=0.7,
                   r2=to_nhwc(r2, torch.float32), beta2=-1.5, mask=to_nhwc(mk, torch.float32), mask_slope=0.2)
    assert relerr(from_nhwc(ys), y) < 2e-5


@pytest.mark.parametrize("c", [3, 64])
def test_layout_roundtrip_and_adjoint(c):
    from srcgan_b200 import ops
    x = rand((2, c, 10, 12), 1)
    s = ops.Slice(torch.zeros((2, 10, 12, c + 5), device=DEV), 5, c)
    ops.nchw_to_nhwc(x.to(DEV), s)
    assert torch.equal(from_nhwc(s), x)
    assert torch.equal(ops.nhwc_to_nchw(s).cpu(), x)
    # adjoint of nearest x2
    g = rand((2, c, 20, 24), 2)
    ref = F.avg_pool2d(g, 2) * 4
    d = ops.Slice(torch.zeros((2, 10, 12, c), device=DEV))
    ops.upsample2x_adjoint(to_nhwc(g, torch.float32), d)
    assert relerr(from_nhwc(d), ref) < 1e-6
    if c % 8 == 0:      # the vectorised bf16 kernel (with and without the LeakyReLU mask), channel slices of wider buffers
        gb = g.to(torch.bfloat16).float()
        mk = rand((2, c, 10, 12), 3) - 0.5
        src = ops.Slice(torch.zeros((2, 20, 24, c + 8), dtype=torch.bfloat16, device=DEV), 8, c)
        src.view().copy_(gb.permute(0, 2, 3, 1).to(DEV))
        msk = ops.Slice(mk.permute(0, 2, 3, 1).contiguous().to(DEV).to(torch.bfloat16))
        for m in (None, msk):
            db = ops.Slice(torch.zeros((2, 10, 12, 2 * c), dtype=torch.bfloat16, device=DEV), c, c)
            ops.upsample2x_adjoint(src, db, m, 0.2)
            want = F.avg_pool2d(gb, 2) * 4
            if m is not None:
                want = want * torch.where(mk.to(torch.bfloat16).float() > 0, torch.ones_like(mk), torch.full_like(mk, 0.2))
            got = db.view().float().permute(0, 3, 1, 2).cpu()
            assert relerr(got, want) < 1e-2 and torch.equal(got, want.to(torch.bfloat16).float())
            assert float(db.buf[..., :c].abs().max()) == 0.0           # the neighbouring slice is untouched


@pytest.mark.parametrize("training", [True, False])
def test_batchnorm_lrelu(training):
    from srcgan_b200 import ops
    n, c, h, w = 3, 128, 9, 7
    x = rand((n, c, h, w), 1, 2.0) + 0.3
    gamma, beta = rand((c,), 2) + 1.5, rand((c,), 3)
    rm, rv = rand((c,), 4) * 0.1, rand((c,), 5).abs() + 0.5
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    y_ref = F.leaky_relu(F.batch_norm(xr, rm_ref, rv_ref, gr, br, training, 0.1, 1e-5), 0.2)
    gy = rand((n, c, h, w), 6)
    y_ref.backward(gy)

    xs = to_nhwc(x, torch.float32)
    ys = ops.Slice(torch.zeros((n, h, w, c), device=DEV))
    rm_d, rv_d = rm.to(DEV), rv.to(DEV)
    g_d, b_d = gamma.to(DEV), beta.to(DEV)
    sm, si = ops.bn_forward(xs, ys, g_d, b_d, rm_d, rv_d, training, 0.2)
    assert relerr(from_nhwc(ys), y_ref.detach()) < 2e-5
    assert relerr(rm_d.cpu(), rm_ref) < 1e-5 and relerr(rv_d.cpu(), rv_ref) < 1e-5
    dys = to_nhwc(gy, torch.float32)
    dg, dbt = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
    ops.bn_backward(dys, ys, xs, dys, g_d, sm, si, 0.2, training, dg, dbt)
    assert relerr(from_nhwc(dys), xr.grad) < 1e-4
    assert relerr(dg.cpu(), gr.grad) < 1e-4 and relerr(dbt.cpu(), br.grad) < 1e-4
    # bf16 kernels: with beta the LeakyReLU mask is recomputed from x instead of read from the saved activation -
    # bit-identical gradients (the saved activation is then not touched: it is poisoned with NaNs here)
    xb = to_nhwc(x, torch.bfloat16)
    yb = ops.Slice(torch.zeros((n, h, w, c), dtype=torch.bfloat16, device=DEV))
    smb, sib = ops.bn_forward(xb, yb, g_d, b_d, rm.to(DEV), rv.to(DEV), training, 0.2)
    res = []
    for use_beta in (False, True):
        d = to_nhwc(gy, torch.bfloat16)
        dgb, dbb = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
        ysave = yb if not use_beta else ops.Slice(torch.full((n, h, w, c), float("nan"), dtype=torch.bfloat16, device=DEV))
        ops.bn_backward(d, ysave, xb, d, g_d, smb, sib, 0.2, training, dgb, dbb, beta=b_d if use_beta else None)
        res.append((d.buf.clone(), dgb, dbb))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    # against fp32 autograd on the SAME bf16-rounded x and dy (so the LeakyReLU mask is the same one; a mask flipped by
    # rounding x would change single elements by 0.8*dy and says nothing about the kernel)
    xq = x.bfloat16().float().requires_grad_(True)
    yq = F.leaky_relu(F.batch_norm(xq, rm.clone(), rv.clone(), gamma, beta, training, 0.1, 1e-5), 0.2)
    yq.backward(gy.bfloat16().float())
    got = res[1][0].float().permute(0, 3, 1, 2).cpu()
    assert float((got - xq.grad).norm() / xq.grad.norm()) < 1e-2
    assert float(((got - xq.grad).abs() > 3e-2 * xq.grad.abs().max()).float().mean()) < 1e-3


@pytest.mark.parametrize("shape", [(2, 3, 64, 64), (1, 1, 14, 14), (3, 3, 7, 5)])
def test_fused_losses(shape):
    from srcgan_b200 import losses
    a, b = rand(shape, 1), rand(shape, 2)
    for kind, ref in ((losses.L1, F.l1_loss), (losses.MSE, F.mse_loss)):
        ar = a.clone().requires_grad_(True)
        lr = ref(ar, b) * 3.0
        lr.backward()
        ad = a.to(DEV).requires_grad_(True)
        ld = losses.fused_loss(kind, ad, b.to(DEV)) * 3.0
        ld.backward()
        assert math.isclose(float(ld), float(lr), rel_tol=1e-5)
        assert relerr(ad.grad.cpu(), ar.grad) < 1e-5
    # scalar target (GANLoss label)
    ad = a.to(DEV).requires_grad_(True)
    ld = losses.fused_loss(losses.MSE, ad, 1.0)
    ld.backward()
    ar = a.clone().requires_grad_(True)
    lr = F.mse_loss(ar, torch.ones_like(ar))
    lr.backward()
    assert math.isclose(float(ld), float(lr), rel_tol=1e-5) and relerr(ad.grad.cpu(), ar.grad) < 1e-5
    assert math.isclose(float(losses.PSNRLoss()(a.to(DEV), b.to(DEV))), float(10 * torch.log10(1 / F.mse_loss(a, b))),
                        rel_tol=1e-5)


def test_metrics_against_oracle_and_golden(golden_modules):
    from oracle import srcgan_oracle as O
    from srcgan_b200 import metrics
    g = golden_modules["scalars"]
    a = torch.rand(2, 3, 40, 36, generator=torch.Generator().manual_seed(105))
    b = (a + 0.05 * torch.randn(a.shape, generator=torch.Generator().manual_seed(106))).clamp(0, 1)
    ad, bd = a.to(DEV), b.to(DEV)
    close = lambda x, y, t=2e-5: math.isclose(float(x), float(y), rel_tol=t, abs_tol=1e-7)
    assert close(metrics.MSE()(ad, bd), g["MSEm"]) and close(metrics.PSNR()(ad, bd), g["PSNR"])
    assert close(metrics.SSIM()(ad, bd), g["SSIM"]) and close(metrics.SSIM()(ad, ad), 1.0)
    assert close(metrics.SSIM()(ad * 255, bd * 255), g["SSIM_255"])
    assert close(metrics.SSIM()(ad * 2 - 1, bd * 2 - 1), g["SSIM_neg"])
    assert torch.allclose(metrics.SSIM()(ad, bd, size_average=False).cpu(), g["SSIM_per_image"], rtol=2e-5)
    assert torch.allclose(metrics.AE()(ad, bd).cpu(), g["AE"], rtol=1e-4)
    assert math.isinf(float(metrics.PSNR()(ad, ad)))
    # larger, ragged (non multiple of the 32-pixel tile) image against the oracle
    a = torch.rand(1, 3, 150, 97, generator=torch.Generator().manual_seed(7))
    b = (a + 0.1 * torch.randn(a.shape, generator=torch.Generator().manual_seed(8))).clamp(0, 1)
    assert close(metrics.SSIM()(a.to(DEV), b.to(DEV)), O.ssim(a, b), 5e-5)
    assert torch.allclose(metrics.AE()(a.to(DEV), b.to(DEV)).cpu(), O.angular_error(a, b), rtol=1e-4)
    assert repr(metrics.SSIM()) == "SSIM" and repr(metrics.AE()) == "AE"


@pytest.mark.parametrize("shape,scale,shift", [((2, 3, 40, 36), 1.0, 0.0), ((1, 3, 150, 97), 255.0, 0.0),
                                               ((3, 1, 11, 11), 2.0, -1.0), ((1, 3, 43, 75), 256.0, -1.0),
                                               ((1, 3, 512, 512), 1.0, 0.0)])
def test_fused_eval_metrics_kernel(shape, scale, shift):
    """MSE, PSNR, AE, SSIM (+ the data range L picked on the device: 1 / 255 / 2 / 256) from ONE launch, against the oracle's
    separate functions; ragged sizes (not multiples of the 32-pixel tile, the 11x11 minimum), 1 and 3 channels."""
    from oracle import srcgan_oracle as O
    from srcgan_b200 import _lib, ops
    a = torch.rand(*shape, generator=torch.Generator().manual_seed(41)) * scale + shift
    b = (a + 0.05 * scale * torch.randn(shape, generator=torch.Generator().manual_seed(42)))
    before = _lib.launch_count()
    res = ops.eval_metrics(a.to(DEV), b.to(DEV))
    assert _lib.launch_count() - before == 1 and _lib.last_kernel() == "eval_metrics"
    res = res.cpu()
    n = shape[0]
    want_L = (255 if float(a.max()) > 128 else 1) - (-1 if float(a.min()) < -0.5 else 0)
    assert float(res[4]) == want_L and float(res[5]) == float(a.min()) and float(res[6]) == float(a.max())
    close = lambda x, y, t=5e-5: math.isclose(float(x), float(y), rel_tol=t, abs_tol=1e-6)
    assert close(res[0], O.mse_loss(a, b)) and close(res[1], O.psnr(a, b))
    assert close(res[3], O.ssim(a, b))
    assert torch.allclose(res[8:8 + n], O.ssim(a, b, size_average=False), rtol=5e-5, atol=1e-6)
    if shape[1] == 3:
        assert torch.allclose(res[8 + n:8 + 2 * n], O.angular_error(a, b), rtol=2e-4)
        assert close(res[2], O.angular_error(a, b).mean(), 2e-4)
    # bit-reproducible (fixed-order reduction behind the atomic ticket)
    assert torch.equal(ops.eval_metrics(a.to(DEV), b.to(DEV)).cpu(), res)
    # per-image MSE
    per_mse = ((a - b) ** 2).flatten(1).mean(1)
    assert torch.allclose(res[8 + 2 * n:8 + 3 * n], per_mse, rtol=5e-5)


def test_fused_eval_metrics_scores_every_tile_of_a_batch_on_its_own():
    """A batch whose tiles have different data ranges: the per-image outputs use each tile's OWN range, like a loop that
    scores one tile per call (testCas.py:65-85); evaluate(batch=k) therefore returns the same rows as batch=1."""
    from oracle import srcgan_oracle as O
    from srcgan_b200 import evaluate, ops
    g = torch.Generator().manual_seed(7)
    a = torch.rand(3, 3, 64, 48, generator=g)
    a[1] *= 255.0
    a[2] = a[2] * 2 - 1
    b = a + 0.03 * a.abs().amax(dim=(1, 2, 3), keepdim=True) * torch.randn(a.shape, generator=g)
    res = ops.eval_metrics(a.to(DEV), b.to(DEV)).cpu()
    n = 3
    assert res[8 + 4 * n:8 + 5 * n].tolist() == [1.0, 255.0, 2.0]
    for i in range(n):
        assert math.isclose(float(res[8 + 3 * n + i]), float(O.ssim(a[i:i + 1], b[i:i + 1])), rel_tol=1e-4, abs_tol=1e-6), i
    rows = evaluate.per_image_metrics(a.to(DEV), b.to(DEV)).cpu()
    for i in range(n):
        assert math.isclose(float(rows[i, 0]), float(O.mse_loss(a[i:i + 1], b[i:i + 1])), rel_tol=5e-5)
        assert math.isclose(float(rows[i, 1]), float(O.psnr(a[i:i + 1], b[i:i + 1])), rel_tol=5e-5)
        assert math.isclose(float(rows[i, 2]), float(O.angular_error(a[i:i + 1], b[i:i + 1])), rel_tol=2e-4)
    ident = lambda t: t
    r1, _ = evaluate.evaluate(ident, [(a[i:i + 1].to(DEV), b[i:i + 1].to(DEV)) for i in range(n)], batch=1)
    r3, _ = evaluate.evaluate(lambda t: t, [(a[i:i + 1].to(DEV), b[i:i + 1].to(DEV)) for i in range(n)], batch=3)
    for x, y in zip(r1, r3):
        assert all(math.isclose(x[k], y[k], rel_tol=1e-6, abs_tol=1e-7) for k in x), (x, y)


@pytest.mark.parametrize("shape,scale", [((2, 3, 40, 36), 1.0), ((1, 1, 75, 43), 255.0)])
def test_dssim_loss_backward(shape, scale):
    """losses.DSSIMLoss is differentiable through SSIM (src/losses.py:170-180): gradient against autograd on the oracle."""
    from oracle import srcgan_oracle as O
    from srcgan_b200 import losses
    a = (torch.rand(*shape, generator=torch.Generator().manual_seed(51)) * scale).requires_grad_(True)
    b = (a.detach() + 0.1 * scale * torch.randn(shape, generator=torch.Generator().manual_seed(52)))
    ref = O.dssim_loss(a.double(), b.double())
    ref.backward()
    ad = a.detach().to(DEV).requires_grad_(True)
    out = losses.DSSIMLoss()(ad, b.to(DEV))
    (out * 3.0).backward()
    assert math.isclose(float(out), float(ref), rel_tol=5e-5, abs_tol=1e-6)
    err = (ad.grad.cpu().double() / 3.0 - a.grad.double()).abs().max() / a.grad.double().abs().max()
    assert float(err) < 1e-3, float(err)
    # per-image form
    from srcgan_b200 import metrics
    ad2 = a.detach().to(DEV).requires_grad_(True)
    v = metrics.SSIM()(ad2, b.to(DEV), size_average=False)
    (v * torch.arange(1, shape[0] + 1, device=DEV, dtype=torch.float32)).sum().backward()
    a2 = a.detach().double().requires_grad_(True)
    (O.ssim(a2, b.double(), size_average=False) * torch.arange(1, shape[0] + 1, dtype=torch.float64)).sum().backward()
    err = (ad2.grad.cpu().double() - a2.grad).abs().max() / a2.grad.abs().max()
    assert float(err) < 1e-3, float(err)


def test_color_against_oracle():
    import numpy as np
    from oracle import srcgan_oracle as O
    from srcgan_b200 import color
    rgb = torch.rand(2, 3, 33, 17, generator=torch.Generator().manual_seed(3))
    lab = color.rgb2lab(rgb.to(DEV), True).cpu()
    ref = np.stack([O.lab_normalise(O.rgb2lab(im.permute(1, 2, 0).numpy())) for im in rgb]).transpose(0, 3, 1, 2)
    assert np.abs(lab.numpy() - ref).max() < 2e-5
    back = color.lab2rgb(lab.to(DEV), True).cpu()
    assert float((back - rgb).abs().max()) < 2e-4
    raw = color.rgb2lab(rgb.to(DEV), False).cpu()
    ref_raw = np.stack([O.rgb2lab(im.permute(1, 2, 0).numpy()) for im in rgb]).transpose(0, 3, 1, 2)
    assert np.abs(raw.numpy() - ref_raw).max() < 2e-3


def test_cpu_tensors_are_rejected_loudly():
    from srcgan_b200 import nn as snn
    net = snn.NLayerDiscriminator(3, 64, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.rand(1, 3, 32, 32))


TC_CASES = [
    # n, h, w, cin, cout, k, pad
    (2, 16, 8, 64, 64, 3, 1),        # exactly one tile per image
    (1, 32, 24, 64, 32, 3, 1),
    (2, 20, 13, 96, 32, 3, 1),       # ragged tiles + 32-channel K tail
    (1, 16, 16, 128, 32, 3, 1),
    (1, 17, 9, 160, 32, 3, 1),
    (2, 24, 16, 192, 64, 3, 1),
    (1, 16, 8, 64, 128, 3, 1),
    (1, 16, 16, 256, 128, 3, 1),
    (1, 18, 10, 128, 256, 4, 1),     # discriminator 4x4 s1 (output 17x9)
    (1, 12, 12, 256, 128, 4, 2),     # its dgrad as an fprop (pad = k-1-p = 2)
    (3, 64, 64, 192, 64, 3, 1),      # many tiles per CTA: exercises stage/phase wrap-around
    # column-sweep kernel (h >= 96, cout 32/64, resident weights): ragged strips, ring wrap, segment splits
    (1, 128, 40, 64, 32, 3, 1),
    (2, 130, 35, 96, 32, 3, 1),      # second strip has 2 valid rows; 32-channel K tail
    (1, 256, 70, 64, 64, 3, 1),      # 8-block ring wraps many times
    (1, 200, 19, 160, 32, 3, 1),
    (2, 96, 16, 16, 64, 3, 1),       # single k-step chunk (K = 16)
    (3, 256, 256, 64, 32, 3, 1),     # > 148 units: several units per CTA, segments of 32 / 64 columns
]


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tc_matches_reference(case):
    """tcgen05 engine vs torch fp32 conv on bf16-rounded operands (fp32 accumulate both sides)."""
    from srcgan_b200 import ops
    n, h, w, cin, cout, k, p = case
    x = rand((n, cin, h, w), 11).bfloat16().float()
    wt = rand((cout, cin, k, k), 12, 0.1).bfloat16().float()
    b = rand((cout,), 13)
    y_ref = F.conv2d(x, wt, b, stride=1, padding=p)
    ho, wo = y_ref.shape[2:]
    xs = to_nhwc(x, torch.bfloat16, ctot=cin + 64, c0=0)
    ys = ops.Slice(torch.zeros((n, ho, wo, cout + 64), dtype=torch.bfloat16, device=DEV), 64, cout)
    wp = ops.pack_weights(wt.to(DEV), ops.WL_TC, torch.bfloat16)
    ops.conv_fprop(xs, wp, b.to(DEV), ys, k, 1, p, engine=ops.ENGINE_TC)
    torch.cuda.synchronize()
    got = from_nhwc(ys)
    assert relerr(got, y_ref) < 1e-2, relerr(got, y_ref)
    # untouched neighbouring channels of the concat buffer
    assert float(ys.buf[..., :64].abs().max()) == 0.0


@pytest.mark.parametrize("env", [{}, {"SRCGAN_B200_SWEEP_DYNAMIC": "1"}, {"SRCGAN_B200_SWEEP_TR": "0"},
                                 {"SRCGAN_B200_SWEEP2_CG": "1"}, {"SRCGAN_B200_NO_SWEEP2": "1"}])
def test_conv_tc_sweep_variants(env, monkeypatch):
    """The paired sweep's launch variants (round-robin / dynamic unit queue, row / column orientation, CTA pair / single CTA)
    and the kw-stacked fallback all compute the same convolution; the default (round-robin) launch is bit-reproducible."""
    from srcgan_b200 import ops
    n, h, w, cin, cout = 3, 256, 160, 96, 32
    x = rand((n, cin, h, w), 31).bfloat16().float()
    wt = rand((cout, cin, 3, 3), 32, 0.1).bfloat16().float()
    b = rand((cout,), 33)
    y_ref = F.conv2d(x, wt, b, stride=1, padding=1)
    xs = to_nhwc(x, torch.bfloat16, ctot=192, c0=0)
    wp = ops.pack_weights(wt.to(DEV), ops.WL_TC, torch.bfloat16)

    def run():
        ys = ops.Slice(torch.zeros((n, h, w, 192), dtype=torch.bfloat16, device=DEV), 96, cout)
        ops.conv_fprop(xs, wp, b.to(DEV), ys, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC)
        torch.cuda.synchronize()
        return ys

    base = run()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got = run()
    assert relerr(from_nhwc(got), F.leaky_relu(y_ref, 0.2)) < 1e-2
    if env == {}:
        assert torch.equal(got.buf, base.buf)


@pytest.mark.parametrize("env,exact", [({"SRCGAN_B200_SWEEP_GROUPS": "2"}, True), ({"SRCGAN_B200_NO_PDL": "1"}, True),
                                       ({"SRCGAN_B200_NO_SWEEP_RANGES": "1"}, False)])
def test_conv_tc_lean_epilogue_and_launch_variants(env, exact, monkeypatch):
    """64 -> 64 without residual / mask operands runs the lean epilogue with three groups, range-mode scheduling and
    programmatic dependent launch by default.  Two groups and ordinary stream-ordered launches give the SAME bits (the order of
    the MMAs does not change); whole-image units instead of column ranges move the lap boundaries of the accumulator ring, so
    they agree to fp32 rounding."""
    from srcgan_b200 import ops
    n, h, w, cin, cout = 5, 256, 256, 64, 64
    g0 = torch.Generator(device=DEV).manual_seed(5)
    xb = torch.randn((n, h, w, 192), dtype=torch.bfloat16, device=DEV, generator=g0)
    wt = torch.randn((cout, cin, 3, 3), device=DEV, generator=g0) * 0.05
    wp = ops.pack_weights(wt, ops.WL_TC, torch.bfloat16)
    b = torch.randn(cout, device=DEV, generator=g0)

    def run():
        y = ops.Slice(torch.zeros((n, h, w, cout), dtype=torch.bfloat16, device=DEV))
        # back to back, so that the second launch starts under the first one's tail when dependent launch is on
        ops.conv_fprop(ops.Slice(xb, 0, cin), wp, b, y, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC)
        y2 = ops.Slice(torch.zeros((n, h, w, cout), dtype=torch.bfloat16, device=DEV))
        ops.conv_fprop(y, wp, b, y2, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC)
        torch.cuda.synchronize()
        return y.buf, y2.buf

    base = run()
    ref = F.leaky_relu(F.conv2d(xb[..., :cin].float().permute(0, 3, 1, 2), wt.bfloat16().float(), b, padding=1), 0.2)
    assert relerr(base[0].float().permute(0, 3, 1, 2).cpu(), ref.cpu()) < 1e-2
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got = run()
    if exact:
        assert torch.equal(got[0], base[0]) and torch.equal(got[1], base[1])
    else:
        assert relerr(got[0].float().cpu(), base[0].float().cpu()) < 4e-3 and relerr(got[1].float().cpu(), base[1].float().cpu()) < 8e-3


@pytest.mark.parametrize("env", [{"SRCGAN_B200_WGRAD_PHANTOM": "1"}, {"SRCGAN_B200_NO_L2_64B": "1"}, {"SRCGAN_B200_NO_PDL": "1"}])
def test_wgrad_stack_variants_bit_identical(env, monkeypatch):
    """The stacked wgrad kernel's phantom tap (zeros vs slab data), the L2 promotion of its 32-channel tensor maps and the
    dependent launch change nothing in the result."""
    from srcgan_b200 import ops
    n, h, w = 3, 128, 96
    g0 = torch.Generator(device=DEV).manual_seed(9)
    X = torch.randn((n, h, w, 192), dtype=torch.bfloat16, device=DEV, generator=g0)
    D = torch.randn((n, h, w, 192), dtype=torch.bfloat16, device=DEV, generator=g0)

    def run():
        out = []
        for cin, c0, cout, d0 in ((64, 0, 64, 0), (32, 64, 32, 96), (160, 0, 32, 64)):
            dw, db = torch.empty(cout, cin, 3, 3, device=DEV), torch.empty(cout, device=DEV)
            ops.conv_wgrad(ops.Slice(X, c0, cin), ops.Slice(D, d0, cout), dw, db, 3, 1, 1, engine=ops.ENGINE_TC)
            out += [dw, db]
        torch.cuda.synchronize()
        return out

    base = run()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got = run()
    assert all(torch.equal(a, b) for a, b in zip(base, got))


@pytest.mark.parametrize("cin,cout", [(96, 32), (64, 64)])
def test_conv_tc_packed_masks(cin, cout):
    """signbits out == (stored activation > 0); a launch with maskbits == the same launch with the bf16 activation as mask,
    bit for bit; kernels that do not implement the bits refuse them."""
    from srcgan_b200 import ops
    n, h, w = 2, 128, 100
    g0 = torch.Generator(device=DEV).manual_seed(3)
    xb = torch.randn((n, h, w, 192), dtype=torch.bfloat16, device=DEV, generator=g0)
    x = ops.Slice(xb, 0, cin)
    wp = ops.pack_weights(torch.randn((cout, cin, 3, 3), device=DEV, generator=g0) * 0.05, ops.WL_TC, torch.bfloat16)
    b = torch.randn(cout, device=DEV, generator=g0)
    act = ops.Slice(torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=DEV))
    bits = torch.zeros((n, h, w, cout // 32), dtype=torch.int32, device=DEV)
    ops.conv_fprop(x, wp, b, act, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC, signbits=bits)
    torch.cuda.synchronize()
    pos = (act.buf > 0).view(n, h, w, cout // 32, 32).long()
    want = (pos << torch.arange(32, device=DEV)).sum(-1)
    want = torch.where(want >= 2 ** 31, want - 2 ** 32, want).to(torch.int32)
    assert torch.equal(bits, want)
    # backward-style launch: same input, multiply by LeakyReLU'(act)
    y_bf = ops.Slice(torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=DEV))
    y_bits = ops.Slice(torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=DEV))
    ops.conv_fprop(x, wp, None, y_bf, 3, 1, 1, mask=act, mask_slope=0.2, engine=ops.ENGINE_TC)
    ops.conv_fprop(x, wp, None, y_bits, 3, 1, 1, maskbits=bits, mask_slope=0.2, engine=ops.ENGINE_TC)
    torch.cuda.synchronize()
    assert torch.equal(y_bf.buf, y_bits.buf)
    small = ops.Slice(torch.randn((1, 32, 32, cin), dtype=torch.bfloat16, device=DEV))      # < 96 px: kw-stacked kernel
    ys = ops.Slice(torch.empty((1, 32, 32, cout), dtype=torch.bfloat16, device=DEV))
    with pytest.raises(RuntimeError, match="paired-sweep"):
        ops.conv_fprop(small, wp, None, ys, 3, 1, 1, engine=ops.ENGINE_TC,
                       signbits=torch.zeros((1, 32, 32, cout // 32), dtype=torch.int32, device=DEV))


PAIR_CASES = [
    # n, h, w, cin of the shared prefix
    (2, 40, 256, 64),        # lanes = the 256 pixels of a row (both CTAs full), cluster ranges cut at image boundaries
    (1, 64, 200, 128),       # ragged second strip (72 valid lanes), two K chunks
    (2, 33, 100, 64),        # one strip only: the second CTA of the pair sweeps nothing but zeros
    (1, 160, 48, 64),        # narrow image: lanes = image rows (160), sweep over the 48 columns
    (3, 256, 256, 128),      # conv3 + conv4 at the benchmark's map size
    (4, 256, 256, 64),       # conv1 + conv2
]


@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv_tc_fused_pair(case):
    """srcgan_conv_fprop_pair (two dense-block layers, one launch; csrc/conv_pair.cuh) against torch fp32 convolutions on the
    same bf16 operands, against the two separate launches it replaces, with packed masks out (forward) and in (mirrored
    backward step); the launch is bit-reproducible."""
    from srcgan_b200 import ops
    n, h, w, cin = case
    g0 = torch.Generator().manual_seed(71)
    x = (torch.randn((n, cin, h, w), generator=g0)).bfloat16().float()
    wa = (torch.randn((32, cin, 3, 3), generator=g0) * 0.05).bfloat16().float()
    wb = (torch.randn((32, cin + 32, 3, 3), generator=g0) * 0.05).bfloat16().float()
    ba, bb = torch.randn(32, generator=g0) * 0.1, torch.randn(32, generator=g0) * 0.1
    wpa = ops.pack_weights(wa.to(DEV), ops.WL_TC, torch.bfloat16)
    wpb = ops.pack_weights(wb.to(DEV), ops.WL_TC, torch.bfloat16)

    def fresh():
        buf = torch.zeros((n, h, w, 192), dtype=torch.bfloat16, device=DEV)
        buf[..., :cin] = x.permute(0, 2, 3, 1).to(DEV)
        return buf

    def slices(buf):
        return (ops.Slice(buf, 0, cin), ops.Slice(buf, cin, 32), ops.Slice(buf, 0, cin + 32), ops.Slice(buf, cin + 32, 32))

    # ---- forward form: bias + LeakyReLU, sign bits out
    buf = fresh()
    xa, ya, xb, yb = slices(buf)
    bits = [torch.zeros((n, h, w, 1), dtype=torch.int32, device=DEV) for _ in range(2)]
    assert ops.conv_fprop_pair(xa, wpa, ba.to(DEV), ya, xb, wpb, bb.to(DEV), yb, act=0.2, signbits=bits), "pair not fused"
    torch.cuda.synchronize()
    assert _lib_last_kernel() == "conv3x3_pair_sweep_tc"
    got_a, got_b = from_nhwc(ya), from_nhwc(yb)
    ref_a = F.leaky_relu(F.conv2d(x, wa, ba, padding=1), 0.2)
    assert relerr(got_a, ref_a) < 1e-2, relerr(got_a, ref_a)
    # layer B against a reference that reads the kernel's own (bf16) x_k: isolates layer B from layer A's rounding
    ref_b = F.leaky_relu(F.conv2d(torch.cat([x, got_a], 1), wb, bb, padding=1), 0.2)
    assert relerr(got_b, ref_b) < 1e-2, relerr(got_b, ref_b)
    if cin + 64 < 192:
        assert float(buf[..., cin + 64:].abs().max()) == 0.0             # nothing written past the two slices
    for sl, bt in ((ya, bits[0]), (yb, bits[1])):
        pos = (sl.view() > 0).view(n, h, w, 1, 32).long()
        want = (pos << torch.arange(32, device=DEV)).sum(-1)
        want = torch.where(want >= 2 ** 31, want - 2 ** 32, want).to(torch.int32)
        assert torch.equal(bt, want)
    # the two launches it replaces (fp32 summation order differs: agreement to bf16 rounding, not bit for bit)
    buf2 = fresh()
    xa2, ya2, xb2, yb2 = slices(buf2)
    ops.conv_fprop(xa2, wpa, ba.to(DEV), ya2, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC)
    ops.conv_fprop(xb2, wpb, bb.to(DEV), yb2, 3, 1, 1, act=0.2, engine=ops.ENGINE_TC)
    torch.cuda.synchronize()
    assert relerr(got_a, from_nhwc(ya2)) < 4e-3 and relerr(got_b, from_nhwc(yb2)) < 8e-3
    # bit-reproducible
    buf3 = fresh()
    xa3, ya3, xb3, yb3 = slices(buf3)
    assert ops.conv_fprop_pair(xa3, wpa, ba.to(DEV), ya3, xb3, wpb, bb.to(DEV), yb3, act=0.2)
    torch.cuda.synchronize()
    assert torch.equal(buf3, buf)
    # ---- mirrored-backward form: no bias, no activation, packed masks in
    mb = [torch.randint(-2 ** 31, 2 ** 31 - 1, (n, h, w, 1), dtype=torch.int64, generator=g0).to(torch.int32).to(DEV)
          for _ in range(2)]
    buf4, buf5 = fresh(), fresh()
    xa4, ya4, xb4, yb4 = slices(buf4)
    assert ops.conv_fprop_pair(xa4, wpa, None, ya4, xb4, wpb, None, yb4, maskbits=mb, mask_slope=0.2)
    torch.cuda.synchronize()
    if h >= 96:                                                           # the unfused kernel reads packed masks on maps >= 96 rows
        xa5, ya5, xb5, yb5 = slices(buf5)
        ops.conv_fprop(xa5, wpa, None, ya5, 3, 1, 1, maskbits=mb[0], mask_slope=0.2, engine=ops.ENGINE_TC)
        ops.conv_fprop(xb5, wpb, None, yb5, 3, 1, 1, maskbits=mb[1], mask_slope=0.2, engine=ops.ENGINE_TC)
        torch.cuda.synchronize()
        assert relerr(from_nhwc(ya4), from_nhwc(ya5)) < 4e-3 and relerr(from_nhwc(yb4), from_nhwc(yb5)) < 8e-3
    bit = lambda t: ((t.view(n, h, w, 1, 1).long() >> torch.arange(32, device=DEV)) & 1).view(n, h, w, 32).bool().cpu()
    ref_m = F.conv2d(x, wa, None, padding=1).permute(0, 2, 3, 1)
    ref_m = torch.where(bit(mb[0]), ref_m, 0.2 * ref_m)
    got_m = ya4.view().float().cpu()
    assert relerr(got_m, ref_m) < 1e-2
    ref_n = F.conv2d(torch.cat([x, got_m.permute(0, 3, 1, 2)], 1), wb, None, padding=1).permute(0, 2, 3, 1)
    ref_n = torch.where(bit(mb[1]), ref_n, 0.2 * ref_n)
    assert relerr(yb4.view().float().cpu(), ref_n) < 1e-2


def test_conv_tc_fused_pair_declines():
    """Shapes the fused kernel does not cover return False without launching (the caller falls back to two launches)."""
    from srcgan_b200 import ops
    n, h, w = 1, 64, 320                                                  # lane extent 320 > 256
    buf = torch.zeros((n, h, w, 192), dtype=torch.bfloat16, device=DEV)
    wpa = ops.pack_weights(torch.zeros((32, 64, 3, 3), device=DEV), ops.WL_TC, torch.bfloat16)
    wpb = ops.pack_weights(torch.zeros((32, 96, 3, 3), device=DEV), ops.WL_TC, torch.bfloat16)
    before = _lib_launch_count()
    assert not ops.conv_fprop_pair(ops.Slice(buf, 0, 64), wpa, None, ops.Slice(buf, 64, 32),
                                   ops.Slice(buf, 0, 96), wpb, None, ops.Slice(buf, 96, 32))
    # layer B must read [prefix | layer A's output]
    buf = torch.zeros((1, 128, 128, 192), dtype=torch.bfloat16, device=DEV)
    assert not ops.conv_fprop_pair(ops.Slice(buf, 0, 64), wpa, None, ops.Slice(buf, 96, 32),
                                   ops.Slice(buf, 0, 96), wpb, None, ops.Slice(buf, 128, 32))
    assert _lib_launch_count() == before


@pytest.mark.parametrize("cin,cout", [(64, 32), (160, 32), (192, 64)])
def test_conv_tc_full_size_identities(cin, cout):
    """BASELINE's full size (batch 64, 256 x 256, the dense block's 192-channel concat buffer) is far beyond what the CPU
    oracle can check element-wise; two size-independent properties tie the three tcgen05 kernels together there:
      * linearity: conv(2x) == 2 conv(x) BIT FOR BIT (a power-of-two scale commutes with every bf16 / fp32 rounding), with
        and without the LeakyReLU-mask epilogue of the dgrad steps;
      * the trilinear form: <conv(x; W), g> == <W, wgrad(x, g)> (both equal sum x*W*g; bf16 output rounding of the left side
        and fp32 accumulation order are the only differences)."""
    from srcgan_b200 import ops
    n, h, w = 64, 256, 256
    g0 = torch.Generator(device=DEV).manual_seed(7)
    xb = torch.randn((n, h, w, 192), dtype=torch.bfloat16, device=DEV, generator=g0)
    x = ops.Slice(xb, 0, cin)
    wt = (torch.randn((cout, cin, 3, 3), device=DEV, generator=g0) * 0.05).bfloat16().float()
    wp = ops.pack_weights(wt, ops.WL_TC, torch.bfloat16)
    y1 = ops.Slice(torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=DEV))
    ops.conv_fprop(x, wp, None, y1, 3, 1, 1, engine=ops.ENGINE_TC)
    x2 = ops.Slice(xb * 2, 0, cin)
    y2 = ops.Slice(torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=DEV))
    ops.conv_fprop(x2, wp, None, y2, 3, 1, 1, engine=ops.ENGINE_TC)
    assert torch.equal(y2.buf, y1.buf * 2)
    mk = ops.Slice(torch.randn((n, h, w, cout), dtype=torch.bfloat16, device=DEV, generator=g0))
    m1 = ops.Slice(torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=DEV))
    m2 = ops.Slice(torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=DEV))
    ops.conv_fprop(x, wp, None, m1, 3, 1, 1, mask=mk, mask_slope=0.25, engine=ops.ENGINE_TC)
    ops.conv_fprop(x2, wp, None, m2, 3, 1, 1, mask=mk, mask_slope=0.25, engine=ops.ENGINE_TC)
    assert torch.equal(m2.buf, m1.buf * 2)
    assert torch.equal(m1.buf, torch.where(mk.buf > 0, y1.buf, y1.buf * 0.25))      # x0.25 is exact in bf16
    del x2, y2, m1, m2, mk
    gy = ops.Slice(torch.randn((n, h, w, cout), dtype=torch.bfloat16, device=DEV, generator=g0))
    dw = torch.empty((cout, cin, 3, 3), device=DEV)
    db = torch.empty((cout,), device=DEV)
    ops.conv_wgrad(x, gy, dw, db, 3, 1, 1, engine=ops.ENGINE_TC)
    torch.cuda.synchronize()
    lhs = float((y1.buf.double() * gy.buf.double()).sum())
    rhs = float((dw.double() * wt.double()).sum())
    scale = float((y1.buf.double() * gy.buf.double()).abs().sum()) ** 0.5 * 16      # ~ std of a sum of that many rounded terms
    assert abs(lhs - rhs) <= max(2e-3 * abs(rhs), scale * 2 ** -8), (lhs, rhs, scale)
    # bias gradient summed inside the wgrad kernel == column sums of dY
    ref_db = gy.buf.double().sum((0, 1, 2))
    assert float((db.double() - ref_db).abs().max()) <= 1e-3 * float(ref_db.abs().max()) + 1.0


@pytest.mark.parametrize("shape", [(2, 32, 16, 128, 64), (1, 160, 24, 64, 64), (1, 128, 50, 128, 32)])
def test_conv_tc_epilogue_matches_simt(shape):
    from srcgan_b200 import ops
    n, h, w, cin, cout = shape
    x, wt, b = rand((n, cin, h, w), 1), rand((cout, cin, 3, 3), 2, 0.1), rand((cout,), 3)
    r1, r2, mk = rand((n, cout, h, w), 5), rand((n, cout, h, w), 6), rand((n, cout, h, w), 7)
    outs = []
    for eng, layout in ((ops.ENGINE_SIMT, ops.WL_RSCK), (ops.ENGINE_TC, ops.WL_TC)):
        ys = ops.Slice(torch.zeros((n, h, w, cout), dtype=torch.bfloat16, device=DEV))
        ops.conv_fprop(to_nhwc(x, torch.bfloat16), ops.pack_weights(wt.to(DEV), layout, torch.bfloat16), b.to(DEV), ys,
                       3, 1, 1, act=0.2, alpha=0.3, r1=to_nhwc(r1, torch.bfloat16), beta1=0.7,
                       r2=to_nhwc(r2, torch.bfloat16), beta2=-1.5, mask=to_nhwc(mk, torch.bfloat16), mask_slope=0.2,
                       engine=eng)
        outs.append(from_nhwc(ys))
    assert relerr(outs[1], outs[0]) < 1e-2


WGRAD_TC_CASES = [
    # n, h, w, cin, cout, k, pad
    (2, 16, 8, 64, 64, 3, 1),
    (2, 20, 13, 96, 32, 3, 1),
    (1, 32, 32, 160, 32, 3, 1),
    (2, 24, 16, 192, 64, 3, 1),
    (1, 16, 16, 64, 128, 3, 1),
    (1, 16, 16, 256, 128, 3, 1),
    (1, 18, 10, 128, 256, 4, 1),
    (4, 64, 64, 64, 64, 3, 1),       # many pixel tiles per CTA
    # kw-stacked kernel: channel blocks of 128 (two atoms) and <= 64 (two vertical taps per M), ragged tiles, cout halves
    (1, 40, 50, 128, 32, 3, 1),
    (2, 24, 40, 160, 64, 3, 1),
    (3, 33, 17, 192, 32, 3, 1),
    (2, 64, 64, 32, 64, 3, 1),
    (1, 9, 70, 64, 32, 3, 1),
]


@pytest.mark.parametrize("case", WGRAD_TC_CASES)
def test_conv_wgrad_tc_matches_reference(case):
    from srcgan_b200 import ops
    n, h, w, cin, cout, k, p = case
    x = rand((n, cin, h, w), 21).bfloat16().float()
    wt = torch.zeros(cout, cin, k, k, requires_grad=True)
    y = F.conv2d(x, wt, None, stride=1, padding=p)
    gy = rand(tuple(y.shape), 22).bfloat16().float()
    y.backward(gy)
    xs = to_nhwc(x, torch.bfloat16, ctot=cin + 64, c0=64)
    gys = to_nhwc(gy, torch.bfloat16, ctot=cout + 32, c0=32) if cout % 64 else to_nhwc(gy, torch.bfloat16, ctot=cout + 64, c0=0)
    dw = torch.empty((cout, cin, k, k), device=DEV)
    db = torch.empty((cout,), device=DEV)
    ops.conv_wgrad(xs, gys, dw, db, k, 1, p, engine=ops.ENGINE_TC)
    torch.cuda.synchronize()
    assert relerr(dw.cpu(), wt.grad) < 5e-3, relerr(dw.cpu(), wt.grad)
    assert relerr(db.cpu(), gy.sum((0, 2, 3))) < 5e-3
    ops.conv_wgrad(xs, gys, dw, db, k, 1, p, engine=ops.ENGINE_TC, accumulate=True, alpha=0.5)
    assert relerr(dw.cpu(), 1.5 * wt.grad) < 5e-3


@pytest.mark.parametrize("ka", [1, 3])
def test_conv_wgrad_split_pairs_dense_block_layers(ka):
    """srcgan_conv_wgrad_split: the weight / bias gradients of conv_k and conv_(k+1) of a dense block from one 64-output-channel
    launch over their common input prefix + one 32 -> 32 launch, against two ordinary launches (and torch on the CPU)."""
    from srcgan_b200 import ops
    n, h, w, nf, gc = 2, 40, 52, 64, 32
    cin_a, d0 = nf + gc * (ka - 1), nf + gc * (3 - ka)
    Cb = (torch.rand((n, h, w, 192), generator=torch.Generator().manual_seed(71)) - 0.5).to(torch.bfloat16).to(DEV)
    Db = (torch.rand((n, h, w, 192), generator=torch.Generator().manual_seed(72)) - 0.5).to(torch.bfloat16).to(DEV)
    ref = {}
    for name, cin, dch in (("a", cin_a, d0 + gc), ("b", cin_a + gc, d0)):
        dw, db = torch.empty((gc, cin, 3, 3), device=DEV), torch.empty((gc,), device=DEV)
        ops.conv_wgrad(ops.Slice(Cb, 0, cin), ops.Slice(Db, dch, gc), dw, db, 3, 1, 1, engine=ops.ENGINE_TC)
        ref[name] = (dw, db)
    for acc in (False, True):
        dwa, dba = torch.full((gc, cin_a, 3, 3), 0.25, device=DEV), torch.full((gc,), 0.25, device=DEV)
        dwb, dbb = torch.full((gc, cin_a + gc, 3, 3), 0.25, device=DEV), torch.full((gc,), 0.25, device=DEV)
        ops.conv_wgrad_split(ops.Slice(Cb, 0, cin_a), ops.Slice(Db, d0, 2 * gc), (dwb, 0, dbb), (dwa, 0, dba), gc, accumulate=acc)
        ops.conv_wgrad_split(ops.Slice(Cb, cin_a, gc), ops.Slice(Db, d0, gc), (dwb, cin_a, None), None, gc, accumulate=acc)
        off = 0.25 if acc else 0.0
        assert relerr((dwa - off).cpu(), ref["a"][0].cpu()) < 1e-4 and relerr((dba - off).cpu(), ref["a"][1].cpu()) < 1e-4
        assert relerr((dwb - off).cpu(), ref["b"][0].cpu()) < 1e-4 and relerr((dbb - off).cpu(), ref["b"][1].cpu()) < 1e-4
    # and against torch
    x = Cb[..., :cin_a + gc].float().permute(0, 3, 1, 2).cpu()
    wt = torch.zeros(gc, cin_a + gc, 3, 3, requires_grad=True)
    gy = Db[..., d0:d0 + gc].float().permute(0, 3, 1, 2).cpu()
    F.conv2d(x, wt, None, padding=1).backward(gy)
    assert relerr(dwb.cpu() - 0.25, wt.grad) < 5e-3


@pytest.mark.parametrize("shape", [(2, 40, 52, 32), (1, 131, 64, 32), (3, 16, 16, 16), (1, 9, 70, 24)])
def test_conv_wgrad_r32_kernel(shape, monkeypatch):
    """conv3x3_wgrad_r32_tc (32-channel slice x 32-channel slice: four vertical taps stacked into M, one MMA per 16-pixel row;
    dY through TMA - the default - or copied by cp.async) against torch and against conv3x3_wgrad_stack_tc<32> on the same
    slices (ragged tile grids, thin slices, accumulate / alpha)."""
    from srcgan_b200 import ops
    n, h, w, cin = shape
    g0 = torch.Generator().manual_seed(123)
    X = (torch.rand((n, h, w, 192), generator=g0) - 0.5).to(torch.bfloat16).to(DEV)
    D = (torch.rand((n, h, w, 192), generator=g0) - 0.5).to(torch.bfloat16).to(DEV)
    xs, ds = ops.Slice(X, 128, cin), ops.Slice(D, 160, 32)

    def run(acc):
        dw = torch.full((32, cin, 3, 3), 0.5, device=DEV)
        ops.conv_wgrad(xs, ds, dw, None, 3, 1, 1, engine=ops.ENGINE_TC, accumulate=acc, alpha=0.5 if acc else 1.0)
        torch.cuda.synchronize()
        return dw

    x = X[..., 128:128 + cin].float().permute(0, 3, 1, 2).cpu()
    wt = torch.zeros(32, cin, 3, 3, requires_grad=True)
    F.conv2d(x, wt, None, padding=1).backward(D[..., 160:192].float().permute(0, 3, 1, 2).cpu())
    monkeypatch.setenv("SRCGAN_B200_WGRAD_R32", "0")
    old = run(False)
    for mode in ("tma", "lsu"):
        monkeypatch.setenv("SRCGAN_B200_WGRAD_R32", mode)
        new = run(False)
        new_acc = run(True)
        assert relerr(new.cpu(), wt.grad) < 5e-3, mode
        assert relerr(new_acc.cpu() - 0.5, 0.5 * wt.grad) < 5e-3, mode
        assert relerr(new.cpu(), old.cpu()) < 1e-5, "r32 kernel (%s) vs the stacked <32> kernel" % mode


@pytest.mark.parametrize("shape", [(2, 72, 40), (1, 33, 64), (2, 16, 17), (1, 9, 130)])
@pytest.mark.parametrize("env", [{}, {"SRCGAN_B200_WGRAD_T22": "0"}, {"SRCGAN_B200_WSTACK32": "1"}, {"SRCGAN_B200_NO_WSTACK": "1"}])
def test_conv_wgrad_tc_variants(env, shape, monkeypatch):
    """kw-stacked wgrad for 64 output channels - with the phantom-free third tap (default: M = two X pixel shifts x N = two dY
    shifts), with (tap | zero row) x N = 192, as two N = 96 halves - and the per-tap halo kernel all give the same weight / bias
    gradients (image widths around the 16-pixel tile: the shifted X window at both borders); the default launch is
    bit-reproducible."""
    from srcgan_b200 import ops
    (n, h, w), cin, cout = shape, 192, 64
    x = rand((n, cin, h, w), 41).bfloat16().float()
    wt = torch.zeros(cout, cin, 3, 3, requires_grad=True)
    y = F.conv2d(x, wt, None, stride=1, padding=1)
    gy = rand(tuple(y.shape), 42).bfloat16().float()
    y.backward(gy)
    xs = to_nhwc(x, torch.bfloat16, ctot=cin, c0=0)
    gys = to_nhwc(gy, torch.bfloat16, ctot=cout + 64, c0=64)

    def run():
        dw = torch.empty((cout, cin, 3, 3), device=DEV)
        db = torch.empty((cout,), device=DEV)
        ops.conv_wgrad(xs, gys, dw, db, 3, 1, 1, engine=ops.ENGINE_TC)
        torch.cuda.synchronize()
        return dw, db

    base_dw, base_db = run()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    dw, db = run()
    assert relerr(dw.cpu(), wt.grad) < 5e-3
    assert relerr(db.cpu(), gy.sum((0, 2, 3))) < 5e-3
    if env == {}:
        assert torch.equal(dw, base_dw) and torch.equal(db, base_db)


S2_CASES = [
    # n, h, w, cin, cout, k
    (2, 32, 16, 128, 128, 3),      # Decoder conv3
    (1, 32, 32, 128, 256, 3),      # Decoder conv4
    (2, 32, 32, 64, 128, 4),       # discriminator conv1
    (1, 34, 18, 64, 64, 3),        # ragged tile grid
    (1, 36, 20, 64, 64, 4),
]


@pytest.mark.parametrize("case", S2_CASES)
def test_conv_tc_stride2_fprop_dgrad_wgrad(case):
    """tcgen05 engine on stride-2 convolutions: element-strided TMA boxes (fprop, wgrad) and the four
    output-phase convolutions (dgrad), against torch fp32 on bf16-rounded operands."""
    from srcgan_b200 import ops
    n, h, w, cin, cout, k = case
    x = rand((n, cin, h, w), 31).bfloat16().float().requires_grad_(True)
    wt = (rand((cout, cin, k, k), 32, 0.1)).bfloat16().float().requires_grad_(True)
    b = rand((cout,), 33)
    y_ref = F.conv2d(x, wt, b, stride=2, padding=1)
    ho, wo = y_ref.shape[2:]
    gy = rand(tuple(y_ref.shape), 34).bfloat16().float()
    y_ref.backward(gy)
    xs = to_nhwc(x.detach(), torch.bfloat16)
    ys = ops.Slice(torch.zeros((n, ho, wo, cout), dtype=torch.bfloat16, device=DEV))
    ops.conv_fprop(xs, ops.pack_weights(wt.detach().to(DEV), ops.WL_TC_S2, torch.bfloat16), b.to(DEV), ys, k, 2, 1,
                   engine=ops.ENGINE_TC)
    assert relerr(from_nhwc(ys), y_ref.detach()) < 1e-2
    gys = to_nhwc(gy, torch.bfloat16)
    mk = rand((n, cin, h, w), 35)
    dxs = ops.Slice(torch.zeros((n, h, w, cin), dtype=torch.bfloat16, device=DEV))
    ops.conv_dgrad(gys, ops.pack_weights(wt.detach().to(DEV), ops.WL_TC_DGRAD_S2, torch.bfloat16), dxs, k, 2, 1,
                   mask=to_nhwc(mk, torch.bfloat16), mask_slope=0.2, engine=ops.ENGINE_TC)
    ref_dx = x.grad * torch.where(mk.bfloat16().float() > 0, torch.ones_like(mk), torch.full_like(mk, 0.2))
    assert relerr(from_nhwc(dxs), ref_dx) < 1e-2
    dw = torch.empty((cout, cin, k, k), device=DEV)
    db = torch.empty((cout,), device=DEV)
    ops.conv_wgrad(xs, gys, dw, db, k, 2, 1, engine=ops.ENGINE_TC)
    assert relerr(dw.cpu(), wt.grad) < 5e-3
    assert relerr(db.cpu(), gy.sum((0, 2, 3))) < 5e-3


@pytest.mark.parametrize("case", [(2, 20, 24, 64, 256, 1), (1, 16, 16, 512, 256, 1), (2, 32, 24, 64, 128, 2), (1, 30, 18, 256, 256, 2),
                                  (2, 16, 8, 256, 64, 1)])
def test_conv_tc_1x1(case):
    """1x1 convolutions (ResDeconv's down-sampling shortcuts and its k2 s2 deconvolutions as 1x1 -> depth-to-space; the
    reference's src/model/resdeconv.py) on the tcgen05 implicit-GEMM kernels: fprop, wgrad, and the stride-1 dgrad."""
    from srcgan_b200 import _lib, ops
    n, h, w, cin, cout, s_ = case
    x = rand((n, cin, h, w), 61).bfloat16().float().requires_grad_(True)
    wt = rand((cout, cin, 1, 1), 62, 0.1).bfloat16().float().requires_grad_(True)
    b = rand((cout,), 63)
    y_ref = F.conv2d(x, wt, b, stride=s_, padding=0)
    ho, wo = y_ref.shape[2:]
    gy = rand(tuple(y_ref.shape), 64).bfloat16().float()
    y_ref.backward(gy)
    xs = to_nhwc(x.detach(), torch.bfloat16)
    ys = ops.Slice(torch.zeros((n, ho, wo, cout), dtype=torch.bfloat16, device=DEV))
    layout = ops.WL_TC if s_ == 1 else ops.WL_TC_S2
    ops.conv_fprop(xs, ops.pack_weights(wt.detach().to(DEV), layout, torch.bfloat16), b.to(DEV), ys, 1, s_, 0, engine=ops.ENGINE_TC)
    assert _lib.last_kernel().startswith("conv_igemm_tc")
    assert relerr(from_nhwc(ys), y_ref.detach()) < 1e-2
    gys = to_nhwc(gy, torch.bfloat16)
    dw = torch.empty((cout, cin, 1, 1), device=DEV)
    db = torch.empty((cout,), device=DEV)
    ops.conv_wgrad(xs, gys, dw, db, 1, s_, 0, engine=ops.ENGINE_TC)
    assert relerr(dw.cpu(), wt.grad) < 5e-3
    assert relerr(db.cpu(), gy.sum((0, 2, 3))) < 5e-3
    if s_ == 1 and cin in (32, 64, 128, 256):
        wtp = ops.pack_weights(wt.detach().transpose(0, 1).contiguous().to(DEV), ops.WL_TC, torch.bfloat16)
        dxs = ops.Slice(torch.zeros((n, h, w, cin), dtype=torch.bfloat16, device=DEV))
        ops.conv_fprop(gys, wtp, None, dxs, 1, 1, 0, engine=ops.ENGINE_TC)
        assert relerr(from_nhwc(dxs), x.grad) < 1e-2


THIN_TC_CASES = [
    # n, h, w, cin, cout   (3x3 stride 1 pad 1; thin side padded to a 16-byte pixel pitch)
    (2, 32, 24, 64, 3),
    (1, 20, 13, 64, 1),
    (2, 32, 24, 3, 64),
    (1, 17, 9, 1, 64),
]


@pytest.mark.parametrize("case", THIN_TC_CASES)
def test_conv_tc_thin_sides(case):
    """Image convs (64->3, 3->64) on the tcgen05 engine: N padded to 16 / K padded to 16 by TMA zero fill."""
    from srcgan_b200 import ops
    n, h, w, cin, cout = case
    x = rand((n, cin, h, w), 41).bfloat16().float()
    wt = rand((cout, cin, 3, 3), 42, 0.1).bfloat16().float().requires_grad_(True)
    b = rand((cout,), 43)
    xr = x.clone().requires_grad_(True)
    y_ref = F.conv2d(xr, wt, b, padding=1)
    gy = rand(tuple(y_ref.shape), 44).bfloat16().float()
    y_ref.backward(gy)
    pad = lambda c: c if c >= 8 else 8
    xs = to_nhwc(x, torch.bfloat16, ctot=pad(cin))
    ys = ops.Slice(torch.zeros((n, h, w, pad(cout)), dtype=torch.bfloat16, device=DEV), 0, cout)
    ops.conv_fprop(xs, ops.pack_weights(wt.detach().to(DEV), ops.WL_TC, torch.bfloat16), b.to(DEV), ys, 3, 1, 1,
                   engine=ops.ENGINE_TC)
    assert relerr(from_nhwc(ys), y_ref.detach()) < 1e-2
    gys = to_nhwc(gy, torch.bfloat16, ctot=pad(cout))
    dw = torch.empty((cout, cin, 3, 3), device=DEV)
    db = torch.empty((cout,), device=DEV)
    ops.conv_wgrad(xs, gys, dw, db, 3, 1, 1, engine=ops.ENGINE_TC)
    assert relerr(dw.cpu(), wt.grad) < 5e-3
    assert relerr(db.cpu(), gy.sum((0, 2, 3))) < 5e-3
    # dgrad as an fprop over transposed weights
    dxs = ops.Slice(torch.zeros((n, h, w, pad(cin)), dtype=torch.bfloat16, device=DEV), 0, cin)
    wtp = ops.pack_weights(wt.detach().transpose(0, 1).flip(2, 3).contiguous().to(DEV), ops.WL_TC, torch.bfloat16)
    ops.conv_fprop(gys, wtp, None, dxs, 3, 1, 1, engine=ops.ENGINE_TC)
    assert relerr(from_nhwc(dxs), xr.grad) < 1e-2


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 3e-2)])
@pytest.mark.parametrize("shape", [(2, 64, 9, 7), (3, 128, 16, 16), (1, 512, 4, 4)])
def test_groupnorm_fwd_bwd(shape, dtype, tol):
    """srcgan_gn_forward / backward (GroupNorm(32, C) with the fused residual + LeakyReLU tail) vs torch on CPU."""
    from srcgan_b200 import ops
    n, c, h, w = shape
    x = rand(shape, 1, 2.0) + 0.3
    res = rand(shape, 2)
    gy = rand(shape, 3)
    gamma = (1.0 + 0.2 * rand((c,), 4)).requires_grad_(True)
    beta = (0.2 * rand((c,), 5)).requires_grad_(True)
    if dtype == torch.bfloat16:
        x, res, gy = x.bfloat16().float(), res.bfloat16().float(), gy.bfloat16().float()
    xr = x.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True)
    y_ref = F.leaky_relu(F.group_norm(xr, 32, gamma, beta, 1e-5) + rr, 0.2)
    y_ref.backward(gy)
    xs, rs, gs = to_nhwc(x, dtype), to_nhwc(res, dtype), to_nhwc(gy, dtype)
    ys = ops.Slice(torch.empty_like(xs.buf))
    g_dev, b_dev = gamma.detach().to(DEV), beta.detach().to(DEV)
    mean, rstd = ops.gn_forward(xs, ys, g_dev, b_dev, 32, 1e-5, residual=rs, act=0.2)
    assert relerr(from_nhwc(ys), y_ref.detach()) < tol
    gz = ops.Slice(torch.empty_like(xs.buf))
    ops.act_backward(gs, ys, gz, 0.2)
    assert relerr(from_nhwc(gz), rr.grad) < tol
    dx = ops.Slice(torch.empty_like(xs.buf))
    dgamma = torch.empty(c, dtype=torch.float32, device=DEV)
    dbeta = torch.empty(c, dtype=torch.float32, device=DEV)
    ops.gn_backward(gz, xs, dx, g_dev, mean, rstd, 32, dgamma, dbeta)
    assert relerr(from_nhwc(dx), xr.grad) < 2 * tol
    assert relerr(dgamma.cpu(), gamma.grad) < tol and relerr(dbeta.cpu(), beta.grad) < tol


@pytest.mark.parametrize("cin,cout", [(64, 32), (96, 32), (128, 32), (160, 32), (192, 64)])
def test_planar_concat_buffer_matches_interleaved(cin, cout):
    """ops.PlanarBuf (a dense block's concat buffer as three 64-channel groups, read across groups by the paired sweep through
    a 5-D tensor map): same numbers as the interleaved 192-channel buffer, bit for bit for the forward convolution (same
    arithmetic order), to fp32 rounding for the weight gradient (one launch per group)."""
    from srcgan_b200 import _lib, ops
    n, h, w = 2, 128, 40
    g = torch.Generator().manual_seed(91)
    inter = (torch.rand((n, h, w, 192), generator=g) - 0.5).to(torch.bfloat16).to(DEV)
    planar = ops.PlanarBuf(n, h, w, 192, torch.bfloat16, DEV)
    for k in range(3):
        planar.groups[k].copy_(inter[..., 64 * k:64 * k + 64])
    wt = ((torch.rand((cout, cin, 3, 3), generator=g) - 0.5) * 0.2).to(DEV)
    b = torch.rand((cout,), generator=g).to(DEV)
    wp = ops.pack_weights(wt, ops.WL_TC, torch.bfloat16)
    oc0 = cin if cout == 32 else 0                       # a dense-block layer writes the slice right after its input prefix
    out_i = torch.zeros((n, h, w, 192), dtype=torch.bfloat16, device=DEV)
    out_p = ops.PlanarBuf(n, h, w, 192, torch.bfloat16, DEV)
    out_p.groups.zero_()
    ops.conv_fprop(ops.Slice(inter, 0, cin), wp, b, ops.Slice(out_i, oc0, cout), 3, 1, 1, act=0.2, engine=ops.ENGINE_TC)
    ops.conv_fprop(ops.Slice(planar, 0, cin), wp, b, ops.Slice(out_p, oc0, cout), 3, 1, 1, act=0.2, engine=ops.ENGINE_TC)
    assert _lib.last_kernel().startswith("conv3x3_sweep2_tc")
    got = torch.cat([out_p.groups[k] for k in range(3)], dim=-1)
    assert torch.equal(got, out_i)
    assert float(out_i[..., oc0:oc0 + cout].abs().max()) > 0
    # weight gradient: dY = the slice just written
    dw_i, db_i = torch.empty((cout, cin, 3, 3), device=DEV), torch.empty((cout,), device=DEV)
    dw_p, db_p = torch.empty((cout, cin, 3, 3), device=DEV), torch.empty((cout,), device=DEV)
    ops.conv_wgrad(ops.Slice(inter, 0, cin), ops.Slice(out_i, oc0, cout), dw_i, db_i, 3, 1, 1, engine=ops.ENGINE_TC)
    ops.conv_wgrad(ops.Slice(planar, 0, cin), ops.Slice(out_p, oc0, cout), dw_p, db_p, 3, 1, 1, engine=ops.ENGINE_TC)
    assert relerr(dw_p.cpu(), dw_i.cpu()) < 1e-4 and relerr(db_p.cpu(), db_i.cpu()) < 1e-4
