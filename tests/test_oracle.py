"""Pins the CPU oracle (oracle/srcgan_oracle.py) against outputs of the REAL reference:
the committed golden fixtures (any box) and the live reference modules (build container)."""
import math
import random

import pytest
import torch

from oracle import ref_harness, srcgan_oracle as O

TOL = 2e-5


def rand(shape, seed):
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed))


def probe_like(t, seed):
    return torch.randn(t.shape, generator=torch.Generator().manual_seed(seed))


def relerr(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def check_grad_norms(sd, want, tol=1e-4):
    for k, n in want.items():
        got = float(sd[k].grad.double().norm())
        assert math.isclose(got, n, rel_tol=tol, abs_tol=1e-12), (k, got, n)
    used = {k for k, v in sd.items() if not O.is_buffer_key(k) and v.grad is not None}
    assert used == set(want)


@pytest.mark.parametrize("mode", ["x4", "x2"])
def test_rddbnet_b_golden(golden_modules, mode):
    fx = golden_modules[f"G_A_{mode}"]
    sd = O.as_leaf_params(O.init_rddbnet_b(11))
    x = rand((2, 3, 16, 16), 101).requires_grad_(True)
    y = O.rddbnet_b(sd, x, mode)
    assert y.shape == fx["out"].shape
    assert relerr(y.detach(), fx["out"]) < TOL
    (y * probe_like(y, 7)).sum().backward()
    check_grad_norms(sd, fx["grad_norms"])
    assert relerr(x.grad, fx["dx"]) < 1e-4


def test_rddbnet_a_shim_golden(golden_modules):
    fx = golden_modules["G_B"]
    sd = O.as_leaf_params(O.init_rddbnet_a(12))
    x = rand((2, 3, 32, 32), 102).requires_grad_(True)
    y = O.rddbnet_a(sd, x)
    assert y.shape == (2, 3, 8, 8)
    assert relerr(y.detach(), fx["out"]) < TOL
    (y * probe_like(y, 8)).sum().backward()
    check_grad_norms(sd, fx["grad_norms"])
    assert relerr(x.grad, fx["dx"]) < 1e-4
    for k, v in fx["buffers"].items():
        assert torch.allclose(sd[k].to(v.dtype), v, rtol=1e-5, atol=1e-7), k
    y_eval = O.rddbnet_a(sd, x.detach(), training=False)
    assert relerr(y_eval.detach(), fx["out_eval"]) < TOL


def test_discriminator_golden(golden_modules):
    fx = golden_modules["D"]
    sd = O.as_leaf_params(O.init_discriminator(13))
    x = rand((3, 3, 64, 64), 103).requires_grad_(True)
    y = O.nlayer_discriminator(sd, x)
    assert y.shape == (3, 1, 14, 14)
    assert relerr(y.detach(), fx["out"]) < TOL
    (y * probe_like(y, 9)).sum().backward()
    check_grad_norms(sd, fx["grad_norms"])
    assert relerr(x.grad, fx["dx"]) < 1e-4
    for k, v in fx["buffers"].items():
        assert torch.allclose(sd[k].to(v.dtype), v, rtol=1e-5, atol=1e-7), k


def test_rrdb_golden(golden_modules):
    fx = golden_modules["RRDB"]
    full = O.init_rddbnet_b(14)
    x = rand((1, 64, 12, 10), 104) - 0.5
    assert relerr(O.rrdb(full, "RRDB_trunk.0", x), fx["out"]) < TOL
    assert relerr(O.rdb5(full, "RRDB_trunk.0.RDB1", x), fx["rdb1_out"]) < TOL


def test_losses_metrics_golden(golden_modules):
    s = golden_modules["scalars"]
    a = rand((2, 3, 40, 36), 105)
    b = (a + 0.05 * torch.randn(a.shape, generator=torch.Generator().manual_seed(106))).clamp(0, 1)
    close = lambda x, y: math.isclose(float(x), y, rel_tol=1e-5, abs_tol=1e-7)
    assert close(O.l1_loss(a, b), s["L1"])
    assert close(O.mse_loss(a, b), s["MSE"]) and close(O.mse_loss(a, b), s["MSEm"])
    assert close(O.psnr(a, b), s["PSNRLoss"]) and close(O.psnr(a, b), s["PSNR"])
    assert close(O.ssim(a, b), s["SSIM"])
    assert close(O.dssim_loss(a, b), s["DSSIM"])
    assert close(O.ssim(a * 255.0, b * 255.0), s["SSIM_255"])
    assert close(O.ssim(a * 2 - 1, b * 2 - 1), s["SSIM_neg"])
    assert close(O.ssim(a, a), 1.0) and close(s["SSIM_self"], 1.0)
    assert torch.allclose(O.ssim(a, b, size_average=False), s["SSIM_per_image"], rtol=1e-5)
    assert torch.allclose(O.angular_error(a, b), s["AE"], rtol=1e-5)
    assert math.isinf(float(O.psnr(a, a)))


def test_step_golden(golden_step):
    """Two full G+D steps of the oracle reproduce the reference's losses and updated weights."""
    random.seed(5)
    step = O.CycleGANStepOracle(O.default_states(0))
    for it, rec in enumerate(golden_step["steps"]):
        real_A, real_B = O.synthetic_batch(2, lr=16, scale=4, seed=1234 + it)
        got = step.optimize_parameters(real_A, real_B)
        for n, v in rec["losses"].items():
            assert math.isclose(got[n], v, rel_tol=2e-4, abs_tol=1e-6), (it, n, got[n], v)
        if "fake_B" in rec:
            assert relerr(step.fake_B.detach(), rec["fake_B"]) < 1e-4
            assert relerr(step.fake_A.detach(), rec["fake_A"]) < 1e-4
    for name, norms in golden_step["param_norms"].items():
        sd = getattr(step, name)
        for k, n in norms.items():
            assert math.isclose(float(sd[k].detach().double().norm()), n, rel_tol=1e-5), (name, k)
    for name, bufs in golden_step["buffers"].items():
        sd = getattr(step, name)
        for k, v in bufs.items():
            assert torch.allclose(sd[k].to(v.dtype), v, rtol=1e-4, atol=1e-6), (name, k)


@pytest.mark.parametrize("net", ["2", "SRdens"])
def test_step_golden_gray(golden_step_gray, net):
    """opt.net = '2' (gray LR, RDDB pair) and 'SRdens' (SRDenseNetA/B pair): two oracle steps vs the real train.py."""
    fx = golden_step_gray[net]
    random.seed(5)
    step = O.CycleGANStepOracle(O.default_states(0, net), O.StepOptions(net=net))
    for it, rec in enumerate(fx["steps"]):
        real_A, real_B = O.synthetic_gray_batch(2, lr=16, scale=4, seed=1234 + it)
        got = step.optimize_parameters(real_A, real_B)
        for n, v in rec["losses"].items():
            assert math.isclose(got[n], v, rel_tol=2e-4, abs_tol=1e-6), (net, it, n, got[n], v)
        if "fake_B" in rec:
            assert relerr(step.fake_B.detach(), rec["fake_B"]) < 1e-4
            assert relerr(step.fake_A.detach(), rec["fake_A"]) < 1e-4
    for name, norms in fx["param_norms"].items():
        sd = getattr(step, name)
        for k, n in norms.items():
            assert math.isclose(float(sd[k].detach().double().norm()), n, rel_tol=1e-5), (name, k)


def test_state_dict_keys_match_reference_layout():
    ga = O.init_rddbnet_b(0)
    assert len(ga) == 102 and sum(v.numel() for v in ga.values()) == 2309507
    gb = O.init_rddbnet_a(0)
    assert sum(v.numel() for k, v in gb.items() if not O.is_buffer_key(k)) == 3195715
    d = O.init_discriminator(0)
    assert sum(v.numel() for k, v in d.items() if not O.is_buffer_key(k)) == 663361
    assert d["model.5.weight"].shape == (256, 128, 4, 4)


def test_lab_round_trip():
    import numpy as np
    rng = np.random.default_rng(0)
    rgb = rng.random((16, 16, 3))
    lab = O.rgb2lab(rgb)
    back = O.lab2rgb(lab)
    assert np.abs(back - rgb).max() < 1e-5  # matrix constants are 6-digit: inverse is not exact
    # white / black known answers (D65): L=100,a=b=0 and L=0
    assert np.allclose(O.rgb2lab(np.ones((1, 1, 3))), [[[100.0, 0.0, 0.0]]], atol=2e-2)
    assert np.allclose(O.rgb2lab(np.zeros((1, 1, 3))), [[[0.0, 0.0, 0.0]]], atol=1e-9)
    n = O.lab_normalise(lab)
    assert np.allclose(O.lab_denormalise(n), lab)


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not on this box")
def test_oracle_equals_live_reference():
    """Same weights, same input -> the oracle and the imported reference modules agree."""
    _pkg, M, losses, metrics = ref_harness.import_reference()
    net = M.RDDBNetB(3, 3, 64, nb=3, mode="x4")
    sd = O.init_rddbnet_b(21)
    net.load_state_dict(sd, strict=True)
    x = rand((1, 3, 12, 12), 201)
    assert relerr(O.rddbnet_b(sd, x, "x4"), net(x).detach()) < TOL
    A = ref_harness.make_rddbneta_shim(M)(3, 3, 64, nb=3, mode="x4")
    sda = O.init_rddbnet_a(22)
    A.load_state_dict(sda, strict=True)
    x = rand((2, 3, 16, 16), 202)
    assert relerr(O.rddbnet_a({k: v.clone() for k, v in sda.items()}, x), A(x).detach()) < TOL
    D = M.NLayerDiscriminator(3, 64, 2)
    sdd = O.init_discriminator(23)
    D.load_state_dict(sdd, strict=True)
    x = rand((2, 3, 32, 32), 203)
    assert relerr(O.nlayer_discriminator({k: v.clone() for k, v in sdd.items()}, x), D(x).detach()) < TOL
    a, b = rand((1, 3, 30, 30), 204), rand((1, 3, 30, 30), 205)
    assert math.isclose(float(O.ssim(a, b)), float(metrics.SSIM()(a, b)), rel_tol=1e-5)
    assert math.isclose(float(O.l1_loss(a, b)), float(losses.L1Loss()(a, b)), rel_tol=1e-6)


def cascade_case(which):
    """-> (state_dict, oracle_fn(sd, x), input, probe seed, constructor args) for one cascaded-trainer generator"""
    if which == "SRDN":
        return O.init_srdn(32), (lambda s, t: O.srdn(s, t)), rand((2, 1, 16, 12), 302), 18
    if which.startswith("RDDBNet"):
        up = int(which[-1])
        return O.init_rddbnet_pkg(31, 1, 1, up), (lambda s, t: O.rddbnet_pkg(s, t, up)), rand((2, 1, 16, 12), 301), 17
    if which == "ESPCN_x2":
        return O.init_espcn(33, 1, 1, 2), (lambda s, t: O.espcn(s, t, 2)), rand((2, 1, 14, 10), 303), 19
    if which == "ESPCN_x4_rgb":
        return O.init_espcn(34, 3, 3, 4), (lambda s, t: O.espcn(s, t, 4)), rand((2, 3, 14, 10), 303), 19
    if which == "SRCNN":
        return O.init_srcnn(35, 1, 3), (lambda s, t: O.srcnn(s, t)), rand((2, 1, 14, 10), 303), 19
    raise KeyError(which)


@pytest.mark.parametrize("which", ["ESPCN_x2", "ESPCN_x4_rgb", "SRCNN"])
def test_plain_conv_stacks_golden(golden_cascade, which):
    """ESPCN (espcn.py) / SRCNN (srcnn.py): oracle vs outputs of the real reference."""
    fx = golden_cascade[which]
    sd0, fn, x, seed = cascade_case(which)
    sd = O.as_leaf_params(sd0)
    x = x.requires_grad_(True)
    y = fn(sd, x)
    assert y.shape == fx["out"].shape and relerr(y.detach(), fx["out"]) < TOL
    (y * probe_like(y, seed)).sum().backward()
    check_grad_norms(sd, fx["grad_norms"])
    assert relerr(x.grad, fx["dx"]) < 1e-4


@pytest.mark.parametrize("which", ["RDDBNet_x2", "RDDBNet_x4", "SRDN"])
def test_cascade_generators_golden(golden_cascade, which):
    """package RDDBNet (rddb.py) / SRDN (srdn.py): oracle vs outputs of the real reference."""
    fx = golden_cascade[which]
    if which == "SRDN":
        sd = O.as_leaf_params(O.init_srdn(32))
        x = rand((2, 1, 16, 12), 302).requires_grad_(True)
        y = O.srdn(sd, x)
        seed = 18
    else:
        up = int(which[-1])
        sd = O.as_leaf_params(O.init_rddbnet_pkg(31, 1, 1, up))
        x = rand((2, 1, 16, 12), 301).requires_grad_(True)
        y = O.rddbnet_pkg(sd, x, up)
        seed = 17
    assert y.shape == fx["out"].shape and relerr(y.detach(), fx["out"]) < TOL
    (y * probe_like(y, seed)).sum().backward()
    check_grad_norms(sd, fx["grad_norms"])
    assert relerr(x.grad, fx["dx"]) < 1e-4


ZOO = ["ResDeconv", "EDSR_x2", "EDSR_x4_rgb", "SRDenseNetA_x2", "SRDenseNetA_x4", "SRDenseNetB_x2", "SRDenseNetB_x4"]


def zoo_case(which):
    """-> (state_dict, oracle fn, input, probe seed): same table the fixture generator uses"""
    from oracle.make_golden import zoo_cases
    _ctor, sd, fn, x, seed = zoo_cases()[which]
    return sd, fn, x, seed


@pytest.mark.parametrize("which", ZOO)
def test_zoo_generators_golden(golden_zoo, which):
    """ResDeconv / EDSR / SRDenseNetA,B: oracle vs outputs and gradients of the real reference."""
    fx = golden_zoo[which]
    sd0, fn, x, seed = zoo_case(which)
    sd = O.as_leaf_params(sd0)
    x = x.clone().requires_grad_(True)
    y = fn(sd, x)
    assert y.shape == fx["out"].shape and relerr(y.detach(), fx["out"]) < TOL
    (y * probe_like(y, seed)).sum().backward()
    check_grad_norms(sd, fx["grad_norms"])
    assert relerr(x.grad, fx["dx"]) < 1e-4


def cas_oracle(variant):
    lab, const = "LAB" in variant, "Const" in variant
    sr_state = O.init_srcnn(51, 1, 1) if const else O.init_espcn(51, 1, 1, 2)
    sr_fn = (lambda s, t: O.srcnn(s, t)) if const else (lambda s, t: O.espcn(s, t, 2))
    return O.CascadeStepOracle(sr_state, O.init_srcnn(52, 1, 2 if lab else 3), sr_fn, lambda s, t: O.srcnn(s, t), 2, variant)


@pytest.mark.parametrize("variant", ["", "ConstLAB"])
def test_cascade_step_golden(golden_cas_step, variant):
    """Two CasSRC iterations of the oracle reproduce the real trainCas*.py (losses, PSNRs, transfer output)."""
    step = cas_oracle(variant)
    for it, rec in enumerate(golden_cas_step[variant or "plain"]):
        got = step.optimize_parameters(rand((2, 1, 32, 32), 700 + it), rand((2, 3, 32, 32), 600 + it))
        for k in ("loss_SR", "loss_C", "psnr_SR", "psnr_C"):
            assert math.isclose(got[k], rec[k], rel_tol=1e-4, abs_tol=1e-6), (variant, it, k, got[k], rec[k])
        assert relerr(step.fake_AB, rec["fake_AB"]) < 1e-4
