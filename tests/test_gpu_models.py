"""Module- and step-level parity of the CUDA path against the CPU oracle and the golden fixtures
(outputs of the real reference).  fp32 mode: <= 1e-3 relative error (north_star); bf16 mode:
SR-image PSNR within 0.05 dB."""
import math
import random

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-3          # north_star tolerance: outputs, losses, BN buffers, D gradients
# Generator gradients: LeakyReLU (and the L1 loss' sign) are discontinuous.  A pre-activation within
# fp32 rounding of zero takes the other branch under a different summation order; that element's
# gradient then differs by 80%.  Measured on B200 (scripts/diag_step.py, diag_tail.py): the
# REFERENCE's own fp32 arithmetic deviates from an fp64 evaluation by 2.6e-3 (median L2 over the
# generator parameters) in a full step, the CUDA fp32 path by 2.8e-3; discriminator gradients and all
# forward outputs agree to <1e-5.  So generator gradients are judged against the fp64 oracle with
# tolerance max(3e-3, 3 x the CPU-fp32 oracle's own error), capped at 1e-2 (the floor is the level
# at which the reference's own fp32 gradients are defined; a single flipped element in a module-level
# test measures 1.1e-3..2.5e-3).
GRAD_FLOOR = 3e-3
GRAD_CAP = 1e-2


def rand(shape, seed):
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed))


def probe_like(t, seed):
    return torch.randn(t.shape, generator=torch.Generator().manual_seed(seed))


def relerr(a, b):
    """max-norm relative error - used for forward outputs and losses"""
    return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def l2err(a, b):
    """L2 relative error - used for gradients.  A LeakyReLU pre-activation that lands within fp32
    rounding of zero can take the other branch than on the CPU (measured: ~1 element in 4M, see
    scripts/diag_tail.py); that single element then differs by 80%, which max-norm would report as a
    1e-2 error of the whole tensor although every other element agrees to 1e-6."""
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.fixture(autouse=True)
def fp32_mode():
    from srcgan_b200 import nn as snn
    old = snn.precision()
    snn.set_precision("fp32")
    yield
    snn.set_precision(old)


class Ref:
    """Oracle results in one precision."""

    def __init__(self, oracle_fn, sd, x, pr, dt):
        from oracle import srcgan_oracle as O
        self.sd = O.as_leaf_params({k: (v.to(dt) if v.is_floating_point() else v.clone()) for k, v in sd.items()})
        self.x = x.detach().clone().to(dt).requires_grad_(True)
        self.y = oracle_fn(self.sd, self.x)
        (self.y * pr.to(dt)).sum().backward()
        self.dx = self.x.grad


def run_both(net, oracle_fn, sd, x, probe_seed):
    """-> (y_gpu, named params (with .grad), dx_gpu, Ref fp32, Ref fp64)"""
    net.load_state_dict(sd, strict=True)
    net.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    y = net(xg)
    pr = probe_like(y, probe_seed)
    (y * pr.to(DEV)).sum().backward()
    return (y, dict(net.named_parameters()), xg.grad,
            Ref(oracle_fn, sd, x, pr, torch.float32), Ref(oracle_fn, sd, x, pr, torch.float64))


def grad_tol(cpu32, truth, floor=GRAD_FLOOR):
    return min(GRAD_CAP, max(floor, 3 * l2err(cpu32, truth)))


def check_grads(named, r32, r64, strict=False, floor=GRAD_FLOOR):
    """every parameter gradient vs the fp64 oracle; ``strict``: plain 1e-3 (no kink allowance)"""
    from oracle import srcgan_oracle as O
    errs = []
    for k, v in r64.sd.items():
        if O.is_buffer_key(k):
            continue
        if v.grad is None:
            assert named[k].grad is None, k
            continue
        assert named[k].grad is not None, k
        e = l2err(named[k].grad, v.grad)
        tol = TOL if strict else grad_tol(r32.sd[k].grad, v.grad, floor)
        assert e < tol, (k, e, tol)
        errs.append(e)
    errs.sort()
    assert errs[len(errs) // 2] < (TOL if strict else 3 * TOL)
    return errs


@pytest.mark.parametrize("mode", ["x4", "x2"])
def test_rddbnet_b_parity(golden_modules, mode):
    from oracle import srcgan_oracle as O
    from srcgan_b200 import nn as snn
    sd = O.init_rddbnet_b(11)
    x = rand((2, 3, 16, 16), 101)
    y, named, dx, r32, r64 = run_both(snn.RDDBNetB(3, 3, 64, nb=3, mode=mode),
                                      lambda s, t: O.rddbnet_b(s, t, mode), sd, x, 7)
    assert relerr(y, r64.y) < 1e-4 and relerr(y, r32.y) < 1e-4
    assert relerr(y, golden_modules[f"G_A_{mode}"]["out"]) < 1e-4          # the real reference's output
    check_grads(named, r32, r64)
    assert l2err(dx, r64.dx) < grad_tol(r32.dx, r64.dx)
    for k, nrm in golden_modules[f"G_A_{mode}"]["grad_norms"].items():
        assert math.isclose(float(named[k].grad.double().norm()), nrm, rel_tol=GRAD_CAP), k


def test_rddbnet_a_shim_parity(golden_modules):
    from oracle import srcgan_oracle as O
    from srcgan_b200 import nn as snn
    sd = O.init_rddbnet_a(12)
    x = rand((2, 3, 32, 32), 102)
    net = snn.RDDBNetA(3, 3, 64, nb=3, mode="x4")
    y, named, dx, r32, r64 = run_both(net, lambda s, t: O.rddbnet_a(s, t), sd, x, 8)
    fx = golden_modules["G_B"]
    assert relerr(y, r64.y) < 1e-4 and relerr(y, fx["out"]) < 1e-4
    check_grads(named, r32, r64)
    assert l2err(dx, r64.dx) < grad_tol(r32.dx, r64.dx)
    state = net.state_dict()
    for k, v in fx["buffers"].items():
        assert torch.allclose(state[k].cpu().to(v.dtype), v, rtol=1e-4, atol=1e-6), k
    net.eval()
    with torch.no_grad():
        assert relerr(net(x.to(DEV)), fx["out_eval"]) < TOL


def test_discriminator_parity(golden_modules):
    from oracle import srcgan_oracle as O
    from srcgan_b200 import nn as snn
    sd = O.init_discriminator(13)
    x = rand((3, 3, 64, 64), 103)
    net = snn.NLayerDiscriminator(3, 64, 2)
    y, named, dx, r32, r64 = run_both(net, lambda s, t: O.nlayer_discriminator(s, t), sd, x, 9)
    fx = golden_modules["D"]
    assert tuple(y.shape) == (3, 1, 14, 14)
    assert relerr(y, r64.y) < 1e-4 and relerr(y, fx["out"]) < 1e-4
    check_grads(named, r32, r64, strict=True)
    assert l2err(dx, r64.dx) < TOL
    state = net.state_dict()
    for k, v in fx["buffers"].items():
        assert torch.allclose(state[k].cpu().to(v.dtype), v, rtol=1e-4, atol=1e-6), k
    # requires_grad=False on the parameters (backward_G phase): only the input gradient flows
    for p in net.parameters():
        p.requires_grad = False
        p.grad = None
    xg = x.to(DEV).requires_grad_(True)
    net(xg).sum().backward()
    assert xg.grad is not None and all(p.grad is None for p in net.parameters())
    # odd sizes (ragged tiles) and the 1-channel variant
    net1 = snn.NLayerDiscriminator(1, 64, 2)
    sd1 = O.init_discriminator(5, input_nc=1)
    x1 = rand((2, 1, 37, 29), 104)
    y1, named1, dx1, q32, q64 = run_both(net1, lambda s, t: O.nlayer_discriminator(s, t), sd1, x1, 3)
    assert relerr(y1, q64.y) < 1e-4 and l2err(dx1, q64.dx) < TOL
    check_grads(named1, q32, q64, strict=True)


def test_state_dict_keys_and_loading():
    from oracle import srcgan_oracle as O
    from srcgan_b200 import nn as snn
    assert list(snn.RDDBNetB(3, 3, 64, nb=3, mode="x4").state_dict().keys()) == list(O.init_rddbnet_b(0).keys())
    assert set(snn.RDDBNetA(3, 3, 64, nb=3, mode="x4").state_dict().keys()) == set(O.init_rddbnet_a(0).keys())
    assert list(snn.NLayerDiscriminator(3, 64, 2).state_dict().keys()) == list(O.init_discriminator(0).keys())


def make_trainer(states, mode="x4", net="1"):
    from srcgan_b200 import trainer
    opt = trainer.params()
    opt.device = torch.device(DEV)
    opt.mode = mode
    opt.net = net
    m = trainer.SRCycleGAN(opt)
    for name in ("G_A", "G_B", "D_A", "D_B"):
        getattr(m, "net" + name).load_state_dict(states[name], strict=True)
    return m


def test_full_step_against_golden_and_oracle(golden_step):
    """Two optimize_parameters calls: the nine losses, the generated images, the updated weights and
    the BN buffers agree with the real reference's (golden) within 1e-3."""
    from oracle import srcgan_oracle as O
    random.seed(5)
    m = make_trainer(O.default_states(0))
    for it, rec in enumerate(golden_step["steps"]):
        real_A, real_B = O.synthetic_batch(2, lr=16, scale=4, seed=1234 + it)
        m.optimize_parameters(real_A.to(DEV), real_B.to(DEV))
        got = m.current_losses()
        for n, v in rec["losses"].items():
            assert math.isclose(got[n], v, rel_tol=TOL, abs_tol=1e-5), (it, n, got[n], v)
        if "fake_B" in rec:
            assert relerr(m.fake_B.detach(), rec["fake_B"]) < TOL
            assert relerr(m.fake_A.detach(), rec["fake_A"]) < TOL
    # Adam's first updates are ~ lr*sign(g): an element whose gradient is ~0 may move the other way, so
    # norms are compared with an absolute allowance of a few flipped elements (2 steps x lr each)
    for name, norms in golden_step["param_norms"].items():
        net = getattr(m, "net" + name)
        lr = 1e-4 if name.startswith("G") else 1e-5
        for k, p in net.named_parameters():
            assert math.isclose(float(p.detach().double().norm()), norms[k], rel_tol=TOL, abs_tol=8 * lr), (name, k)
    # element-wise: the CUDA path may disagree with an fp64 run of the oracle on no more weights than
    # twice what the reference's own fp32 arithmetic (the fp32 oracle) disagrees on
    def run_oracle(dt):
        random.seed(5)
        st = {n: {k: (v.to(dt) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
              for n, sd in O.default_states(0).items()}
        ref = O.CycleGANStepOracle(st)
        for it in range(len(golden_step["steps"])):
            real_A, real_B = O.synthetic_batch(2, lr=16, scale=4, seed=1234 + it)
            ref.optimize_parameters(real_A.to(dt), real_B.to(dt))
        return ref

    r32, r64 = run_oracle(torch.float32), run_oracle(torch.float64)
    for name in ("G_A", "G_B", "D_A", "D_B"):
        lr = 1e-4 if name.startswith("G") else 1e-5
        bad_gpu = bad_cpu = total = 0
        for k, p in getattr(m, "net" + name).named_parameters():
            truth = getattr(r64, name)[k].detach()
            bad_gpu += int(((p.detach().cpu().double() - truth).abs() > 0.25 * lr).sum())
            bad_cpu += int(((getattr(r32, name)[k].detach().double() - truth).abs() > 0.25 * lr).sum())
            total += truth.numel()
        assert bad_gpu <= 2 * bad_cpu + 1e-3 * total, (name, bad_gpu, bad_cpu, total)


@pytest.mark.parametrize("net", ["2", "SRdens"])
def test_full_step_gray_against_golden(golden_step_gray, net):
    """opt.net = '2' (gray LR through the RDDB pair) and 'SRdens' (SRDenseNetA/B pair, train.py:166-170): two
    optimize_parameters calls vs the real reference's."""
    from oracle import srcgan_oracle as O
    fx = golden_step_gray[net]

    # The first step must match to 1e-3 (measured 1e-6).  After the first Adam update (~ lr*sign(g) per weight)
    # weights whose gradient is ~0 move the other way under a different summation order - with a gray input
    # replicated to three identical channels there are many such ties - so the second step's losses carry that
    # noise (measured on B200, scripts/diag_gray.py: 2e-5..1.1e-3 for net '2', <1e-5 for 'SRdens'; the CPU
    # fp32-vs-fp64 oracle pair shows up to 4.7e-4 on the same losses).  They are judged with GRAD_FLOOR = 3e-3,
    # the level at which the reference's own fp32 step is defined (see the header).
    random.seed(5)
    m = make_trainer(O.default_states(0, net), net=net)
    for it, rec in enumerate(fx["steps"]):
        real_A, real_B = O.synthetic_gray_batch(2, lr=16, scale=4, seed=1234 + it)
        m.optimize_parameters(real_A.to(DEV), real_B.to(DEV))
        got = m.current_losses()
        for n, v in rec["losses"].items():
            assert math.isclose(got[n], v, rel_tol=TOL if it == 0 else GRAD_FLOOR, abs_tol=1e-5), (net, it, n, got[n], v)
        if "fake_B" in rec:
            assert relerr(m.fake_B.detach(), rec["fake_B"]) < TOL
            assert relerr(m.fake_A.detach(), rec["fake_A"]) < TOL
    for name, norms in fx["param_norms"].items():
        lr = 1e-4 if name.startswith("G") else 1e-5
        for k, p in getattr(m, "net" + name).named_parameters():
            assert math.isclose(float(p.detach().double().norm()), norms[k], rel_tol=TOL, abs_tol=8 * lr), (name, k)


def test_bf16_step_at_bench_patch_size_is_deterministic_and_finite():
    """The benchmark's patch size (64x64 -> 256x256, where every dense-block layer of G_B runs on the paired sweep / stacked
    wgrad kernels with their split reductions) cannot be checked element-wise on the CPU; what must
    hold at any size: two runs from the same state give bit-identical losses and updated weights, and everything is finite."""
    from oracle import srcgan_oracle as O
    from srcgan_b200 import nn as snn
    snn.set_precision("bf16")
    results = []
    for _ in range(2):
        random.seed(11)
        torch.manual_seed(11)
        m = make_trainer(O.default_states(0))
        real_A, real_B = O.synthetic_batch(4, lr=64, scale=4, seed=77)
        m.optimize_parameters(real_A.to(DEV), real_B.to(DEV))
        m.optimize_parameters(real_A.to(DEV), real_B.to(DEV))
        torch.cuda.synchronize()
        results.append((m.current_losses(), [p.detach().clone() for p in m.netG_B.parameters()]))
    (l0, w0), (l1, w1) = results
    assert all(math.isfinite(v) for v in l0.values()), l0
    assert l0 == l1
    assert all(torch.equal(a, b) for a, b in zip(w0, w1))


def test_bf16_mode_psnr_matches_fp32():
    """north_star: bf16 mode must match the SR image's PSNR within 0.05 dB."""
    from oracle import srcgan_oracle as O
    from srcgan_b200 import nn as snn
    sd = O.init_rddbnet_b(31)
    x = rand((2, 3, 24, 24), 301)
    target = torch.nn.functional.interpolate(x, scale_factor=4, mode="bicubic").clamp(0, 1)
    y_ref = O.rddbnet_b(sd, x, "x4")
    net = snn.RDDBNetB(3, 3, 64, nb=3, mode="x4")
    net.load_state_dict(sd)
    net.to(DEV)
    snn.set_precision("bf16")
    with torch.no_grad():
        y = net(x.to(DEV)).cpu()
    psnr = lambda a, b: float(10 * torch.log10(1.0 / ((a - b) ** 2).mean()))
    assert abs(psnr(y, target) - psnr(y_ref, target)) < 0.05
    assert relerr(y, y_ref) < 0.1


def test_eval_sweep_512_tiles():
    """BASELINE config 5 shape: 512x512 tiles through the generator + MSE/PSNR/AE/SSIM on the device,
    against the CPU oracle on the same weights (bf16 inference: PSNR within 0.05 dB of fp32)."""
    from oracle import srcgan_oracle as O
    from srcgan_b200 import evaluate, nn as snn
    sd = O.init_rddbnet_b(41)
    net = snn.RDDBNetB(3, 3, 64, nb=3, mode="x4")
    net.load_state_dict(sd)
    net.to(DEV).eval()
    hr = [rand((1, 3, 512, 512), 500 + i) for i in range(2)]
    lr = [torch.nn.functional.interpolate(h, scale_factor=0.25, mode="nearest") for h in hr]
    snn.set_precision("bf16")
    rows, mean = evaluate.evaluate(net, [(l.to(DEV), h.to(DEV)) for l, h in zip(lr, hr)])
    assert len(rows) == 2 and set(mean) == {"MSE", "PSNR", "AE", "SSIM"}
    for l, h, row in zip(lr, hr, rows):
        out = O.rddbnet_b(sd, l, "x4")
        assert abs(row["PSNR"] - float(O.psnr(out, h))) < 0.05
        assert math.isclose(row["MSE"], float(O.mse_loss(out, h)), rel_tol=2e-2)
        assert math.isclose(row["SSIM"], float(O.ssim(out, h)), rel_tol=5e-2, abs_tol=5e-3)
        assert math.isclose(row["AE"], float(O.angular_error(out, h).mean()), rel_tol=2e-2)


@pytest.mark.parametrize("const", [False, True])
def test_cascade_eval_from_checkpoint_files(tmp_path, const):
    """testCas.py / testCasConst.py flow: networks rebuilt from checkpoint file NAMES, weights from the files,
    luma -> 1/up -> SR -> colouriser, four metrics per tile; against the CPU oracle on the same weights."""
    import torch.nn.functional as F
    from oracle import srcgan_oracle as O
    from srcgan_b200 import checkpoint as C, evaluate, nn as snn
    sr_name = "SRCNN" if const else "ESPCN"
    sd_a = O.init_srcnn(61, 1, 1) if const else O.init_espcn(61, 1, 1, 2)
    sd_b = O.init_resdeconv(62, 3)
    pa = str(tmp_path / C.cas_checkpoint_name(sr_name, "A2C", 2, 3))
    pb = str(tmp_path / C.cas_checkpoint_name("ResDeconv", "C2B", 2, 3))
    torch.save(sd_a, pa)                      # as the reference's trainCas.py:224-225 would have written them
    torch.save(sd_b, pb)
    net_a, net_b, na, nb = C.load_cascade_pair(pa, pb, DEV)
    assert (na.model, na.up, nb.model, nb.lab) == (sr_name, 2, "ResDeconv", False) and not net_a.training
    tiles = [(rand((1, 1, 64, 64), 800 + i), rand((1, 3, 64, 64), 810 + i)) for i in range(2)]
    rows, mean = evaluate.evaluate_cascade(net_a, net_b, [(a.to(DEV), b.to(DEV)) for a, b in tiles], 2, const=const)
    for (_a, b), row in zip(tiles, rows):
        luma = 0.2125 * b[:, :1] + 0.7154 * b[:, 1:2] + 0.0721 * b[:, 2:3]
        if const:
            lr = F.interpolate(F.interpolate(luma, scale_factor=0.5, mode="bilinear"), scale_factor=2, mode="bilinear")
            sr = O.srcnn(sd_a, lr)
        else:
            sr = O.espcn(sd_a, F.interpolate(luma, scale_factor=0.5), 2)
        out = O.resdeconv(sd_b, sr)
        assert math.isclose(row["MSE"], float(O.mse_loss(out, b)), rel_tol=TOL)
        assert abs(row["PSNR"] - float(O.psnr(out, b))) < 1e-2
        assert math.isclose(row["SSIM"], float(O.ssim(out, b)), rel_tol=5e-3, abs_tol=1e-4)
        assert math.isclose(row["AE"], float(O.angular_error(out, b).mean()), rel_tol=TOL)


@pytest.mark.parametrize("which", ["RDDBNet_x2", "RDDBNet_x4", "SRDN"])
def test_cascade_generators_parity(golden_cascade, which):
    """package RDDBNet (k2 s2 transposed-conv upsampling) and SRDN on the CUDA path, fp32 mode."""
    from oracle import srcgan_oracle as O
    from srcgan_b200 import nn as snn
    fx = golden_cascade[which]
    if which == "SRDN":
        sd, net, fn, x, seed = O.init_srdn(32), snn.SRDN(1, 1, 2), (lambda s, t: O.srdn(s, t)), rand((2, 1, 16, 12), 302), 18
    else:
        up = int(which[-1])
        sd, net, x, seed = O.init_rddbnet_pkg(31, 1, 1, up), snn.RDDBNet(1, 1, up), rand((2, 1, 16, 12), 301), 17
        fn = lambda s, t: O.rddbnet_pkg(s, t, up)
    y, named, dx, r32, r64 = run_both(net, fn, sd, x, seed)
    assert relerr(y, r64.y) < 1e-4 and relerr(y, fx["out"]) < 1e-4
    # SRDN stacks 18 dense blocks (72 LeakyReLU layers): one flipped kink in the last blocks measures
    # 2e-3..7e-3 on the neighbouring layers (scripts/diag_srdn.py) -> judged at the cap
    floor = GRAD_CAP if which == "SRDN" else GRAD_FLOOR
    check_grads(named, r32, r64, floor=floor)
    assert l2err(dx, r64.dx) < grad_tol(r32.dx, r64.dx, floor)
    # bf16 mode runs and stays close
    snn.set_precision("bf16")
    with torch.no_grad():
        yb = net(x.to(DEV))
    assert relerr(yb, r64.y) < 0.1


@pytest.mark.parametrize("which", ["ESPCN_x2", "ESPCN_x4_rgb", "SRCNN"])
def test_plain_conv_stacks_parity(golden_cascade, which):
    """ESPCN (5x5 conv, PixelShuffle) and SRCNN (9x9 / 1x1 / 5x5, trailing ReLU) on the CUDA path."""
    from oracle import srcgan_oracle as O
    from srcgan_b200 import nn as snn
    from tests.test_oracle import cascade_case
    fx = golden_cascade[which]
    sd, fn, x, seed = cascade_case(which)
    net = {"ESPCN_x2": lambda: snn.ESPCN(1, 1, 2), "ESPCN_x4_rgb": lambda: snn.ESPCN(3, 3, 4),
           "SRCNN": lambda: snn.SRCNN(1, 3, 2)}[which]()
    y, named, dx, r32, r64 = run_both(net, fn, sd, x, seed)
    assert relerr(y, r64.y) < 1e-4 and relerr(y, fx["out"]) < 1e-4
    check_grads(named, r32, r64)
    assert l2err(dx, r64.dx) < grad_tol(r32.dx, r64.dx)
    snn.set_precision("bf16")
    with torch.no_grad():
        yb = net(x.to(DEV))
    assert relerr(yb, r64.y) < 0.1


def zoo_net(which):
    from srcgan_b200 import zoo
    return {"ResDeconv": lambda: zoo.ResDeconv(1, 3), "EDSR_x2": lambda: zoo.EDSR(1, 1, 2, num_residuals=3),
            "EDSR_x4_rgb": lambda: zoo.EDSR(3, 3, 4, num_residuals=2),
            "SRDenseNetA_x2": lambda: zoo.SRDenseNetA(1, 3, mode="x2", num_blocks=2, num_layers=2),
            "SRDenseNetA_x4": lambda: zoo.SRDenseNetA(1, 3, mode="x4", num_blocks=2, num_layers=2),
            "SRDenseNetB_x2": lambda: zoo.SRDenseNetB(3, 1, mode="x2", num_blocks=2, num_layers=2),
            "SRDenseNetB_x4": lambda: zoo.SRDenseNetB(3, 1, mode="x4", num_blocks=2, num_layers=2)}[which]()


@pytest.mark.parametrize("which", ["ResDeconv", "EDSR_x2", "EDSR_x4_rgb", "SRDenseNetA_x2", "SRDenseNetA_x4",
                                   "SRDenseNetB_x2", "SRDenseNetB_x4"])
def test_zoo_generators_parity(golden_zoo, which):
    """ResDeconv (GroupNorm ResNet-18 + k2 s2 deconvs), EDSR (shared-GN residual blocks), SRDenseNetA/B (shared
    k3 s2 transposed / strided resampler) on the CUDA path: fp32 mode vs oracle + the real reference's outputs."""
    from srcgan_b200 import nn as snn
    from tests.test_oracle import zoo_case
    fx = golden_zoo[which]
    sd, fn, x, seed = zoo_case(which)
    net = zoo_net(which)
    y, named, dx, r32, r64 = run_both(net, fn, sd, x, seed)
    assert relerr(y, r64.y) < 1e-4 and relerr(y, fx["out"]) < 1e-4
    check_grads(named, r32, r64)
    assert l2err(dx, r64.dx) < grad_tol(r32.dx, r64.dx)
    snn.set_precision("bf16")
    with torch.no_grad():
        yb = net(x.to(DEV))
    assert relerr(yb, r64.y) < 0.1
    # bf16 training step runs through the tensor-core paths (wide layers) without refusing a shape
    net.zero_grad()
    xb = x.to(DEV).requires_grad_(True)
    net(xb).sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


@pytest.mark.parametrize("which,shape", [("ResDeconv", (2, 1, 64, 96)), ("EDSR_x4_rgb", (1, 3, 24, 20))])
def test_operator_layer_thin_tensors_bf16(which, shape, monkeypatch):
    """bf16 mode of the operator layer (functional.py): thin tensors in pitch-8 buffers put the image convolutions on the tcgen05
    kernels and the <= 4-input-channel stem on thin_in_tiled; outputs, input gradient and every parameter gradient agree with the
    dense-pitch / old-kernel path (same bf16 operands, fp32 accumulation on both sides)."""
    from srcgan_b200 import nn as snn
    snn.set_precision("bf16")
    torch.manual_seed(5)
    net = zoo_net(which).to(DEV)
    x = rand(shape, 77).to(DEV)
    gy = None

    def run():
        nonlocal gy
        net.zero_grad()
        xin = x.clone().requires_grad_(True)
        y = net(xin)
        if gy is None:
            gy = rand(tuple(y.shape), 78).to(DEV)
        y.backward(gy)
        torch.cuda.synchronize()
        return y.detach().clone(), xin.grad.clone(), {k: p.grad.clone() for k, p in net.named_parameters()}

    y1, dx1, g1 = run()
    monkeypatch.setenv("SRCGAN_B200_NO_THIN_PITCH", "1")
    monkeypatch.setenv("SRCGAN_B200_NO_THIN_TILED", "1")
    monkeypatch.setenv("SRCGAN_B200_THIN_WGRAD7_ROWS", "1")
    y0, dx0, g0 = run()
    assert relerr(y1, y0) < 2e-2 and l2err(dx1, dx0) < 2e-2
    for k in g0:
        assert l2err(g1[k], g0[k]) < 2e-2, k


@pytest.mark.parametrize("variant", ["", "ConstLAB"])
def test_cascade_step_against_golden(golden_cas_step, variant):
    """trainer_cas.CasSRC (mirror of trainCas*.py) on the CUDA path vs the real reference's iterations."""
    from oracle import srcgan_oracle as O
    from srcgan_b200 import trainer_cas
    opt = trainer_cas.params()
    opt.device = torch.device(DEV)
    opt.up, opt.variant = 2, variant
    opt.SRModel, opt.CModel = ("SRCNN" if "Const" in variant else "ESPCN"), "SRCNN"
    m = trainer_cas.CasSRC(opt)
    m.netG_A2C.load_state_dict(O.init_srcnn(51, 1, 1) if "Const" in variant else O.init_espcn(51, 1, 1, 2), strict=True)
    m.netG_C2B.load_state_dict(O.init_srcnn(52, 1, 2 if "LAB" in variant else 3), strict=True)
    for it, rec in enumerate(golden_cas_step[variant or "plain"]):
        m.init_log()
        m.optimize_parameters(rand((2, 1, 32, 32), 700 + it).to(DEV), rand((2, 3, 32, 32), 600 + it).to(DEV))
        got = m.log_values()
        for k, g in (("loss_SR", "loss_sr"), ("loss_C", "loss_c"), ("psnr_SR", "psnr_sr"), ("psnr_C", "psnr_c")):
            assert math.isclose(got[g], rec[k], rel_tol=TOL, abs_tol=1e-5), (variant, it, k, got[g], rec[k])
        assert relerr(m.fake_AB, rec["fake_AB"]) < TOL
    with pytest.raises(NotImplementedError):
        trainer_cas.build_model("NoSuchNet", 1, 3)


def test_batched_repack_matches_single_packs():
    """csrc/pack_batch.cu (one launch per network) writes bit for bit what srcgan_pack_weights writes from the torch-side
    flip / cat / scale expressions: forward, transposed and mirrored-dense-block tensors of G_B and the discriminator; after an
    in-place parameter update ONE launch refreshes everything; invalidate_packs() covers writes through ``.data``."""
    from oracle import srcgan_oracle as O
    from srcgan_b200 import nn as snn, ops
    snn.set_precision("bf16")

    def legacy(e):
        parts = []
        for (p, transposed, n_lo, k_lo, k_len, scale) in e.blocks:
            w = p.detach()
            blk = snn._wT(w[k_lo:k_lo + k_len, n_lo:n_lo + e.cout]) if transposed else w[n_lo:n_lo + e.cout, k_lo:k_lo + k_len]
            parts.append(blk * scale)
        return ops.pack_weights(torch.cat(parts, dim=1).contiguous(), ops.WL_TC, torch.bfloat16)

    for net, sd, shape in ((snn.RDDBNetA(3, 3, 64, nb=3, mode="x4"), O.init_rddbnet_a(12), (1, 3, 128, 128)),
                           (snn.NLayerDiscriminator(3, 64, 2), O.init_discriminator(13), (2, 3, 64, 64))):
        net.load_state_dict(sd)
        net.to(DEV)
        x = rand(shape, 5).to(DEV).requires_grad_(True)
        net(x).sum().backward()
        pk = net._pk()
        assert len(pk.tc) >= (90 if isinstance(net, snn.RDDBNetA) else 1)
        for key, e in pk.tc.items():
            assert torch.equal(e.dest, legacy(e)), key
        before = pk.repacks
        with torch.no_grad():
            for p in net.parameters():
                p.add_(0.01 * torch.randn_like(p))               # what an optimizer step does (bumps the version)
        net(x).sum().backward()
        assert pk.repacks == before + 1                          # ONE launch refreshed every tensor
        for key, e in pk.tc.items():
            assert torch.equal(e.dest, legacy(e)), key
        # a write through .data is invisible to the version counter: stale until invalidate_packs()
        first = next(iter(pk.tc.values()))
        first.params[0].data.mul_(2.0)
        net(x)
        assert not torch.equal(first.dest, legacy(first))
        net.invalidate_packs()
        net(x)
        assert torch.equal(first.dest, legacy(first))


def test_gradients_are_bucket_views_and_accumulate_in_place():
    """Every parameter gradient of a network is a view of ONE flat buffer (what the data-parallel all-reduce sends); a
    network used twice in one backward pass adds in place and still matches two separate passes."""
    from oracle import srcgan_oracle as O
    from srcgan_b200 import nn as snn
    net = snn.NLayerDiscriminator(3, 64, 2)
    net.load_state_dict(O.init_discriminator(13))
    net.to(DEV)
    a, b = rand((2, 3, 64, 64), 1).to(DEV), rand((2, 3, 64, 64), 2).to(DEV)
    (net(a).square().mean() + net(b).square().mean()).backward()          # backward_D_basic's shape: real + fake
    bucket = net.grad_bucket()
    assert all(bucket.holds(p) for p in net.parameters())
    flat = bucket.flat
    assert flat.numel() >= sum(p.numel() for p in net.parameters())
    got = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    net.load_state_dict(O.init_discriminator(13))                          # reset the BN buffers
    for p in net.parameters():
        p.grad = None
    net(a).square().mean().backward()
    ga = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    for p in net.parameters():
        p.grad = None
    net(b).square().mean().backward()
    for k, p in net.named_parameters():
        assert l2err(got[k], ga[k] + p.grad) < 1e-5, k


def test_tall_image_mode_for_small_maps(monkeypatch):
    """Maps below 96 px (the G_A trunk at 64x64) run on the paired tcgen05 sweep by stacking the batch into one tall image with
    zero separator rows (nn._tall_plan).  Same network, same inputs: tall mode vs the small-tile kernels vs the fp32 oracle."""
    from oracle import srcgan_oracle as O
    from srcgan_b200 import _lib, nn as snn
    snn.set_precision("bf16")
    sd = O.init_rddbnet_b(21)
    x = rand((6, 3, 24, 40), 211)                       # 6 * 25 + 1 = 151 rows >= 128; ragged last strip; w = 40
    pr = None
    res = {}
    for mode in ("tall", "plain"):
        if mode == "plain":
            monkeypatch.setenv("SRCGAN_B200_NO_TALL", "1")
        net = snn.RDDBNetB(3, 3, 64, nb=3, mode="x4")
        net.load_state_dict(sd)
        net.to(DEV)
        assert (net._tall_plan(6, 24, 40, 3, torch.bfloat16) > 0) == (mode == "tall")
        xg = x.to(DEV).requires_grad_(True)
        y = net(xg)
        kernels_seen = _lib.last_kernel()
        pr = probe_like(y, 5) if pr is None else pr
        (y * pr.to(DEV)).sum().backward()
        res[mode] = (y.detach().cpu(), {k: p.grad.detach().cpu() for k, p in net.named_parameters()}, xg.grad.cpu())
    r32 = Ref(lambda s, t: O.rddbnet_b(s, t, "x4"), sd, x, pr, torch.float32)
    yt, gt, dxt = res["tall"]
    yp, gp, dxp = res["plain"]
    e_t, e_p = l2err(yt, r32.y), l2err(yp, r32.y)
    assert e_t < max(2e-2, 1.5 * e_p), (e_t, e_p)
    assert l2err(dxt, r32.dx) < max(6e-2, 1.5 * l2err(dxp, r32.dx))
    worse = 0
    for k, v in r32.sd.items():
        if O.is_buffer_key(k) or v.grad is None:
            continue
        a, b = l2err(gt[k], v.grad), l2err(gp[k], v.grad)
        assert a < max(8e-2, 2.0 * b), (k, a, b)
        worse += a > b
    # neither path is systematically closer to the oracle than the other
    assert worse < 0.8 * len(gt)


def test_tall_image_conv_kernel_separators():
    """One 3x3 layer on a tall image: equals the per-image convolution, separator rows come out as exact zeros."""
    from srcgan_b200 import ops
    n, h, w, cin, cout = 5, 31, 48, 64, 32
    x = rand((n, cin, h, w), 1) - 0.5
    wt, b = (rand((cout, cin, 3, 3), 2) - 0.5) * 0.2, rand((cout,), 3)
    xb = x.to(torch.bfloat16).float()
    want = F.leaky_relu(F.conv2d(xb, wt.to(torch.bfloat16).float(), b, padding=1), 0.2)
    period = h + 1
    tall = torch.zeros((1, n * period + 1, w, cin), dtype=torch.bfloat16, device=DEV)
    tall[0, 1:].view(n, period, w, cin)[:, :h].copy_(xb.permute(0, 2, 3, 1).to(DEV))
    out = torch.full((1, n * period + 1, w, 96), 7.0, dtype=torch.bfloat16, device=DEV)
    bits = torch.full((1, n * period + 1, w, 1), -1, dtype=torch.int32, device=DEV)
    wp = ops.pack_weights(wt.to(DEV), ops.WL_TC, torch.bfloat16)
    ops.conv_fprop(ops.Slice(tall), wp, b.to(DEV), ops.Slice(out, 32, cout), 3, 1, 1, act=0.2, engine=ops.ENGINE_TC,
                   signbits=bits, zero_rows=period)
    from srcgan_b200 import _lib
    assert _lib.last_kernel().startswith("conv3x3_sweep2_tc")
    got = out[0, 1:].view(n, period, w, 96)[:, :h, :, 32:64].float().permute(0, 3, 1, 2).cpu()
    assert relerr(got, want) < 1e-2
    sep = out[0, ::period, :, 32:64]
    assert float(sep.abs().max()) == 0.0 and int(bits[0, ::period].abs().max()) == 0
    assert float((out[..., :32] - 7.0).abs().max()) == 0.0 and float((out[..., 64:] - 7.0).abs().max()) == 0.0
    sb = bits[0, 1:].view(n, period, w)[:, :h].cpu()
    wantbits = (want.to(torch.bfloat16) > 0).permute(0, 2, 3, 1).to(torch.int64)
    packed = (wantbits << torch.arange(32, dtype=torch.int64)).sum(-1)
    agree = ((sb.to(torch.int64) & 0xFFFFFFFF) == packed).float().mean()
    assert float(agree) > 0.999          # (a value that rounds to +-0 in bf16 may differ)


@pytest.mark.gpu
def test_device_prefetcher_yields_every_batch_once_in_order():
    """srcgan_b200.data.DevicePrefetcher: the copies run one batch ahead on a side stream; values, order and count are those of
    the host iterable (tensors and tuples)."""
    from srcgan_b200 import data
    host = [(torch.full((3, 5), float(i)).pin_memory(), torch.arange(4) + i) for i in range(5)]
    got = list(data.DevicePrefetcher(host, "cuda:0"))
    assert len(got) == 5
    for i, (a, b) in enumerate(got):
        assert a.is_cuda and b.is_cuda
        assert torch.equal(a.cpu(), host[i][0]) and torch.equal(b.cpu(), host[i][1])
    single = list(data.DevicePrefetcher([torch.ones(2).pin_memory()], "cuda:0"))
    assert len(single) == 1 and torch.equal(single[0].cpu(), torch.ones(2))
