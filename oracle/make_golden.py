"""Generates tests/golden/*.pt by running the REAL reference (imported from /root/reference,
never copied) on seeded inputs.  Run in the build container:  python oracle/make_golden.py

Fixtures hold only outputs (inputs and weights are regenerated from seeds through
``oracle/srcgan_oracle.py``'s deterministic init + ``load_state_dict(strict=True)``):
  modules_tiny.pt   per-module outputs, per-parameter gradient norms, BN buffers,
                    loss / metric known answers
  zoo_tiny.pt       ResDeconv / EDSR / SRDenseNetA,B outputs, gradient norms, input gradients
  step_gray_tiny.pt the same two iterations for opt.net = '2' (gray LR, RDDB pair) and 'SRdens' (SRDenseNet pair)
  step_tiny.pt      two consecutive ``train.SRCycleGAN.optimize_parameters`` calls
                    (x4, net '1', B=2, 16x16 -> 64x64): the 9 losses per step, parameter
                    norms after the Adam updates, BN buffers
"""
from __future__ import annotations

import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_harness, srcgan_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def probe_like(t: torch.Tensor, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(t.shape, generator=g)


def grad_norms(module) -> dict:
    return {k: float(p.grad.double().norm()) for k, p in module.named_parameters() if p.grad is not None}


def buffers(module) -> dict:
    return {k: v.clone() for k, v in module.state_dict().items() if O.is_buffer_key(k)}


def rand(shape, seed):
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed))


def modules_fixture() -> dict:
    _pkg, M, losses, metrics = ref_harness.import_reference()
    RDDBNetA = ref_harness.make_rddbneta_shim(M)
    fx = {}

    # G_A : RDDBNetB x4 and x2
    for mode in ("x4", "x2"):
        net = M.RDDBNetB(3, 3, 64, nb=3, mode=mode)
        net.load_state_dict(O.init_rddbnet_b(11), strict=True)
        x = rand((2, 3, 16, 16), 101).requires_grad_(True)
        y = net(x)
        (y * probe_like(y, 7)).sum().backward()
        fx[f"G_A_{mode}"] = {"out": y.detach().clone(), "grad_norms": grad_norms(net),
                             "dx": x.grad.clone()}

    # G_B : RDDBNetA shim
    net = RDDBNetA(3, 3, 64, nb=3, mode="x4")
    net.load_state_dict(O.init_rddbnet_a(12), strict=True)
    x = rand((2, 3, 32, 32), 102).requires_grad_(True)
    y = net(x)
    (y * probe_like(y, 8)).sum().backward()
    fx["G_B"] = {"out": y.detach().clone(), "grad_norms": grad_norms(net), "dx": x.grad.clone(),
                 "buffers": buffers(net)}
    net.eval()
    fx["G_B"]["out_eval"] = net(x.detach()).detach().clone()

    # D
    net = M.NLayerDiscriminator(3, 64, 2)
    net.load_state_dict(O.init_discriminator(13), strict=True)
    x = rand((3, 3, 64, 64), 103).requires_grad_(True)
    y = net(x)
    (y * probe_like(y, 9)).sum().backward()
    fx["D"] = {"out": y.detach().clone(), "grad_norms": grad_norms(net), "dx": x.grad.clone(),
               "buffers": buffers(net)}

    # RDB5 / RRDB alone (model.py:193-233)
    rr = M.RRDB(64, 32)
    sd = {k[len("RRDB_trunk.0."):]: v for k, v in O.init_rddbnet_b(14).items() if k.startswith("RRDB_trunk.0.")}
    rr.load_state_dict(sd, strict=True)
    x = (rand((1, 64, 12, 10), 104) - 0.5)
    fx["RRDB"] = {"out": rr(x).detach().clone(), "rdb1_out": rr.RDB1(x).detach().clone()}

    # losses / metrics known answers
    a = rand((2, 3, 40, 36), 105)
    b = (a + 0.05 * torch.randn(a.shape, generator=torch.Generator().manual_seed(106))).clamp(0, 1)
    a255 = a * 255.0
    fx["scalars"] = {
        "L1": float(losses.L1Loss()(a, b)), "MSE": float(losses.MSELoss()(a, b)),
        "PSNRLoss": float(losses.PSNRLoss()(a, b)), "DSSIM": float(losses.DSSIMLoss()(a, b)),
        "SSIM": float(metrics.SSIM()(a, b)), "SSIM_255": float(metrics.SSIM()(a255, b * 255.0)),
        "SSIM_neg": float(metrics.SSIM()(a * 2 - 1, b * 2 - 1)),
        "SSIM_self": float(metrics.SSIM()(a, a)),
        "SSIM_per_image": metrics.SSIM()(a, b, size_average=False).clone(),
        "PSNR": float(metrics.PSNR()(a, b)), "MSEm": float(metrics.MSE()(a, b)),
        "AE": metrics.AE()(a, b).clone(),
    }
    return fx


def cascade_fixture() -> dict:
    """Generators of the cascaded trainers that share the dense-block kernels: package RDDBNet
    (src/model/rddb.py) for up = 2, 4 and SRDN (src/model/srdn.py), as built by trainCas.py:30."""
    pkg, _M, _losses, _metrics = ref_harness.import_reference()
    fx = {}
    for up in (2, 4):
        net = pkg.RDDBNet(1, 1, up)
        net.load_state_dict(O.init_rddbnet_pkg(31, 1, 1, up), strict=True)
        x = rand((2, 1, 16, 12), 301).requires_grad_(True)
        y = net(x)
        (y * probe_like(y, 17)).sum().backward()
        fx[f"RDDBNet_x{up}"] = {"out": y.detach().clone(), "grad_norms": grad_norms(net), "dx": x.grad.clone()}
    net = pkg.SRDN(1, 1, 2)
    net.load_state_dict(O.init_srdn(32), strict=True)
    x = rand((2, 1, 16, 12), 302).requires_grad_(True)
    y = net(x)
    (y * probe_like(y, 18)).sum().backward()
    fx["SRDN"] = {"out": y.detach().clone(), "grad_norms": grad_norms(net), "dx": x.grad.clone()}
    # plain conv stacks: ESPCN (5x5 + PixelShuffle) and SRCNN (9x9 / 1x1 / 5x5, trailing ReLU)
    for name, net, sd in (("ESPCN_x2", pkg.ESPCN(1, 1, 2), O.init_espcn(33, 1, 1, 2)),
                          ("ESPCN_x4_rgb", pkg.ESPCN(3, 3, 4), O.init_espcn(34, 3, 3, 4)),
                          ("SRCNN", pkg.SRCNN(1, 3, 2), O.init_srcnn(35, 1, 3))):
        net.load_state_dict(sd, strict=True)
        cin = sd["conv1.weight"].shape[1]
        x = rand((2, cin, 14, 10), 303).requires_grad_(True)
        y = net(x)
        (y * probe_like(y, 19)).sum().backward()
        fx[name] = {"out": y.detach().clone(), "grad_norms": grad_norms(net), "dx": x.grad.clone()}
    return fx


def zoo_cases():
    """name -> (reference ctor(pkg, M), state_dict, oracle fn, input, probe seed); shared with the tests."""
    return {
        "ResDeconv": (lambda pkg, M: pkg.ResDeconv(1, 3), O.init_resdeconv(41, 3), lambda s, t: O.resdeconv(s, t),
                      rand((2, 1, 32, 32), 401), 21),
        "EDSR_x2": (lambda pkg, M: pkg.EDSR(1, 1, 2, num_residuals=3), O.init_edsr(42, 1, 1, 2, num_residuals=3),
                    lambda s, t: O.edsr(s, t, 2, 3), rand((2, 1, 16, 12), 402), 22),
        "EDSR_x4_rgb": (lambda pkg, M: pkg.EDSR(3, 3, 4, num_residuals=2), O.init_edsr(43, 3, 3, 4, num_residuals=2),
                        lambda s, t: O.edsr(s, t, 4, 2), rand((2, 3, 8, 8), 403), 23),
        "SRDenseNetA_x2": (lambda pkg, M: M.SRDenseNetA(1, 3, mode="x2", num_blocks=2, num_layers=2),
                           O.init_srdensenet(44, "A", 1, 3), lambda s, t: O.srdensenet(s, t, "A", "x2"),
                           rand((2, 1, 16, 12), 404), 24),
        "SRDenseNetA_x4": (lambda pkg, M: M.SRDenseNetA(1, 3, mode="x4", num_blocks=2, num_layers=2),
                           O.init_srdensenet(45, "A", 1, 3), lambda s, t: O.srdensenet(s, t, "A", "x4"),
                           rand((2, 1, 8, 8), 405), 25),
        "SRDenseNetB_x2": (lambda pkg, M: M.SRDenseNetB(3, 1, mode="x2", num_blocks=2, num_layers=2),
                           O.init_srdensenet(46, "B", 3, 1), lambda s, t: O.srdensenet(s, t, "B", "x2"),
                           rand((2, 3, 32, 24), 406), 26),
        "SRDenseNetB_x4": (lambda pkg, M: M.SRDenseNetB(3, 1, mode="x4", num_blocks=2, num_layers=2),
                           O.init_srdensenet(47, "B", 3, 1), lambda s, t: O.srdensenet(s, t, "B", "x4"),
                           rand((2, 3, 32, 32), 407), 27),
    }


def zoo_fixture() -> dict:
    """ResDeconv (resdeconv.py), EDSR (edsr.py), SRDenseNetA/B (model.py:659-778) run by the REAL reference."""
    pkg, M, _losses, _metrics = ref_harness.import_reference()
    fx = {}
    for name, (ctor, sd, _fn, x, seed) in zoo_cases().items():
        net = ctor(pkg, M)
        net.load_state_dict(sd, strict=True)
        x = x.clone().requires_grad_(True)
        y = net(x)
        (y * probe_like(y, seed)).sum().backward()
        fx[name] = {"out": y.detach().clone(), "grad_norms": grad_norms(net), "dx": x.grad.clone()}
    return fx


def step_fixture(net: str = "1") -> dict:
    train = ref_harness.import_train()
    opt = train.params()
    opt.device = torch.device("cpu")
    opt.mode = "x4"
    opt.net = net
    torch.manual_seed(0)
    model = train.SRCycleGAN(opt)
    st = O.default_states(0, net)
    model.netG_A.load_state_dict(st["G_A"], strict=True)
    model.netG_B.load_state_dict(st["G_B"], strict=True)
    model.netD_A.load_state_dict(st["D_A"], strict=True)
    model.netD_B.load_state_dict(st["D_B"], strict=True)
    random.seed(5)
    steps = []
    for it in range(2):
        real_A, real_B = (O.synthetic_batch if net == "1" else O.synthetic_gray_batch)(2, lr=16, scale=4, seed=1234 + it)
        model.optimize_parameters(real_A, real_B)
        rec = {"losses": {n: float(getattr(model, "loss_" + n)) for n in O.CycleGANStepOracle.LOSS_NAMES}}
        if it == 0:
            rec["fake_B"] = model.fake_B.detach().clone()
            rec["fake_A"] = model.fake_A.detach().clone()
        steps.append(rec)
    out = {"steps": steps, "param_norms": {}, "buffers": {}}
    for name in ("G_A", "G_B", "D_A", "D_B"):
        net = getattr(model, "net" + name)
        out["param_norms"][name] = {k: float(p.detach().double().norm()) for k, p in net.named_parameters()}
        out["buffers"][name] = buffers(net)
    return out


def cas_step_fixture() -> dict:
    """Two iterations of the REAL ``trainCas*.CasSRC.optimize_parameters``: plain variant with ESPCN x2 + SRCNN
    colouriser, ConstLAB variant (full-resolution SR input, LAB targets) with SRCNN + SRCNN as in runConstLAB.sh:
    losses, PSNRs, transfer outputs."""
    out = {}
    for variant in ("", "ConstLAB"):
        mod = ref_harness.import_traincas(variant)
        opt = mod.params()
        opt.device = torch.device("cpu")
        opt.up, opt.SRModel, opt.CModel = 2, ("SRCNN" if "Const" in variant else "ESPCN"), "SRCNN"
        model = mod.CasSRC(opt)
        lab = "LAB" in variant
        model.netG_A2C.load_state_dict(O.init_srcnn(51, 1, 1) if "Const" in variant else O.init_espcn(51, 1, 1, 2),
                                       strict=True)
        model.netG_C2B.load_state_dict(O.init_srcnn(52, 1, 2 if lab else 3), strict=True)
        model.init_log()
        steps = []
        for it in range(2):
            real_B = rand((2, 3, 32, 32), 600 + it)
            real_A = rand((2, 1, 32, 32), 700 + it)
            model.optimize_parameters(real_A, real_B)
            steps.append({"loss_SR": float(model.loss_SR), "loss_C": float(model.loss_C),
                          "psnr_SR": float(model.psnr_SR), "psnr_C": float(model.psnr_C),
                          "fake_AB": model.fake_AB.detach().clone()})
        out[variant or "plain"] = steps
    return out


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    torch.save(modules_fixture(), os.path.join(OUT, "modules_tiny.pt"))
    torch.save(step_fixture(), os.path.join(OUT, "step_tiny.pt"))
    torch.save({net: step_fixture(net) for net in ("2", "SRdens")}, os.path.join(OUT, "step_gray_tiny.pt"))
    torch.save(cascade_fixture(), os.path.join(OUT, "cascade_tiny.pt"))
    torch.save(cas_step_fixture(), os.path.join(OUT, "cas_step_tiny.pt"))
    torch.save(zoo_fixture(), os.path.join(OUT, "zoo_tiny.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
