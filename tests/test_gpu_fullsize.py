"""Parity AT THE BENCHMARKED CONFIGURATION: one full-size G+D step (64x64 -> 256x256 patches, where every dense-block
layer of G_B runs on the paired tcgen05 sweep / stacked wgrad kernels) on the CUDA path against the fp32 CPU oracle on the
same weights and inputs - losses, generated images (PSNR), and EVERY parameter gradient of G_A, G_B, D_A, D_B.

Batch 1 keeps the oracle step at a few seconds of host time; the per-patch arithmetic is the benchmark's.  The measured
statistics are also written to gpurun_out/fullsize_parity_<precision>.json (copied to profiles/ by hand)."""
import json
import math
import os
import random

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def psnr(a, b):
    return float(10 * torch.log10(1.0 / ((a.double().cpu() - b.double().cpu()) ** 2).mean()))


def _clone_states(states, dt=None, dev=None):
    out = {}
    for n, sd in states.items():
        out[n] = {k: (v.to(dt) if (dt is not None and v.is_floating_point()) else v.clone()).to(dev or v.device) for k, v in sd.items()}
    return out


@pytest.fixture(scope="module")
def oracle_step():
    """(states, inputs, fp32 oracle, fp64 oracle) after one optimize_parameters each - shared by the two precisions."""
    from oracle import srcgan_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    states = O.default_states(0)
    real_A, real_B = O.synthetic_batch(1, lr=64, scale=4, seed=4321)
    random.seed(3)
    ref = O.CycleGANStepOracle(_clone_states(states))
    losses = ref.optimize_parameters(real_A, real_B)
    random.seed(3)
    ref64 = O.CycleGANStepOracle(_clone_states(states, torch.float64))        # the "truth" the fp32 evaluations are judged by
    ref64.optimize_parameters(real_A.double(), real_B.double())
    return states, real_A, real_B, ref, losses, ref64


def _grad_table(get_grad, truth_nets, O):
    """per parameter tensor: cosine / L2 error / fraction of elements within 1e-3 of the tensor's max, against ``truth``"""
    rows = {}
    for net, sd in truth_nets.items():
        for k, p in sd.items():
            if O.is_buffer_key(k) or p.grad is None:
                continue
            g = get_grad(net, k)
            if g is None:
                rows["%s.%s" % (net, k)] = None
                continue
            g, r = g.detach().double().cpu().flatten(), p.grad.detach().double().cpu().flatten()
            rows["%s.%s" % (net, k)] = {
                "cos": float(torch.dot(g, r) / (g.norm() * r.norm()).clamp_min(1e-300)),
                "l2": float((g - r).norm() / r.norm().clamp_min(1e-300)),
                "within": float(((g - r).abs() <= 1e-3 * r.abs().max()).double().mean()), "numel": g.numel()}
    return rows


def _summary(rows):
    vs = [v for v in rows.values() if v is not None]
    l2s, coss = sorted(v["l2"] for v in vs), sorted(v["cos"] for v in vs)
    tot = sum(v["numel"] for v in vs)
    return {"tensors": len(vs), "median_l2": l2s[len(l2s) // 2], "worst_l2": l2s[-1], "median_cos": coss[len(coss) // 2],
            "worst_cos": coss[0], "frac_elements_within_1e-3_of_tensor_max": sum(v["within"] * v["numel"] for v in vs) / tot}


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_full_size_step_matches_the_oracle(oracle_step, precision):
    """Tolerances.  Forward quantities (nine losses, four generated images) are held to north_star's numbers: fp32 1e-3,
    bf16 PSNR within 0.05 dB.  Gradients of this random-init network are ILL-CONDITIONED at full size - LeakyReLU gates whose
    pre-activation lies within rounding of zero flip under any change of summation order or precision, and a flipped gate
    changes that pixel's contribution by 80 % - so every floating-point evaluation, the reference's own included, carries a
    noise floor that is measured here and used as the yardstick:
      fp32: error against an fp64 evaluation of the oracle; >= 99.5 % of all gradient elements within 1e-3 of their tensor's
            maximum, median tensor error <= 1.2e-3, every tensor within 3x the REFERENCE's own fp32 error;
      bf16: error against the fp32 oracle; must stay within 1.5x what STOCK PyTorch bf16 autocast (cuDNN) shows on the same
            step, same weights, same inputs."""
    from oracle import srcgan_oracle as O
    from srcgan_b200 import nn as snn, trainer
    states, real_A, real_B, ref, ref_losses, ref64 = oracle_step
    old = snn.precision()
    snn.set_precision(precision)
    try:
        opt = trainer.params()
        opt.device, opt.mode, opt.net = torch.device(DEV), "x4", "1"
        m = trainer.SRCycleGAN(opt)
        for name in ("G_A", "G_B", "D_A", "D_B"):
            getattr(m, "net" + name).load_state_dict(states[name], strict=True)
        random.seed(3)
        m.optimize_parameters(real_A.to(DEV), real_B.to(DEV))
        torch.cuda.synchronize()
        got = m.current_losses()
    finally:
        snn.set_precision(old)
    bf = precision == "bf16"
    nets = ("G_A", "G_B", "D_A", "D_B")
    named = {n: dict(getattr(m, "net" + n).named_parameters()) for n in nets}
    report = {"precision": precision, "losses": {}, "images": {}}
    snr = lambda a, b: float(20 * torch.log10(b.double().norm() / (a.double().cpu() - b.double().cpu()).norm().clamp_min(1e-300)))
    for n, v in ref_losses.items():
        report["losses"][n] = {"cuda": got[n], "oracle": v, "rel": abs(got[n] - v) / max(abs(v), 1e-30)}
    for name, tgt in (("fake_B", real_B), ("fake_A", real_A), ("recl_B", real_B), ("recl_A", real_A)):
        a, b = getattr(m, name).detach(), getattr(ref, name).detach()
        report["images"][name] = {"psnr_cuda_vs_data": psnr(a, tgt), "psnr_oracle_vs_data": psnr(b, tgt),
                                  "snr_cuda_vs_oracle_db": snr(a, b)}
    truth = {n: getattr(ref64 if not bf else ref, n) for n in nets}
    ours = _grad_table(lambda n, k: named[n][k].grad, truth, O)
    if bf:
        # yardstick: the same step through stock PyTorch (the oracle's functional torch code on cuda under bf16 autocast)
        random.seed(3)
        ac = O.CycleGANStepOracle(_clone_states(states, dev=DEV))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ac.optimize_parameters(real_A.to(DEV), real_B.to(DEV))
        yard = _grad_table(lambda n, k: getattr(ac, n)[k].grad, truth, O)
        yname = "torch_bf16_autocast_vs_fp32_oracle"
        for name in ("fake_B", "fake_A", "recl_B", "recl_A"):
            report["images"][name]["snr_torch_autocast_vs_oracle_db"] = snr(getattr(ac, name).detach().float(), getattr(ref, name).detach())
    else:
        yard = _grad_table(lambda n, k: getattr(ref, n)[k].grad, truth, O)
        yname = "reference_fp32_vs_fp64_oracle"
    report["grads"] = {k: {"cuda": ours[k], yname: yard.get(k)} for k in ours}
    report["summary"] = {"cuda_vs_%s" % ("fp32_oracle" if bf else "fp64_oracle"): _summary(ours), yname: _summary(yard),
                         "worst_loss_rel": max(v["rel"] for v in report["losses"].values()),
                         "min_image_snr_db": min(v["snr_cuda_vs_oracle_db"] for v in report["images"].values())}
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "fullsize_parity_%s.json" % precision), "w") as f:
            json.dump(report, f, indent=1)
    print(json.dumps(report["summary"]))
    # ---- the bar
    assert all(v is not None for v in ours.values()), [k for k, v in ours.items() if v is None]
    for n, v in report["losses"].items():       # 1. the nine losses
        assert math.isclose(v["cuda"], v["oracle"], rel_tol=2e-2 if bf else 1e-3, abs_tol=2e-3 if bf else 1e-5), (n, v)
    for name, v in report["images"].items():    # 2. generated images
        assert abs(v["psnr_cuda_vs_data"] - v["psnr_oracle_vs_data"]) < 0.05, (name, v)        # north_star, bf16 mode
        if bf:
            assert v["snr_cuda_vs_oracle_db"] > v["snr_torch_autocast_vs_oracle_db"] - 3.0, (name, v)
        else:
            assert v["snr_cuda_vs_oracle_db"] > 80.0, (name, v)                                  # 1e-4 of the image norm
    # 3. every parameter gradient against the measured noise floor.  Measured on B200 (profiles/r2_fullsize_parity_*.json):
    #    bf16: median L2 error 0.149 against the fp32 oracle - stock bf16 autocast 0.146 (cosines 0.9896 / 0.9900);
    #    fp32: median 9.1e-4 / worst 3.4e-3 against the fp64 oracle, 99.8 % of the elements within 1e-3 - the reference's own
    #          fp32 arithmetic: 1.3e-3 / 5.2e-3, 99.0 % (before the two-level accumulation of csrc/conv_simt.cu the CUDA fp32
    #          engine stood at 2.7e-3 / 1.3e-2, 93.6 %).
    for k, v in ours.items():
        y = yard[k]
        assert v["l2"] <= max(5e-2 if bf else 4e-3, (2.5 if bf else 3.0) * y["l2"]), (k, v, y)
    so, sy = _summary(ours), _summary(yard)
    assert so["median_l2"] <= (1.15 if bf else 1.1) * sy["median_l2"] + 1e-4, (so, sy)
    assert so["median_cos"] >= sy["median_cos"] - (5e-3 if bf else 1e-5), (so, sy)
    if bf:
        assert so["frac_elements_within_1e-3_of_tensor_max"] >= sy["frac_elements_within_1e-3_of_tensor_max"] - 0.03, (so, sy)
    else:   # north_star: fp32 gradients within 1e-3 - element by element against the fp64 truth
        assert so["frac_elements_within_1e-3_of_tensor_max"] >= 0.995, so
        assert so["median_l2"] <= 1.2e-3, so
    if not bf:      # discriminator gradients are well conditioned: north_star's 1e-3 holds outright
        for k, v in ours.items():
            if k.startswith("D_"):
                assert v["l2"] < 1e-3, (k, v)
