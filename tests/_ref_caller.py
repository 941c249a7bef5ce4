"""Child process of tests/test_gpu_reference_callers.py: imports the reference's UNMODIFIED ``train`` / ``trainCas*``
module from baseline/_ref/src with srcgan_b200/dropin first on sys.path, so that ``from model import ...``, ``import
losses`` and ``import metrics`` inside the reference's scripts resolve to this package, then drives real
``optimize_parameters`` calls on cuda:0.  visdom / skimage (logging and image I/O, not installed) are stubbed.
Never imports ``oracle``.

    python tests/_ref_caller.py <train|trainCas|trainCasConstLAB|...> <in.pt> <out.pt>
"""
import os
import random
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.path.join(ROOT, "baseline", "_ref", "src")


def main():
    which, inp, outp = sys.argv[1:4]
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF_SRC)
    sys.path.insert(0, os.path.join(ROOT, "srcgan_b200", "dropin"))        # wins over the reference's model/losses/metrics
    sys.modules["visdom"] = types.SimpleNamespace(Visdom=lambda *a, **k: None)
    sk = types.ModuleType("skimage")
    sk.io, sk.color = types.ModuleType("skimage.io"), types.ModuleType("skimage.color")
    sk.color.lab2rgb = sk.color.rgb2lab = sk.color.rgb2gray = None
    sk.io.imsave = sk.io.imread = None
    sys.modules.update({"skimage": sk, "skimage.io": sk.io, "skimage.color": sk.color})
    import importlib
    import torch
    mod = importlib.import_module(which)
    assert os.path.samefile(os.path.dirname(mod.__file__), REF_SRC), mod.__file__
    job = torch.load(inp, weights_only=False)
    dev = torch.device("cuda:0")
    from srcgan_b200 import _lib
    before = _lib.launch_count()
    out = {"module_file": mod.__file__, "steps": []}
    if which == "train":
        opt = mod.params()
        opt.device, opt.mode, opt.net = dev, "x4", "1"
        m = mod.SRCycleGAN(opt)
        out["classes"] = {n: type(getattr(m, n)).__module__ + "." + type(getattr(m, n)).__name__
                          for n in ("netG_A", "netG_B", "netD_A", "netD_B", "criterionGAN", "criterionCycle")}
        for name in ("G_A", "G_B", "D_A", "D_B"):
            getattr(m, "net" + name).load_state_dict(job["states"][name], strict=True)
        random.seed(job["seed"])
        for real_A, real_B in job["batches"]:
            m.optimize_parameters(real_A.to(dev), real_B.to(dev))
            rec = {n: float(getattr(m, "loss_" + n)) for n in
                   ("D_A", "G_A", "cycle_A", "iden_A", "D_B", "G_B", "cycle_B", "iden_B", "G")}
            out["steps"].append({"losses": rec, "fake_B": m.fake_B.detach().cpu(), "fake_A": m.fake_A.detach().cpu()})
        m.update_lr(opt)                                  # the epoch loop's call (train.py:378)
        out["lr_after_update"] = [g["lr"] for o in m.optimizers for g in o.param_groups]
    else:
        opt = mod.params()
        opt.device = dev
        opt.up, opt.SRModel, opt.CModel = job["up"], job["SRModel"], job["CModel"]
        m = mod.CasSRC(opt)
        out["classes"] = {n: type(getattr(m, n)).__module__ + "." + type(getattr(m, n)).__name__
                          for n in ("netG_A2C", "netG_C2B", "criterionSR", "criterionPSNR")}
        m.netG_A2C.load_state_dict(job["states"]["A2C"], strict=True)
        m.netG_C2B.load_state_dict(job["states"]["C2B"], strict=True)
        m.init_log()
        for real_A, real_B in job["batches"]:
            m.optimize_parameters(real_A.to(dev), real_B.to(dev))
            out["steps"].append({"loss_SR": float(m.loss_SR), "loss_C": float(m.loss_C), "psnr_SR": float(m.psnr_SR),
                                 "psnr_C": float(m.psnr_C), "fake_AB": m.fake_AB.detach().cpu()})
    torch.cuda.synchronize()
    out["launches"] = _lib.launch_count() - before
    out["last_kernel"] = _lib.last_kernel()
    torch.save(out, outp)
    print("ok", which, out["launches"], "launches;", out["classes"])


if __name__ == "__main__":
    main()
