"""Step driver of the cascaded SR + colourisation trainers, mirroring the caller-side class ``CasSRC`` of the
reference's ``src/trainCas.py:18-153`` and its three variants (``trainCasConst.py:89-92`` blurs by
down-then-up bilinear resizing so the SR net works at full resolution; ``trainCasLAB.py:83-84`` /
``trainCasConstLAB.py`` train on LAB tensors: SR target = L, colouriser target = ab, 2 output channels).

Two generators, each trained with a plain L1 loss by its own Adam; an eval-mode "transfer" pass and a PSNR
monitor complete the iteration.  The reference's unmodified scripts also run on the drop-in modules; this
mirror exists because the reference tree is not on the GPU box.  Differences kept deliberately small:
models are looked up in a registry instead of ``eval(opt.SRModel)``, the four ``.item()`` host syncs per
iteration are replaced by one batched read in ``log_values()``, and the transfer pass runs under no_grad
(the reference builds and discards an autograd graph there)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import losses
from . import nn as snn
from . import zoo

MODELS = {"RDDBNet": snn.RDDBNet, "SRDN": snn.SRDN, "ESPCN": snn.ESPCN, "SRCNN": snn.SRCNN,
          "EDSR": zoo.EDSR, "ResDeconv": zoo.ResDeconv}


def build_model(name: str, *args):
    if name not in MODELS:
        raise NotImplementedError("srcgan_b200: model %r is not built on the B200 path (have: %s)"
                                  % (name, ", ".join(sorted(MODELS))))
    return MODELS[name](*args)


class params(object):
    """Same names / defaults as trainCas.py:155-164 plus the argparse values (:168-176)."""

    def __init__(self):
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.lr = 1e-4
        self.batch_size = 1
        self.num_works = 2
        self.num_epochs = 50
        self.matrix = 0
        self.lr_policy = "cosine"
        self.SRModel = "ESPCN"
        self.CModel = "ResDeconv"
        self.up = 2
        self.variant = ""              # "", "Const", "LAB", "ConstLAB"


class CasSRC(object):
    def __init__(self, opt):
        self.opt = opt
        self.lab = "LAB" in opt.variant
        self.const = "Const" in opt.variant
        self.netG_A2C = build_model(opt.SRModel, 1, 1, opt.up).to(opt.device)
        self.netG_C2B = build_model(opt.CModel, 1, 2 if self.lab else 3).to(opt.device)
        self.criterionSR = losses.L1Loss()
        self.criterionC = losses.L1Loss()
        self.criterionPSNR = losses.PSNRLoss()
        self.optimizer_G = torch.optim.Adam(self.netG_A2C.parameters(), lr=opt.lr)
        self.optimizer_D = torch.optim.Adam(self.netG_C2B.parameters(), lr=opt.lr)     # "D" optimises the colouriser
        self.optimizers = [self.optimizer_G, self.optimizer_D]
        self.init_log()

    def update_lr(self, opt) -> None:
        """trainCas.py:45-61; called once per epoch by the reference's loop (trainCas.py:189)."""
        from .trainer import update_lr
        update_lr(self.optimizers, opt)

    def init_log(self):
        self._log = {"loss_sr": [], "loss_c": [], "psnr_sr": [], "psnr_c": []}

    def _blur(self, x: torch.Tensor) -> torch.Tensor:
        y = F.interpolate(x, scale_factor=1.0 / self.opt.up, mode="bilinear")
        if self.const:
            y = F.interpolate(y, scale_factor=self.opt.up, mode="bilinear")
        return y

    def forwardSR(self, realB: torch.Tensor) -> None:
        if self.lab:
            self.real_B, self.real_BC = realB[:, 1:], realB[:, :1]                      # ab / L
        else:
            self.real_B = realB
            self.real_BC = 0.2125 * realB[:, :1] + 0.7154 * realB[:, 1:2] + 0.0721 * realB[:, 2:3]
        self.real_BA = self._blur(self.real_BC)
        self.fake_BC = self.netG_A2C(self.real_BA)

    def forwardC(self) -> None:
        self.fake_BB = self.netG_C2B(self.real_BC)

    @torch.no_grad()
    def transfer(self, realA: torch.Tensor) -> None:
        # Const variants feed real_A at full resolution (trainCasConst.py:103-106)
        self.real_A = realA if self.const else self._blur(realA)
        self.netG_A2C.eval()
        self.netG_C2B.eval()
        self.fake_AC = self.netG_A2C(self.real_A)
        self.fake_AB = self.netG_C2B(self.fake_AC)

    def backward_G(self) -> None:
        self.loss_SR = self.criterionSR(self.fake_BC, self.real_BC)
        self.loss_SR.backward()
        self._log["loss_sr"].append(self.loss_SR.detach())

    def backward_D(self) -> None:
        self.loss_C = self.criterionC(self.fake_BB, self.real_B)
        self.loss_C.backward()
        self._log["loss_c"].append(self.loss_C.detach())

    def validate(self) -> None:
        self.psnr_SR = self.criterionPSNR(self.fake_BC.detach(), self.real_BC.detach())
        self.psnr_C = self.criterionPSNR(self.fake_BB.detach(), self.real_B.detach())
        self._log["psnr_sr"].append(self.psnr_SR)
        self._log["psnr_c"].append(self.psnr_C)

    def optimize_parameters(self, realA: torch.Tensor, realB: torch.Tensor) -> None:
        self.netG_A2C.train()
        self.netG_C2B.train()
        self.forwardSR(realB)
        self.optimizer_G.zero_grad()
        self.backward_G()
        self.optimizer_G.step()
        self.forwardC()
        self.optimizer_D.zero_grad()
        self.backward_D()
        self.optimizer_D.step()
        self.transfer(realA)
        self.validate()

    def log_values(self) -> Dict[str, float]:
        """Means of the logged scalars since init_log(), with ONE device->host read."""
        keys = [k for k, v in self._log.items() if v]
        if not keys:
            return {}
        vals = torch.stack([torch.stack(self._log[k]).float().mean() for k in keys]).tolist()
        return dict(zip(keys, vals))
