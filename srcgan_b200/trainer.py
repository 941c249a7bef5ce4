"""Step driver for the G+D training step, mirroring the caller-side classes of the reference's
``src/train.py`` (ImagePool :20-64, GANLoss :67-128, SRCycleGAN :145-340, params :344-361) so the
same code that drives the reference drives this package.  The reference's own ``train.py`` also
runs unchanged on the drop-in modules (``srcgan_b200/dropin`` on PYTHONPATH); this mirror exists
because the reference tree is not available on the GPU box and because it routes the adversarial
loss through the fused loss kernel as well.

``opt.net == '1'`` (RGB <-> RGB, the benchmarked hot path) and ``'2'`` (gray <-> RGB) use the RDDB generators;
``'SRdens'`` the SRDenseNetA/B pair of train.py:166-170.
"""
from __future__ import annotations

import itertools
import random
from typing import Dict, List

import torch
import torch.nn.functional as F

from . import losses
from .nn import NLayerDiscriminator, RDDBNetA, RDDBNetB
from .zoo import SRDenseNetA, SRDenseNetB


def update_lr(optimizers, opt) -> None:
    """Per-epoch learning-rate update, behaviour of train.py:196-213 / trainCas.py:45-61: a FRESH scheduler is built for
    every optimizer on every call and stepped once.  With the default ``lr_policy='cosine'`` that multiplies the learning
    rate by (1 + cos(pi / num_epochs)) / 2 each epoch; 'step' (StepLR 50, 0.1) leaves it unchanged on a fresh
    scheduler's first step, 'plateau' (factor 0.2, patience 5) never fires on a fresh scheduler.  'linear' raises: the
    reference's branch reads an undefined name (train.py:199) and cannot run."""
    from torch.optim import lr_scheduler
    import warnings
    for optimizer in optimizers:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")     # "lr_scheduler.step() before optimizer.step()" on the first epoch, as upstream
            if opt.lr_policy == "step":
                lr_scheduler.StepLR(optimizer, step_size=50, gamma=0.1).step()
            elif opt.lr_policy == "plateau":
                lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=0.2, threshold=0.01, patience=5).step(opt.matrix)
            elif opt.lr_policy == "cosine":
                lr_scheduler.CosineAnnealingLR(optimizer, T_max=opt.num_epochs, eta_min=0).step()
            else:
                raise NotImplementedError("learning rate policy [%s] is not implemented" % opt.lr_policy)


class ImagePool:
    """Buffer of previously generated images (pool_size 0 disables it); 50% of the queries swap the
    incoming image with a stored one.  Uses python ``random`` exactly as the reference does."""

    def __init__(self, pool_size: int):
        self.pool_size = pool_size
        self.num_imgs = 0
        self.images: List[torch.Tensor] = []

    def query(self, images: torch.Tensor) -> torch.Tensor:
        if self.pool_size == 0:
            return images
        picked = []
        for img in images.detach():
            img = img.unsqueeze(0)
            if self.num_imgs < self.pool_size:
                self.num_imgs += 1
                self.images.append(img)
                picked.append(img)
                continue
            if random.uniform(0, 1) > 0.5:
                slot = random.randint(0, self.pool_size - 1)
                picked.append(self.images[slot].clone())
                self.images[slot] = img
            else:
                picked.append(img)
        return torch.cat(picked, 0)


class GANLoss(torch.nn.Module):
    """lsgan objective: MSE between the patch logits and an all-ones / all-zeros label, computed by the
    fused loss kernel against a scalar label (no expanded label tensor is materialised)."""

    def __init__(self, gan_mode: str = "lsgan", device=None, target_real_label: float = 1.0,
                 target_fake_label: float = 0.0):
        super().__init__()
        if gan_mode != "lsgan":
            raise NotImplementedError("srcgan_b200.trainer.GANLoss: only 'lsgan' (train.py:186) is built")
        self.gan_mode = gan_mode
        self.real_label, self.fake_label = float(target_real_label), float(target_fake_label)

    def __call__(self, prediction: torch.Tensor, target_is_real: bool) -> torch.Tensor:
        return losses.fused_loss(losses.MSE, prediction, self.real_label if target_is_real else self.fake_label)


class params(object):
    """Hyper-parameters, same names and defaults as train.py:344-361."""

    def __init__(self):
        self.device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
        self.lr = 1e-4
        self.beta1 = 0.5
        self.batch_size = 1
        self.num_works = 2
        self.num_epochs = 25
        self.pool_size = 4
        self.lambda_identity = 1.0
        self.lambda_A = 10
        self.lambda_B = 10
        self.n_epochs_decay = 100
        self.matrix = 0
        self.lr_policy = "cosine"
        self.mode = "x2"
        self.net = "1"
        self.scaling_factor = 2


class SRCycleGAN(object):
    """Owns G_A (LR->HR), G_B (HR->LR), D_A, D_B, both Adam optimisers and the step schedule."""

    loss_names = ["D_A", "G_A", "cycle_A", "iden_A", "D_B", "G_B", "cycle_B", "iden_B"]

    def __init__(self, opt):
        self.opt = opt
        dev = opt.device
        gray = opt.net != "1"
        if opt.net == "SRdens":
            self.netG_A = SRDenseNetA(1, 3, mode=opt.mode, num_blocks=2, num_layers=2).to(dev)
            self.netG_B = SRDenseNetB(3, 1, mode=opt.mode, num_blocks=2, num_layers=2).to(dev)
        else:
            self.netG_A = RDDBNetB(1 if gray else 3, 3, 64, nb=3, mode=opt.mode).to(dev)
            self.netG_B = RDDBNetA(3, 1 if gray else 3, 64, nb=3, mode=opt.mode).to(dev)
        self.netD_A = NLayerDiscriminator(3, 64, 2).to(dev)
        self.netD_B = NLayerDiscriminator(1 if gray else 3, 64, 2).to(dev)
        self.fake_A_pool = ImagePool(opt.pool_size)
        self.fake_B_pool = ImagePool(opt.pool_size)
        self.criterionGAN = GANLoss("lsgan", device=dev)
        self.criterionCycle = losses.L1Loss()
        self.criterionIdt = losses.L1Loss()
        self.optimizer_G = torch.optim.Adam(
            itertools.chain(self.netG_A.parameters(), self.netG_B.parameters()), lr=opt.lr, betas=(opt.beta1, 0.999))
        self.optimizer_D = torch.optim.Adam(
            itertools.chain(self.netD_A.parameters(), self.netD_B.parameters()), lr=1e-5, betas=(opt.beta1, 0.999))
        self.optimizers = [self.optimizer_G, self.optimizer_D]

    def update_lr(self, opt) -> None:
        """train.py:196-213; called once per epoch by the reference's loop (train.py:378)."""
        update_lr(self.optimizers, opt)

    @staticmethod
    def set_requires_grad(nets, requires_grad: bool = False) -> None:
        for net in (nets if isinstance(nets, list) else [nets]):
            if net is not None:
                for p in net.parameters():
                    p.requires_grad = requires_grad

    def forward(self, realA: torch.Tensor, realB: torch.Tensor) -> None:
        scale = 2 if self.opt.mode == "x2" else 4
        self.real_A, self.real_B = realA, realB
        self.fake_B = self.netG_A(realA)
        self.recl_A = self.netG_B(self.fake_B)
        self.fake_A = self.netG_B(realB)
        self.recl_B = self.netG_A(self.fake_A)
        if self.opt.net == "1":
            hr_as_lr, lr_as_hr = realB, realA
        else:   # luma of the HR image; LR gray replicated to 3 channels (train.py:252-257)
            hr_as_lr = 0.2125 * realB[:, :1] + 0.7154 * realB[:, 1:2] + 0.0721 * realB[:, 2:3]
            lr_as_hr = torch.cat([realA, realA, realA], dim=1)
        self.real_B_Gray = F.interpolate(hr_as_lr, scale_factor=1.0 / scale)
        self.iden_A = self.netG_A(self.real_B_Gray)
        self.real_A_RGB = F.interpolate(lr_as_hr, scale_factor=scale)
        self.iden_B = self.netG_B(self.real_A_RGB)

    def backward_D_basic(self, netD, real, fake):
        loss_real = self.criterionGAN(netD(real), True)
        loss_fake = self.criterionGAN(netD(fake.detach()), False)
        loss_D = (loss_real + loss_fake) * 0.5
        loss_D.backward()
        return loss_D

    def backward_D_A(self):
        self.loss_D_A = self.backward_D_basic(self.netD_A, self.real_B, self.fake_B_pool.query(self.fake_B))

    def backward_D_B(self):
        self.loss_D_B = self.backward_D_basic(self.netD_B, self.real_A, self.fake_A_pool.query(self.fake_A))

    def backward_G(self):
        o = self.opt
        if o.lambda_identity > 0:
            self.loss_iden_A = self.criterionIdt(self.iden_A, self.real_B) * o.lambda_B / 2 * o.lambda_identity
            self.loss_iden_B = self.criterionIdt(self.iden_B, self.real_A) * o.lambda_A / 2 * o.lambda_identity
        else:
            self.loss_iden_A = 0
            self.loss_iden_B = 0
        self.loss_G_A = self.criterionGAN(self.netD_A(self.fake_B), True)
        self.loss_G_B = self.criterionGAN(self.netD_B(self.fake_A), True)
        self.loss_cycle_A = self.criterionCycle(self.recl_A, self.real_A) * o.lambda_A * 0.5
        self.loss_cycle_B = self.criterionCycle(self.recl_B, self.real_B) * o.lambda_B * 0.5
        self.loss_G = (self.loss_G_A + self.loss_G_B) + self.loss_cycle_A + self.loss_cycle_B \
            + self.loss_iden_A + self.loss_iden_B
        self.loss_G.backward()

    def optimize_parameters(self, realA, realB):
        self.forward(realA, realB)
        self.set_requires_grad([self.netD_A, self.netD_B], False)
        self.optimizer_G.zero_grad()
        self.backward_G()
        self.optimizer_G.step()
        self.set_requires_grad([self.netD_A, self.netD_B], True)
        self.optimizer_D.zero_grad()
        self.backward_D_A()
        self.backward_D_B()
        self.optimizer_D.step()

    def current_losses(self) -> Dict[str, float]:
        """One device->host read of all nine scalars (the reference's logger does 5 ``.item()`` syncs)."""
        names = self.loss_names + ["G"]
        vals = torch.stack([torch.as_tensor(getattr(self, "loss_" + n), dtype=torch.float32,
                                            device=self.real_A.device).detach() for n in names]).tolist()
        return dict(zip(names, vals))
