"""Checkpoint interchange with the reference: same file names, same ``state_dict`` keys, plain ``torch.save``.

Cascaded trainers write ``checkpoints/<Model>[@G2LAB]_{A2C,C2B}_x<up>_<epoch:04d>.pth`` (trainCas.py:222-225,
trainCasLAB.py:220-223) and the eval scripts rebuild the networks from the file NAME alone
(testCas.py:40-56: ``eval(check[0])(1, 1, int(check[2][1]))`` / ``(1, 3)``; LAB variants strip ``@G2LAB`` and use
2 output channels, testCasLAB.py:63-64).  The CycleGAN trainer writes ``netG_{A2B,B2A}_SRtask_<mode>_<epoch>.pth``
(train.py:407-410).  Weights trained by either side load into the other with ``strict=True``.
"""
from __future__ import annotations

import os
from typing import NamedTuple, Tuple

import torch


class CasName(NamedTuple):
    model: str       # registry name, e.g. "RDDBNet"
    lab: bool        # "@G2LAB" suffix: LAB-space variant
    role: str        # "A2C" (super-resolution) or "C2B" (colourisation)
    up: int
    epoch: int


def cas_checkpoint_name(model: str, role: str, up: int, epoch: int, lab: bool = False) -> str:
    assert role in ("A2C", "C2B")
    return "%s%s_%s_x%d_%04d.pth" % (model, "@G2LAB" if lab else "", role, up, epoch)


def parse_cas_checkpoint_name(path: str) -> CasName:
    """Same split the eval scripts do (testCas.py:40-41); the scale is the single digit after 'x' (:52)."""
    parts = os.path.basename(path).split(".pth")[0].split("_")
    if len(parts) != 4 or parts[1] not in ("A2C", "C2B") or not parts[2].startswith("x"):
        raise ValueError("not a cascaded-trainer checkpoint name: %r" % path)
    return CasName(parts[0].replace("@G2LAB", ""), "@G2LAB" in parts[0], parts[1], int(parts[2][1]), int(parts[3]))


def cyclegan_checkpoint_names(mode: str, epoch: int) -> Tuple[str, str]:
    return ("netG_A2B_SRtask_%s_%04d.pth" % (mode, epoch), "netG_B2A_SRtask_%s_%04d.pth" % (mode, epoch))


def save_state(net: torch.nn.Module, path: str) -> None:
    """``torch.save(net.state_dict(), path)`` with CPU tensors (the reference saves device tensors; its loaders
    call ``torch.load`` without map_location, so CPU tensors load everywhere)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in net.state_dict().items()}, path)


def load_state(net: torch.nn.Module, path: str) -> None:
    net.load_state_dict(torch.load(path, map_location="cpu"), strict=True)
    if hasattr(net, "invalidate_packs"):
        net.invalidate_packs()


def save_cascade(model, ckpt_dir: str, epoch: int) -> Tuple[str, str]:
    """What trainCas*.py:220-225 does at the end of an epoch for a ``trainer_cas.CasSRC``."""
    o = model.opt
    a = os.path.join(ckpt_dir, cas_checkpoint_name(o.SRModel, "A2C", o.up, epoch, model.lab))
    b = os.path.join(ckpt_dir, cas_checkpoint_name(o.CModel, "C2B", o.up, epoch, model.lab))
    save_state(model.netG_A2C, a)
    save_state(model.netG_C2B, b)
    return a, b


def load_cascade_pair(path_a2c: str, path_c2b: str, device):
    """Rebuilds the two generators from the checkpoint file names, like testCas.py:52-58 (eval mode)."""
    from .trainer_cas import build_model
    a, b = parse_cas_checkpoint_name(path_a2c), parse_cas_checkpoint_name(path_c2b)
    if a.role != "A2C" or b.role != "C2B":
        raise ValueError("expected an A2C and a C2B checkpoint, got %s / %s" % (a.role, b.role))
    net_a = build_model(a.model, 1, 1, a.up).to(device)
    net_b = build_model(b.model, 1, 2 if b.lab else 3).to(device)
    load_state(net_a, path_a2c)
    load_state(net_b, path_c2b)
    return net_a.eval(), net_b.eval(), a, b
