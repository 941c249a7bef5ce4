"""Import harness for the REAL reference (huster-wgm/SRCGAN under /root/reference).  TEST
INFRASTRUCTURE ONLY - works only where /root/reference exists (the build container).

The reference's ``train.py`` cannot be imported as shipped:
  * ``visdom`` / ``skimage`` are not installed          -> stub modules in ``sys.modules``
  * ``from model import RDDBNetA, RDDBNetB, ...``       -> names injected into package ``model``
  * ``RDDBNetA`` is defined nowhere                     -> shim class below, built ONLY from
    reference classes: ``model.model.RDDBNet`` ctor + ``model.model.Decoder``, forward =
    the commented-out one at ``model.py:370-378``.
Nothing here is copied from the reference; it is imported from where it lies.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("SRCGAN_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "train.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs() -> None:
    """visdom / skimage are only touched by logging, image saving and the dataset."""
    if "visdom" not in sys.modules:
        _stub("visdom", Visdom=lambda *a, **k: None)
    try:
        import skimage  # noqa: F401
    except Exception:
        def _na(*a, **k):
            raise RuntimeError("skimage is not installed (stubbed by oracle/ref_harness.py)")
        sk = _stub("skimage")
        sk.io = _stub("skimage.io", imread=_na, imsave=_na)
        sk.color = _stub("skimage.color", lab2rgb=_na, rgb2lab=_na, rgb2gray=_na)


def import_reference(model_pkg_path: str | None = None):
    """Returns (model_pkg, model_model, losses, metrics) of the reference.

    ``model_pkg_path``: directory to put in front of ``sys.path`` instead of the reference's
    ``src`` for the ``model`` / ``losses`` / ``metrics`` imports (used to prove the drop-in)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    install_stubs()
    for p in (REF_SRC,):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib
    model_pkg = importlib.import_module("model")
    model_model = importlib.import_module("model.model")
    losses = importlib.import_module("losses")
    metrics = importlib.import_module("metrics")
    return model_pkg, model_model, losses, metrics


def make_rddbneta_shim(model_model):
    """The documented stand-in for the missing class (x4 only)."""
    import torch.nn as nn  # noqa: F401

    class RDDBNetA(model_model.RDDBNet):
        def __init__(self, in_nc, out_nc, nf, nb, gc=32, mode="x2"):
            super().__init__(in_nc, out_nc, nf, nb, gc=gc, mode=mode)
            self.decode = model_model.Decoder()

        def forward(self, x):
            fea = self.conv_first(x)
            fea = fea + self.trunk_conv(self.RRDB_trunk(fea))
            return self.conv_last(self.decode(fea))

    return RDDBNetA


def import_train():
    """Imports the reference's ``train`` module (SRCycleGAN, GANLoss, ImagePool, params)."""
    model_pkg, model_model, _losses, _metrics = import_reference()
    if not hasattr(model_pkg, "RDDBNetA"):
        model_pkg.RDDBNetA = make_rddbneta_shim(model_model)
    for name in ("RDDBNetB", "NLayerDiscriminator", "SRDenseNetA", "SRDenseNetB"):
        if not hasattr(model_pkg, name):
            setattr(model_pkg, name, getattr(model_model, name))
    import importlib
    return importlib.import_module("train")


def import_traincas(variant: str = ""):
    """Imports the reference's ``trainCas{,Const,LAB,ConstLAB}`` module (CasSRC, params)."""
    import_reference()
    import importlib
    return importlib.import_module("trainCas" + variant)
