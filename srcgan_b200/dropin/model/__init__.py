"""Drop-in ``model`` package: put ``srcgan_b200/dropin`` in front of the reference's ``src`` on
PYTHONPATH and ``from model import RDDBNetA, RDDBNetB, NLayerDiscriminator`` (train.py:11) resolves
to the B200-native networks; the cascaded trainers' ``from model import *`` (trainCas.py:11) finds every
generator of src/model/__init__.py here as well."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _root not in sys.path:
    sys.path.insert(0, _root)

from srcgan_b200.nn import ESPCN, NLayerDiscriminator, RDDBNet, RDDBNetA, RDDBNetB, SRCNN, SRDN  # noqa: E402,F401
from srcgan_b200.zoo import EDSR, ResDeconv, SRDenseNetA, SRDenseNetB  # noqa: E402,F401

__all__ = ["RDDBNetA", "RDDBNetB", "NLayerDiscriminator", "SRDenseNetA", "SRDenseNetB",
           "ESPCN", "SRCNN", "EDSR", "RDDBNet", "SRDN", "ResDeconv"]
