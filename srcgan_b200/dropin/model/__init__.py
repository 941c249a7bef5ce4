"""Drop-in ``model`` package: put ``srcgan_b200/dropin`` in front of the reference's ``src`` on
PYTHONPATH and ``from model import RDDBNetA, RDDBNetB, NLayerDiscriminator`` (train.py:11) resolves
to the B200-native networks.  SRDenseNetA/B (train.py:166-168, opt.net=='SRdens') and the cascaded
generators are not built yet: asking for them raises instead of silently falling back."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _root not in sys.path:
    sys.path.insert(0, _root)

from srcgan_b200.nn import ESPCN, NLayerDiscriminator, RDDBNet, RDDBNetA, RDDBNetB, SRCNN, SRDN  # noqa: E402,F401


def _missing(name):
    class _Missing:
        def __init__(self, *a, **k):
            raise NotImplementedError("srcgan_b200 drop-in: %s is not implemented on the B200 path yet" % name)
    _Missing.__name__ = name
    return _Missing


SRDenseNetA = _missing("SRDenseNetA")
SRDenseNetB = _missing("SRDenseNetB")
for _n in ("EDSR", "ResDeconv"):
    globals()[_n] = _missing(_n)
__all__ = ["RDDBNetA", "RDDBNetB", "NLayerDiscriminator", "SRDenseNetA", "SRDenseNetB",
           "ESPCN", "SRCNN", "EDSR", "RDDBNet", "SRDN", "ResDeconv"]
