"""Drop-in ``losses`` module (shadows the reference's src/losses.py on PYTHONPATH)."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
from srcgan_b200.losses import DSSIMLoss, L1Loss, MSELoss, PSNRLoss, SSIM  # noqa: E402,F401
