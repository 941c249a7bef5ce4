"""Drop-in ``metrics`` module (shadows the reference's src/metrics.py on PYTHONPATH)."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
from srcgan_b200.metrics import AE, MSE, PSNR, SSIM  # noqa: E402,F401
