"""Drop-in replacement for the reference's top-level ``unet_model.py``: copy (or put this directory
first on ``sys.path``) and ``from unet_model import UNet`` (reference inference.py:4, train.py:12)
resolves to the B200 implementation.  See INTEGRATION.md."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from tw_invoice_unet_ocr_llm_b200.unet_model import DoubleConv, UNet  # noqa: E402,F401
