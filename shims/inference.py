"""Drop-in replacement for the reference's top-level ``inference.py``: ``from inference import
run_unet`` (reference app_camera.py:16) resolves to the B200 implementation.  See INTEGRATION.md."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from tw_invoice_unet_ocr_llm_b200.inference import (  # noqa: E402,F401
    DEVICE, FIELDS, IMG_SIZE, THRESHOLDS, load_model, preprocess, run_unet, run_unet_batch, run_unet_enhanced)
