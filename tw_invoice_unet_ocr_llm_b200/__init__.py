"""B200-native (sm_100a) drop-in for the U-Net forward of tingyu-c/TW-invoice-unet-ocr-llm.

Public surface mirrors the reference modules:

* ``unet_model``  -- ``DoubleConv``, ``UNet``            (reference unet_model.py)
* ``inference``   -- ``load_model``, ``preprocess``, ``run_unet`` (+ ``run_unet_batch``),
  ``DEVICE``, ``IMG_SIZE``, ``FIELDS``                (reference inference.py)
* ``launcher``    -- multi-GPU batch launcher (new)
* ``engine``      -- host driver of ``libunetb200.so`` (C ABI in ``include/unetb200.h``)
"""
__all__ = ["unet_model", "inference", "engine", "launcher", "synthetic"]
