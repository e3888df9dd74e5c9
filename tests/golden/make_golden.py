#!/usr/bin/env python
"""Generate the golden vectors that pin the oracle -- run in the BUILD container only.

The reference ships no tests or fixtures, so the pins are outputs of the reference itself:
this script imports the UNMODIFIED reference modules from /root/reference (unet_model.UNet,
inference.run_unet), loads the deterministic fixture checkpoint into them and records

  golden_logits.npz   UNet.forward logits for seeded synthetic inputs (2x3x64x64, 1x3x32x48)
  golden_run_unet.npz inference.run_unet on a synthetic 1280x720 frame: the three boolean
                      masks (bit-packed) and the crop rectangles / sizes it produced

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import inference as ref_inference          # noqa: E402  (the reference, unmodified)
import unet_model as ref_unet_model         # noqa: E402

from tw_invoice_unet_ocr_llm_b200.synthetic import (make_fixture_state, synthetic_invoices,  # noqa: E402
                                                    synthetic_invoices_u8)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    state = make_fixture_state()
    model = ref_unet_model.UNet(3, 3)
    model.load_state_dict(state)
    model.eval()
    out = {}
    for tag, (n, h, w, seed) in {"a": (2, 64, 64, 42), "b": (1, 32, 48, 43)}.items():
        x = synthetic_invoices(n, h, w, seed=seed)
        with torch.no_grad():
            z = model(x)
        out[f"x_{tag}"] = x.numpy()
        out[f"z_{tag}"] = z.numpy()
    # a checksum of the fixture so a drifting generator is caught before the logits are
    out["state_checksum"] = np.array([float(sum(v.double().sum() for v in state.values() if v.dtype.is_floating_point))])
    np.savez_compressed(os.path.join(HERE, "golden_logits.npz"), **out)

    # ---- inference.run_unet end to end (reference inference.py:50-129) on CPU
    frame = synthetic_invoices_u8(1, 720, 1280, seed=77)[0]
    pil = Image.fromarray(frame)
    with tempfile.TemporaryDirectory() as d:
        ckpt = os.path.join(d, "best_unet_model.pth")
        torch.save(state, ckpt)
        ref_inference.DEVICE = "cpu"
        masks, crops = ref_inference.run_unet(pil, ckpt)
    rec = {"frame_seed": np.array([77]), "frame_hw": np.array([720, 1280])}
    for k in ref_inference.FIELDS:
        rec[f"mask_{k}"] = np.packbits(masks[k])
        c = crops[k]
        rec[f"crop_size_{k}"] = np.array(c.size if c is not None else (-1, -1))
    # ---- the mask -> crop stage (reference inference.py:84-127) on structured masks: the
    # reference's own run_unet, with load_model stubbed to a module that emits prescribed logits
    rects = {"invoice_no": (40, 60, 200, 90), "date": (300, 400, 470, 430), "total_amount": None}   # x1,y1,x2,y2 in 512-space

    class Stub(torch.nn.Module):
        def forward(self, x):
            z = torch.full((1, 3, 512, 512), -6.0)
            for c, k in enumerate(ref_inference.FIELDS):
                if rects[k] is not None:
                    x1, y1, x2, y2 = rects[k]
                    z[0, c, y1:y2, x1:x2] = 6.0
            return z

    frame2 = synthetic_invoices_u8(1, 750, 1000, seed=78)[0]
    pil2 = Image.fromarray(frame2)
    keep = ref_inference.load_model
    ref_inference.load_model = lambda path: Stub().eval()
    try:
        masks2, crops2 = ref_inference.run_unet(pil2, "unused")
    finally:
        ref_inference.load_model = keep
    rec["crop_frame_seed"] = np.array([78])
    rec["crop_frame_hw"] = np.array([750, 1000])
    for k in ref_inference.FIELDS:
        rec[f"crop_rect_{k}"] = np.array(rects[k] if rects[k] is not None else (-1, -1, -1, -1))
        c = crops2[k]
        rec[f"crop2_size_{k}"] = np.array(c.size if c is not None else (-1, -1))
        rec[f"crop2_sum_{k}"] = np.array([int(np.asarray(c, dtype=np.int64).sum()) if c is not None else -1])
    np.savez_compressed(os.path.join(HERE, "golden_run_unet.npz"), **rec)
    print({k: (crops2[k].size if crops2[k] else None) for k in crops2})
    print("wrote golden_logits.npz, golden_run_unet.npz;",
          {k: int(masks[k].sum()) for k in masks}, {k: (crops[k].size if crops[k] else None) for k in crops})


if __name__ == "__main__":
    main()
