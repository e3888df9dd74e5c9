#!/usr/bin/env python
"""cProfile of ``inference.run_unet`` on a 1920x1080 RGB frame (model cached): where the batch-1
end-to-end latency of BASELINE.json configs[4] goes.   python tools/profile_run_unet.py"""
import cProfile
import os
import pstats
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from PIL import Image  # noqa: E402

from tw_invoice_unet_ocr_llm_b200 import inference as inf  # noqa: E402
from tw_invoice_unet_ocr_llm_b200.synthetic import make_fixture_state, synthetic_invoices_u8  # noqa: E402


def main():
    state = make_fixture_state()
    with tempfile.TemporaryDirectory() as d:
        ckpt = os.path.join(d, "best_unet_model.pth")
        torch.save(state, ckpt)
        im = Image.fromarray(synthetic_invoices_u8(1, 1080, 1920, seed=11)[0])
        for _ in range(10):
            inf.run_unet(im, ckpt)
        t0 = time.perf_counter()
        for _ in range(30):
            inf.run_unet(im, ckpt)
        print(f"run_unet 1080p: {(time.perf_counter() - t0) / 30 * 1e3:.2f} ms per call")
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(30):
            inf.run_unet(im, ckpt)
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
