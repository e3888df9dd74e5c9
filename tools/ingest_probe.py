#!/usr/bin/env python
"""Device-only forward with fp32 NCHW vs uint8 NHWC ingest (batch 64 @ 512^2): whole step and the first conv."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tw_invoice_unet_ocr_llm_b200.engine import Engine
from tw_invoice_unet_ocr_llm_b200.synthetic import make_fixture_state, synthetic_invoices_u8
dev = torch.device("cuda", 0)
eng = Engine(make_fixture_state(), dev)
u8 = np.concatenate([synthetic_invoices_u8(8, 512, 512, seed=7)] * 8)
xu = torch.from_numpy(u8).to(dev)
xf = torch.from_numpy(u8.astype(np.float32) / 255.0).permute(0, 3, 1, 2).contiguous().to(dev)
mask = torch.empty((64, 3, 512, 512 // 8), dtype=torch.uint8, device=dev)
thr = [0.25, 0.40, 0.30]
for name, x in (("fp32 NCHW", xf), ("uint8 NHWC", xu), ("fp32 NCHW", xf), ("uint8 NHWC", xu)):
    run = lambda: eng.run(x, want_logits=False, thresholds=thr, mask_out=mask, mask_bits=True)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    eng.set_option("profile", 1); run(); run(); t = eng.layer_times_ms(); eng.set_option("profile", 0)
    print(f"{name:11s} step {e0.elapsed_time(e1) / 10:.3f} ms   first conv {t[0]:.3f} ms  down1.net.3 {t[1]:.3f} conv1.net.3 {t[21]:.3f}")
