#!/usr/bin/env python
"""cProfile of inference.run_unet_batch on 64 PIL 1080p frames (host-side breakdown); GPU box only."""
import cProfile
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from PIL import Image
from tw_invoice_unet_ocr_llm_b200 import inference as inf
from tw_invoice_unet_ocr_llm_b200.synthetic import make_fixture_state, synthetic_invoices_u8

state = make_fixture_state()
with tempfile.TemporaryDirectory() as d:
    ckpt = os.path.join(d, "best_unet_model.pth")
    torch.save(state, ckpt)
    four = synthetic_invoices_u8(4, 1080, 1920, seed=12)
    pils = [Image.fromarray(four[i % 4]) for i in range(64)]
    for _ in range(2):
        inf.run_unet_batch(pils, ckpt)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        inf.run_unet_batch(pils, ckpt)
        ts.append(time.perf_counter() - t0)
    print("wall ms:", [round(t * 1e3, 1) for t in ts], "cpus", os.cpu_count())
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        inf.run_unet_batch(pils, ckpt, crop_views=True)
        ts.append(time.perf_counter() - t0)
    print("wall ms with crop views:", [round(t * 1e3, 1) for t in ts])
    inf._TRACE = []
    inf.run_unet_batch(pils, ckpt)
    print("trace ms:", [(l, round(t * 1e3, 1)) for l, t in inf._TRACE])
    inf._TRACE = None
    import numpy as np
    a = np.asarray(inf._rgb_host_view(pils[0]))
    dst = torch.empty(a.size, dtype=torch.uint8, pin_memory=True).numpy().reshape(a.shape)
    t0 = time.perf_counter()
    for _ in range(16):
        np.copyto(dst, a)
    print("single-thread staging copy ms per frame:", (time.perf_counter() - t0) / 16 * 1e3, a.shape)
    t0 = time.perf_counter()
    for _ in range(16):
        pils[0].crop((100, 100, 1700, 900))
    print("one large crop ms:", (time.perf_counter() - t0) / 16 * 1e3)
    masks, crops = inf.run_unet(pils[0], ckpt)
    print("crop sizes:", {k: (c.size if c is not None else None) for k, c in crops.items()})
    pr = cProfile.Profile()
    pr.enable()
    inf.run_unet_batch(pils, ckpt)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
