#!/usr/bin/env python
"""Crop-enhancement measurement: the CUDA batch path against OpenCV on the host.

    python tools/enhance_bench.py            # prints one JSON object

* ``single``: one 48x200 "text" crop through ``enhance.enhance_for_ocrspace`` (PIL in, PIL out,
  copies included) vs the same cv2 calls the reference makes (app_camera.py:581-598), p50 of 200.
* ``batch``: 192 crops (64 invoices x 3 fields, ragged sizes, the three recipes) through one
  ``enhance_batch`` call: end to end (host arrays in, host arrays out) and device only (CUDA events
  around ``unetb200_enhance_run``), vs the cv2 loop on one host core and on all host cores.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _cv2_chain(rgb, kind):
    """The reference's OpenCV calls (app_camera.py:581-598 / :689-703) on an RGB array."""
    import cv2
    gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
    gray = cv2.resize(gray, None, fx=4, fy=4, interpolation=cv2.INTER_CUBIC)
    if kind == "date":
        gray = cv2.createCLAHE(clipLimit=3.0, tileGridSize=(8, 8)).apply(gray)
        gray = cv2.GaussianBlur(gray, (3, 3), 0)
        return cv2.threshold(gray, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)[1]
    gray = cv2.filter2D(gray, -1, np.array([[-1, -1, -1], [-1, 9, -1], [-1, -1, -1]]))
    enhanced = cv2.createCLAHE(clipLimit=4.0, tileGridSize=(8, 8)).apply(gray)
    if kind == "text":
        return cv2.threshold(enhanced, 0, 255, cv2.THRESH_OTSU)[1]
    return enhanced


def _p50(fn, n, warm):
    ts = []
    for i in range(n + warm):
        t0 = time.perf_counter()
        fn()
        if i >= warm:
            ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[int(len(ts) * 0.95)]


def algorithmic_bytes(h, w, flags):
    """src read + upscaled image written, read by the LUT pass and by the CLAHE pass + result written
    (+ read and rewritten by the threshold pass)."""
    b = 3 * h * w + 16 * h * w * 4 + 64 * 256
    if flags & 4:
        b += 2 * 16 * h * w
    return b


def measure(quick: bool = False) -> dict:
    import ctypes as C
    import torch
    from PIL import Image
    from tw_invoice_unet_ocr_llm_b200 import _native as nat, enhance
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_crops_u8
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        cv2 = None
    dev = torch.device("cuda", torch.cuda.current_device())
    res = {}

    # ---- single crop, reference-facing call
    rgb = synthetic_crops_u8([(48, 200)], seed=61)[0]
    pil = Image.fromarray(rgb)
    p50, p95 = _p50(lambda: enhance.enhance_for_ocrspace(pil, mode="text"), 50 if quick else 200, 10)
    res["single"] = {"crop": "48x200 text", "gpu_p50_ms": p50, "gpu_p95_ms": p95}
    if cv2 is not None:
        c50, c95 = _p50(lambda: Image.fromarray(_cv2_chain(np.array(pil.convert("RGB")), "text")), 50 if quick else 200, 5)
        res["single"].update({"cv2_p50_ms": c50, "cv2_p95_ms": c95})

    # ---- ragged batch: 64 invoices x 3 fields
    rng = np.random.default_rng(62)
    sizes = [(int(rng.integers(24, 72)), int(rng.integers(90, 320))) for _ in range(192)]
    crops = synthetic_crops_u8(sizes, seed=63)
    kinds = [("text", "date", "amount")[i % 3] for i in range(len(crops))]
    e50, _ = _p50(lambda: enhance.enhance_batch(crops, kinds), 10 if quick else 30, 3)
    table, sb, ob, wb = enhance.plan(sizes, kinds)
    tab_bytes = (C.sizeof(table) + 15) & ~15
    host = np.zeros(tab_bytes + sb, np.uint8)
    host[:C.sizeof(table)] = np.frombuffer(table, dtype=np.uint8)
    for a, t in zip(crops, table):
        host[tab_bytes + t.src_off: tab_bytes + t.src_off + a.size] = a.reshape(-1)
    d_in = torch.from_numpy(host).to(dev)
    d_out = torch.empty(ob, dtype=torch.uint8, device=dev)
    d_ws = torch.empty(wb, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def run():
        nat.check(nat.lib().unetb200_enhance_run(table, d_in.data_ptr(), len(crops), d_in.data_ptr() + tab_bytes,
                                                 d_out.data_ptr(), d_ws.data_ptr(), stream.cuda_stream))
    for _ in range(5):
        run()
    torch.cuda.synchronize(dev)
    reps = 20 if quick else 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        run()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    dev_ms = e0.elapsed_time(e1) / reps
    alg = sum(algorithmic_bytes(h, w, enhance.KINDS[k][0]) for (h, w), k in zip(sizes, kinds))
    res["batch"] = {
        "crops": len(crops), "output_megapixels": sum(16 * h * w for h, w in sizes) / 1e6,
        "e2e_ms": e50, "e2e_crops_per_s": len(crops) / e50 * 1e3,
        "device_ms": dev_ms, "device_crops_per_s": len(crops) / dev_ms * 1e3,
        "algorithmic_bytes": alg, "device_gb_per_s": alg / dev_ms / 1e6,
        "launches": 5, "note": "5 launches per batch; working set fits the L2, so GB/s is not an HBM figure",
    }
    if cv2 is not None:
        t0 = time.perf_counter()
        reps_c = 1 if quick else 3
        for _ in range(reps_c):
            for a, k in zip(crops, kinds):
                _cv2_chain(a, k)
        c_ms = (time.perf_counter() - t0) * 1e3 / reps_c
        res["batch"].update({"cv2_1core_ms": c_ms, "cv2_1core_crops_per_s": len(crops) / c_ms * 1e3})
        from concurrent.futures import ThreadPoolExecutor
        cores = os.cpu_count() or 1
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(lambda ak: _cv2_chain(*ak), zip(crops, kinds)))
            t0 = time.perf_counter()
            for _ in range(reps_c):
                list(ex.map(lambda ak: _cv2_chain(*ak), zip(crops, kinds)))
            m_ms = (time.perf_counter() - t0) * 1e3 / reps_c
        res["batch"].update({"cv2_allcores_ms": m_ms, "cv2_allcores_crops_per_s": len(crops) / m_ms * 1e3,
                             "host_cores": cores})
    return res


if __name__ == "__main__":
    print(json.dumps(measure(quick="--quick" in sys.argv)))
