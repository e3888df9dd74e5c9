#!/usr/bin/env python
"""A/B timing of engine options inside ONE process (same box, same thermal state).

usage: tools/ab.py key=v0,v1[,v2...] [--batch 64] [--size 512] [--steps 5] [--rounds 6]
Alternates the settings round-robin, `rounds` times, `steps` forwards each, CUDA-event timed;
prints median / min ms per step for every setting.
"""
import argparse
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from tw_invoice_unet_ocr_llm_b200.engine import Engine
from tw_invoice_unet_ocr_llm_b200.synthetic import make_fixture_state, synthetic_invoices_u8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("spec")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--rounds", type=int, default=6)
    ap.add_argument("--layers", action="store_true", help="also print per-layer times (profiled pass)")
    a = ap.parse_args()
    key, vals = a.spec.split("=")
    vals = [int(v) for v in vals.split(",")]
    dev = torch.device("cuda", 0)
    eng = Engine(make_fixture_state(), dev)
    B, S = a.batch, a.size
    base = synthetic_invoices_u8(min(8, B), S, S, seed=7)
    u8 = np.concatenate([base] * ((B + len(base) - 1) // len(base)))[:B]
    x = torch.from_numpy(u8.astype(np.float32) / 255.0).permute(0, 3, 1, 2).contiguous().to(dev)
    logits = torch.empty((B, 3, S, S), dtype=torch.float32, device=dev)
    mask = torch.empty((B, 3, S, S), dtype=torch.uint8, device=dev)
    thr = [0.25, 0.40, 0.30]
    res = {v: [] for v in vals}
    for rnd in range(a.rounds + 1):
        for v in vals:
            eng.set_option(key, v)
            for _ in range(2):
                eng.run(x, thresholds=thr, logits_out=logits, mask_out=mask)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                eng.run(x, thresholds=thr, logits_out=logits, mask_out=mask)
            e1.record()
            torch.cuda.synchronize()
            if rnd > 0:
                res[v].append(e0.elapsed_time(e1) / a.steps)
    if a.layers:
        per = {}
        for v in vals:
            eng.set_option(key, v)
            eng.set_option("profile", 1)
            acc = None
            for _ in range(4):
                eng.run(x, thresholds=thr, logits_out=logits, mask_out=mask)
                t = eng.layer_times_ms()
                acc = t if acc is None else [p + q for p, q in zip(acc, t)]
            eng.set_option("profile", 0)
            per[v] = [q / 4 for q in acc]
        print("layer".ljust(20) + "".join(f"{key}={v}".rjust(12) for v in vals))
        for i, l in enumerate(eng.layers[:-1]):
            print(l.name.decode().ljust(20) + "".join(f"{per[v][i]:12.3f}" for v in vals))
        print("sum".ljust(20) + "".join(f"{sum(per[v]):12.3f}" for v in vals))
    for v in vals:
        t = res[v]
        print(f"{key}={v}: median {statistics.median(t):.3f} ms  min {min(t):.3f}  max {max(t):.3f}  "
              f"-> {B / statistics.median(t) * 1e3:.1f} img/s   {[round(q, 2) for q in t]}")


if __name__ == "__main__":
    main()
