#!/usr/bin/env python
"""Where the host time of one batch-1 Engine.run goes (GPU box)."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tw_invoice_unet_ocr_llm_b200.engine import Engine
from tw_invoice_unet_ocr_llm_b200.synthetic import make_fixture_state, synthetic_invoices
dev = torch.device("cuda", 0)
eng = Engine(make_fixture_state(), dev)
x = synthetic_invoices(1, 512, 512, seed=1).to(dev)
l1 = torch.empty((1, 3, 512, 512), dtype=torch.float32, device=dev)
m1 = torch.empty((1, 3, 512, 512), dtype=torch.uint8, device=dev)
thr = [0.25, 0.40, 0.30]
def call():
    eng.run(x, want_logits=True, thresholds=thr, logits_out=l1, mask_out=m1)
for _ in range(50): call()
torch.cuda.synchronize()
ts = []
for _ in range(300):
    t0 = time.perf_counter(); call(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append((t1 - t0, t2 - t0))
ts.sort(key=lambda t: t[1])
print("enqueue p50 %.3f ms, total p50 %.3f ms" % (sorted(t[0] for t in ts)[150] * 1e3, ts[150][1] * 1e3))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): call()
e1.record(); torch.cuda.synchronize()
print("back-to-back device ms per forward: %.3f" % (e0.elapsed_time(e1) / 100))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): call()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
