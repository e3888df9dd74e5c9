#!/usr/bin/env python
"""Where does the end-to-end step lose time against the device-only step at N > 1?  (torchrun, one rank per GPU)
Variants of the launcher pipeline: full, no upload, no download, neither; max over ranks, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from tw_invoice_unet_ocr_llm_b200.launcher import GpuWorker, bind_to_gpu_numa
from tw_invoice_unet_ocr_llm_b200.synthetic import make_fixture_state, synthetic_invoices_u8

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
if world > 1:
    bind_to_gpu_numa(lr)
    dist.init_process_group("nccl", device_id=dev)
B, S, K = 64, 512, 10
w = GpuWorker(make_fixture_state(), dev, chunk=64, packed=True)
frames = torch.from_numpy(np.concatenate([synthetic_invoices_u8(8, S, S, seed=7 + rank)] * 8)).pin_memory()
outs = [torch.empty((B, 3, S, S // 8), dtype=torch.uint8).pin_memory() for _ in range(2)]

def sync():
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier(); torch.cuda.synchronize(dev)

def run(upload, download, k):
    bufs = w._staging(frames.shape, outs[0].shape)
    cs, ks, ds = w.copy_stream, w.compute_stream, w.down_stream
    drained = [None, None]
    for i in range(k):
        slot = i & 1
        xin, mout = bufs[slot]
        ready, done, dr = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        with torch.cuda.stream(cs):
            if drained[slot] is not None: cs.wait_event(drained[slot])
            if upload: xin.copy_(frames, non_blocking=True)
            ready.record(cs)
        with torch.cuda.stream(ks):
            ks.wait_event(ready)
            w.engine.run(xin, want_logits=False, thresholds=w.thresholds, mask_out=mout, mask_bits=True)
            done.record(ks)
        with torch.cuda.stream(ds):
            ds.wait_event(done)
            if download: outs[slot].copy_(mout, non_blocking=True)
            dr.record(ds)
            drained[slot] = dr
    w.join_current_stream()

res = {}
for name, (u, d) in {"full": (1, 1), "no_upload": (0, 1), "no_download": (1, 0), "neither": (0, 0), "full2": (1, 1)}.items():
    run(u, d, 3); w.synchronize(); sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(u, d, K); e1.record(); sync()
    ms = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device=dev)
    mn = ms.clone()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    res[name] = (round(float(ms), 3), round(float(mn), 3))
if rank == 0:
    print("N =", world, "ms per step (max over ranks, min over ranks):", res)
if world > 1:
    dist.destroy_process_group()
