"""Multi-GPU batch launcher (new subsystem; the reference is single-device, batch 1).

Images are independent units, so the batch is split into contiguous shards, one per GPU,
each GPU holds a full packed replica of the weights, and there is NO collective on the data
path: the host scatters uint8 frames (0.79 MB / image at 512x512) and gathers uint8 masks
(0.79 MB / image).  Results are bit-identical to a single-GPU run of the same images.

Two ways to drive it:

* in one process -- ``MultiGpuSegmenter``: one worker thread per GPU, each with its own
  engine, streams and pinned staging, chunks of <= 64 images, copies overlapped with compute;
* one process per GPU (``torchrun``) -- ``shard_bounds`` + ``gather_masks`` (a host-side
  ``torch.distributed.gather`` of the uint8 masks; ``bench.py`` uses this shape).
"""
from __future__ import annotations

import threading
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

MAX_CHUNK = 64
DEFAULT_THRESHOLDS = (0.25, 0.40, 0.30)     # reference inference.py:76-78


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of ``total`` items owned by ``rank`` of ``world`` (sizes differ by <= 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def chunk_bounds(lo: int, hi: int, chunk: int = MAX_CHUNK) -> List[Tuple[int, int]]:
    """Split [lo, hi) into consecutive pieces of at most ``chunk`` items."""
    if chunk <= 0:
        raise ValueError("chunk must be positive")
    return [(a, min(a + chunk, hi)) for a in range(lo, hi, chunk)]


class GpuWorker:
    """One GPU: engine + two streams + double-buffered device staging."""

    def __init__(self, state, device, thresholds: Sequence[float] = DEFAULT_THRESHOLDS,
                 chunk: int = MAX_CHUNK):
        from .engine import Engine
        self.device = torch.device(device)
        self.thresholds = list(thresholds)
        self.chunk = chunk
        self.engine = Engine(state, self.device)
        with torch.cuda.device(self.device):
            self.copy_stream = torch.cuda.Stream(self.device)
            self.compute_stream = torch.cuda.Stream(self.device)
        self._bufs = {}

    def _staging(self, shape_in, shape_out):
        key = (tuple(shape_in[1:]), tuple(shape_out[1:]))
        if key not in self._bufs:
            mk = lambda s: torch.empty((self.chunk, *s[1:]), dtype=torch.uint8, device=self.device)
            self._bufs[key] = [(mk(shape_in), mk(shape_out)) for _ in range(2)]
        return self._bufs[key]

    def segment(self, frames: torch.Tensor, out: torch.Tensor) -> None:
        """``frames`` uint8 [B,H,W,3] (pinned host) -> ``out`` uint8 [B,3,H,W] (pinned host).

        Per chunk: H2D on the copy stream, forward on the compute stream, D2H on the copy
        stream; chunk i+1's upload overlaps chunk i's forward."""
        b, h, w, _ = frames.shape
        bufs = self._staging(frames.shape, out.shape)
        cs, ks = self.copy_stream, self.compute_stream
        with torch.cuda.device(self.device):
            ready = [torch.cuda.Event() for _ in range(2)]     # upload of slot done
            done = [torch.cuda.Event() for _ in range(2)]      # forward of slot done
            drained = [None, None]                             # download of slot done
            for i, (lo, hi) in enumerate(chunk_bounds(0, b, self.chunk)):
                slot = i & 1
                xin, mout = bufs[slot]
                n = hi - lo
                with torch.cuda.stream(cs):
                    if drained[slot] is not None:
                        cs.wait_event(drained[slot])
                    xin[:n].copy_(frames[lo:hi], non_blocking=True)
                    ready[slot].record(cs)
                with torch.cuda.stream(ks):
                    ks.wait_event(ready[slot])
                    self.engine.run(xin[:n], want_logits=False, thresholds=self.thresholds,
                                    mask_out=mout[:n])
                    done[slot].record(ks)
                with torch.cuda.stream(cs):
                    cs.wait_event(done[slot])
                    out[lo:hi].copy_(mout[:n], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                    drained[slot] = ev
            cs.synchronize()
            ks.synchronize()


class MultiGpuSegmenter:
    """Shard a batch of frames over several GPUs of one host (one worker thread per GPU)."""

    def __init__(self, state, devices: Optional[Sequence] = None,
                 thresholds: Sequence[float] = DEFAULT_THRESHOLDS, chunk: int = MAX_CHUNK,
                 worker_factory: Optional[Callable] = None):
        if devices is None:
            devices = [f"cuda:{i}" for i in range(torch.cuda.device_count())]
        if len(devices) == 0:
            raise RuntimeError("MultiGpuSegmenter needs at least one CUDA device; there is no CPU path")
        factory = worker_factory or (lambda dev: GpuWorker(state, dev, thresholds, chunk))
        self.workers = [factory(d) for d in devices]

    def segment(self, frames, out=None):
        """uint8 frames [B,H,W,3] -> uint8 masks [B,3,H,W] (host).  Order is preserved."""
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(np.ascontiguousarray(frames))
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
            raise ValueError(f"expected uint8 [B,H,W,3] frames, got {frames.dtype} {tuple(frames.shape)}")
        b, h, w, _ = frames.shape
        pin = torch.cuda.is_available()
        if pin and not frames.is_pinned():
            frames = frames.pin_memory()
        if out is None:
            out = torch.empty((b, 3, h, w), dtype=torch.uint8, pin_memory=pin)
        errors: List[BaseException] = []

        def work(rank: int):
            lo, hi = shard_bounds(b, len(self.workers), rank)
            if hi > lo:
                try:
                    self.workers[rank].segment(frames[lo:hi], out[lo:hi])
                except BaseException as e:      # surfaced to the caller below
                    errors.append(e)

        threads = [threading.Thread(target=work, args=(r,)) for r in range(len(self.workers))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return out


def gather_masks(local: torch.Tensor, total: int, group=None) -> Optional[torch.Tensor]:
    """Host-side gather for the one-process-per-GPU layout: every rank passes the CPU uint8
    masks of its ``shard_bounds`` range; rank 0 gets the full ``[total, ...]`` tensor in input
    order, other ranks get ``None``.  No GPU collective is involved."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    local = local.cpu().contiguous()
    lo, hi = shard_bounds(total, world, rank)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} items, expected {hi - lo}")
    # gloo gathers equal-sized tensors: pad every shard to the largest one
    biggest = max(b - a for a, b in (shard_bounds(total, world, r) for r in range(world)))
    padded = torch.zeros((biggest, *local.shape[1:]), dtype=local.dtype)
    padded[: hi - lo] = local
    if dist.get_backend(group) == "nccl":
        objs = [None] * world if rank == 0 else None
        dist.gather_object(padded, objs, dst=0, group=group)
        parts = objs
    else:
        parts = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
        dist.gather(padded, parts, dst=0, group=group)
    if rank != 0:
        return None
    out = torch.empty((total, *local.shape[1:]), dtype=local.dtype)
    for r in range(world):
        a, b = shard_bounds(total, world, r)
        out[a:b] = parts[r][: b - a]
    return out
