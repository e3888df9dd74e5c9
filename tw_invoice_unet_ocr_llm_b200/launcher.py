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


def bind_to_gpu_numa(device_index: int) -> Optional[List[int]]:
    """Pin the calling thread (and what it first-touches: its pinned staging memory) to the CPU cores NVML reports
    as local to GPU ``device_index``.  On a multi-socket host the uint8 frames and masks of every GPU then cross
    the socket interconnect zero times instead of once per copy.  Returns the core list, or ``None`` when NVML /
    affinity control is unavailable or the set is empty (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            if visible:                      # map the process-local index to the physical one
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                ident = ids[device_index]
                handle = (pynvml.nvmlDeviceGetHandleByUUID(ident) if ident.startswith(("GPU-", "MIG-"))
                          else pynvml.nvmlDeviceGetHandleByIndex(int(ident)))
            else:
                handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        finally:
            pynvml.nvmlShutdown()
        cores = [64 * i + b for i, wd in enumerate(mask) for b in range(64) if (int(wd) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cores = [c for c in cores if c in allowed]
        if not cores or len(cores) == len(allowed):
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:
        return None


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
    """One GPU: engine + three streams (upload, compute, download) + double-buffered device staging."""

    def __init__(self, state, device, thresholds: Sequence[float] = DEFAULT_THRESHOLDS,
                 chunk: int = MAX_CHUNK):
        from .engine import Engine
        self.device = torch.device(device)
        self.thresholds = list(thresholds)
        self.chunk = chunk
        self.engine = Engine(state, self.device)
        with torch.cuda.device(self.device):
            self.copy_stream = torch.cuda.Stream(self.device)       # host -> device
            self.compute_stream = torch.cuda.Stream(self.device)
            self.down_stream = torch.cuda.Stream(self.device)       # device -> host
        self._bufs = {}
        self._next_slot = 0
        self._drained = [None, None]     # per staging slot: event of its last device->host copy

    def _staging(self, shape_in, shape_out):
        key = (tuple(shape_in[1:]), tuple(shape_out[1:]))
        if key not in self._bufs:
            mk = lambda s: torch.empty((self.chunk, *s[1:]), dtype=torch.uint8, device=self.device)
            self._bufs[key] = [(mk(shape_in), mk(shape_out)) for _ in range(2)]
        return self._bufs[key]

    def segment_async(self, frames: torch.Tensor, out: torch.Tensor, boxes_out: Optional[torch.Tensor] = None
                      ) -> None:
        """Enqueue ``frames`` uint8 [B,H,W,3] (pinned host) -> ``out`` uint8 [B,3,H,W] (pinned host)
        and return without waiting.  Per chunk: H2D on the upload stream, forward on the compute
        stream, D2H on the download stream; the upload of the next chunk (or of the next call) and
        the download of the previous one overlap the forward of the current one.  ``out`` is valid
        after :meth:`synchronize`.  ``boxes_out`` (optional, int32 [B,3,5] pinned host) also receives the
        mask extents {xmin, xmax, ymin, ymax, count} of reference inference.py:85-93, reduced on the GPU."""
        b = frames.shape[0]
        bufs = self._staging(frames.shape, out.shape)
        cs, ks, ds = self.copy_stream, self.compute_stream, self.down_stream
        with torch.cuda.device(self.device):
            for lo, hi in chunk_bounds(0, b, self.chunk):
                slot = self._next_slot
                self._next_slot ^= 1
                xin, mout = bufs[slot]
                n = hi - lo
                ready, done, drained = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
                with torch.cuda.stream(cs):
                    if self._drained[slot] is not None:        # slot's previous masks have left the device
                        cs.wait_event(self._drained[slot])
                    xin[:n].copy_(frames[lo:hi], non_blocking=True)
                    ready.record(cs)
                with torch.cuda.stream(ks):
                    ks.wait_event(ready)
                    self.engine.run(xin[:n], want_logits=False, thresholds=self.thresholds,
                                    mask_out=mout[:n])
                    box = None
                    if boxes_out is not None:
                        from . import prepost
                        box = prepost.mask_bbox(mout[:n])
                        box.record_stream(ds)
                    done.record(ks)
                with torch.cuda.stream(ds):
                    ds.wait_event(done)
                    out[lo:hi].copy_(mout[:n], non_blocking=True)
                    if box is not None:
                        boxes_out[lo:hi].copy_(box, non_blocking=True)
                    drained.record(ds)
                    self._drained[slot] = drained

    def synchronize(self) -> None:
        """Block until everything enqueued by :meth:`segment_async` has landed in host memory."""
        self.copy_stream.synchronize()
        self.compute_stream.synchronize()
        self.down_stream.synchronize()

    def join_current_stream(self) -> None:
        """Make the caller's current stream wait for all enqueued work (for CUDA-event timing)."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.copy_stream)
        cur.wait_stream(self.compute_stream)
        cur.wait_stream(self.down_stream)

    def segment(self, frames: torch.Tensor, out: torch.Tensor, boxes_out: Optional[torch.Tensor] = None) -> None:
        """Synchronous form: enqueue, then wait until ``out`` (and ``boxes_out``) are complete."""
        if boxes_out is None:
            self.segment_async(frames, out)
        else:
            self.segment_async(frames, out, boxes_out)
        self.synchronize()


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

    def segment(self, frames, out=None, return_boxes: bool = False):
        """uint8 frames [B,H,W,3] -> uint8 masks [B,3,H,W] (host).  Order is preserved.  With
        ``return_boxes`` the result is ``(masks, boxes)``, boxes int32 [B,3,5] = xmin, xmax, ymin, ymax,
        count per field (empty mask: W, -1, H, -1, 0), reduced on each GPU next to its masks."""
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
        boxes = torch.empty((b, 3, 5), dtype=torch.int32, pin_memory=pin) if return_boxes else None
        errors: List[BaseException] = []

        def work(rank: int):
            lo, hi = shard_bounds(b, len(self.workers), rank)
            if hi > lo:
                dev = getattr(self.workers[rank], "device", None)
                if len(self.workers) > 1 and dev is not None and dev.index is not None:
                    bind_to_gpu_numa(dev.index)
                try:
                    if boxes is None:
                        self.workers[rank].segment(frames[lo:hi], out[lo:hi])
                    else:
                        self.workers[rank].segment(frames[lo:hi], out[lo:hi], boxes[lo:hi])
                except BaseException as e:      # surfaced to the caller below
                    errors.append(e)

        threads = [threading.Thread(target=work, args=(r,)) for r in range(len(self.workers))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return out if boxes is None else (out, boxes)


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
