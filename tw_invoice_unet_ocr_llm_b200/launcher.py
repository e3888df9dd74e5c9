"""Multi-GPU batch launcher (new subsystem; the reference is single-device, batch 1).

Images are independent units, so the batch is split into contiguous shards, one per GPU,
each GPU holds a full packed replica of the weights, and there is NO collective on the data
path: the host scatters uint8 frames (0.79 MB / image at 512x512) and gathers uint8 masks
(0.79 MB / image).  Results are bit-identical to a single-GPU run of the same images.

Two ways to drive it:

* in one process -- ``MultiGpuSegmenter``: one worker thread per GPU, each with its own
  engine, streams and pinned staging, chunks of <= 64 images, copies overlapped with compute;
* one process per GPU (``torchrun``) -- ``shard_bounds`` + ``HostGather`` (one POSIX shared-memory block,
  pinned in every rank, that each GPU's device->host copy writes its shard of the masks into directly: the
  gather costs no extra copy and no collective) or ``gather_masks`` (a host-side ``torch.distributed.gather``
  for groups without shared memory); ``bench.py`` uses this shape.

Masks can travel bit-packed (``packed=True``: uint8 ``[B, 3, H, W/8]``, one bit per pixel -- everything the
reference's boolean masks hold, reference inference.py:75-79 -- in an eighth of the PCIe bytes);
``engine.unpack_mask_bits`` restores the 0/1 planes.

A worker that fails (device lost, out of memory) does not lose the batch: its shard is re-queued on the
surviving GPUs (``MultiGpuSegmenter.segment``); only when no worker is left does the error propagate.
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
                 chunk: int = MAX_CHUNK, packed: bool = False):
        from .engine import Engine
        self.device = torch.device(device)
        self.thresholds = list(thresholds)
        self.chunk = chunk
        self.packed = bool(packed)       # masks leave the GPU as one bit per pixel: [B, 3, H, W/8]
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
        """Enqueue ``frames`` uint8 [B,H,W,3] (pinned host) -> ``out`` uint8 [B,3,H,W] (pinned host;
        [B,3,H,W/8] bit-packed for a ``packed`` worker) and return without waiting.  Per chunk: H2D on the upload stream, forward on the compute
        stream, D2H on the download stream; the upload of the next chunk (or of the next call) and
        the download of the previous one overlap the forward of the current one.  ``out`` is valid
        after :meth:`synchronize`.  ``boxes_out`` (optional, int32 [B,3,5] pinned host) also receives the
        mask extents {xmin, xmax, ymin, ymax, count} of reference inference.py:85-93, reduced on the GPU."""
        b, h, w = frames.shape[0], frames.shape[1], frames.shape[2]
        want = (b, 3, h, w // 8) if self.packed else (b, 3, h, w)
        if tuple(out.shape) != want or out.dtype != torch.uint8:
            raise ValueError(f"out must be uint8 {want} for {'packed ' if self.packed else ''}masks of "
                             f"{tuple(frames.shape)} frames, got {out.dtype} {tuple(out.shape)}")
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
                                    mask_out=mout[:n], mask_bits=self.packed)
                    box = None
                    if boxes_out is not None:
                        from . import prepost
                        box = prepost.mask_bbox_bits(mout[:n]) if self.packed else prepost.mask_bbox(mout[:n])
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
                 worker_factory: Optional[Callable] = None, packed: bool = False):
        if devices is None:
            devices = [f"cuda:{i}" for i in range(torch.cuda.device_count())]
        if len(devices) == 0:
            raise RuntimeError("MultiGpuSegmenter needs at least one CUDA device; there is no CPU path")
        self.packed = bool(packed)
        factory = worker_factory or (lambda dev: GpuWorker(state, dev, thresholds, chunk, packed=self.packed))
        self.workers = [factory(d) for d in devices]
        self.failures = [0] * len(self.workers)      # per worker: shards it failed (and others took over)

    def segment(self, frames, out=None, return_boxes: bool = False):
        """uint8 frames [B,H,W,3] -> uint8 masks [B,3,H,W] (host; [B,3,H,W/8] bit-packed for a ``packed``
        segmenter).  Order is preserved.  With ``return_boxes`` the result is ``(masks, boxes)``, boxes int32
        [B,3,5] = xmin, xmax, ymin, ymax, count per field (empty mask: W, -1, H, -1, 0), reduced on each GPU
        next to its masks.

        Failure handling: the batch is first split into one contiguous shard per worker.  If a worker raises,
        its shard is split again over the workers that succeeded and re-run there (results are bit-identical on
        any GPU, so the output does not depend on who ran what); the error propagates only when every worker has
        failed.  ``self.failures`` counts the shards each worker gave up."""
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(np.ascontiguousarray(frames))
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
            raise ValueError(f"expected uint8 [B,H,W,3] frames, got {frames.dtype} {tuple(frames.shape)}")
        b, h, w, _ = frames.shape
        pin = torch.cuda.is_available()
        if pin and not frames.is_pinned():
            frames = frames.pin_memory()
        if out is None:
            out = torch.empty((b, 3, h, w // 8 if self.packed else w), dtype=torch.uint8, pin_memory=pin)
        boxes = torch.empty((b, 3, 5), dtype=torch.int32, pin_memory=pin) if return_boxes else None

        def run_shard(rank: int, lo: int, hi: int):
            if boxes is None:
                self.workers[rank].segment(frames[lo:hi], out[lo:hi])
            else:
                self.workers[rank].segment(frames[lo:hi], out[lo:hi], boxes[lo:hi])

        alive = list(range(len(self.workers)))
        todo = [(lo, hi) for lo, hi in (shard_bounds(b, len(alive), r) for r in range(len(alive)))]
        first_error: Optional[BaseException] = None
        while True:
            failed: List[Tuple[int, int, int]] = []          # (rank, lo, hi)
            lock = threading.Lock()

            def work(rank: int, spans):
                nonlocal first_error
                dev = getattr(self.workers[rank], "device", None)
                if len(self.workers) > 1 and dev is not None and getattr(dev, "index", None) is not None:
                    bind_to_gpu_numa(dev.index)
                for lo, hi in spans:
                    if hi <= lo:
                        continue
                    try:
                        run_shard(rank, lo, hi)
                    except BaseException as e:      # the shard goes back to the queue below
                        with lock:
                            failed.append((rank, lo, hi))
                            if first_error is None:
                                first_error = e

            per_worker = {r: [] for r in alive}
            for i, span in enumerate(todo):
                per_worker[alive[i % len(alive)]].append(span)
            threads = [threading.Thread(target=work, args=(r, spans)) for r, spans in per_worker.items() if spans]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            if not failed:
                break
            for r, _, _ in failed:
                self.failures[r] += 1
            dead = {r for r, _, _ in failed}
            alive = [r for r in alive if r not in dead]
            if not alive:
                raise first_error
            # split every lost shard over the survivors
            todo = []
            for _, lo, hi in failed:
                todo += [(lo + a, lo + c) for a, c in (shard_bounds(hi - lo, len(alive), k) for k in range(len(alive)))
                         if c > a]
        return out if boxes is None else (out, boxes)


class HostGather:
    """Host-side gather of the per-rank mask shards for the one-process-per-GPU layout, without a copy: one
    POSIX shared-memory block of ``[total, *item_shape]`` uint8 that every rank maps and pins
    (``cudaHostRegister``), so each GPU's device->host copy lands its shard directly in rank 0's result.
    ``local()`` is the calling rank's ``shard_bounds`` slice (the D2H target), ``full()`` the whole array
    (meaningful on every rank after ``wait()``).  No GPU collective is involved; ``group`` only carries the
    block's name and the completion barrier (use a gloo group for the barrier when the GPUs should stay idle)."""

    def __init__(self, total: int, item_shape: Sequence[int], group=None, pin: Optional[bool] = None):
        import torch.distributed as dist
        from multiprocessing import shared_memory
        self._dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.total, self.item_shape = int(total), tuple(int(v) for v in item_shape)
        nbytes = max(1, self.total * int(np.prod(self.item_shape, dtype=np.int64)))
        name = [None]
        if self.rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name[0] = self._shm.name
        dist.broadcast_object_list(name, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if self.rank != 0:
            self._shm = shared_memory.SharedMemory(name=name[0])
            try:                                  # the creator (rank 0) owns the block's lifetime
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        arr = np.ndarray((self.total, *self.item_shape), dtype=np.uint8, buffer=self._shm.buf)
        self._tensor = torch.from_numpy(arr)
        self._registered = False
        if pin is None:
            pin = torch.cuda.is_available()
        if pin and nbytes > 0:
            rc = torch.cuda.cudart().cudaHostRegister(self._tensor.data_ptr(), nbytes, 0)
            self._registered = int(rc) == 0
        self.pinned = self._registered

    def local(self) -> torch.Tensor:
        lo, hi = shard_bounds(self.total, self.world, self.rank)
        return self._tensor[lo:hi]

    def full(self) -> torch.Tensor:
        return self._tensor

    def wait(self) -> None:
        """Every rank has finished writing its shard (call after the rank's own copies completed)."""
        self._dist.barrier(group=self.group)

    def close(self) -> None:
        if getattr(self, "_shm", None) is None:
            return
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self._tensor.data_ptr())
            self._registered = False
        self._tensor = None
        try:
            self._dist.barrier(group=self.group)      # nobody unlinks while a peer still maps the block
        except Exception:
            pass
        try:
            self._shm.close()
        except BufferError:
            pass
        if self.rank == 0:
            try:
                self._shm.unlink()
            except FileNotFoundError:
                pass
        self._shm = None


def gather_masks(local: torch.Tensor, total: int, group=None) -> Optional[torch.Tensor]:
    """Host-side gather for the one-process-per-GPU layout: every rank passes the CPU uint8
    masks of its ``shard_bounds`` range; rank 0 gets the full ``[total, ...]`` tensor in input
    order, other ranks get ``None``.  No GPU collective is involved: gloo groups use a host
    ``gather``; groups on a device-only backend (NCCL) go through a :class:`HostGather` shared-memory block
    (for the zero-copy form, let the device->host copies write into ``HostGather.local()`` directly)."""
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
        hg = HostGather(total, local.shape[1:], group=group, pin=False)
        hg.local().copy_(local)
        hg.wait()
        out = hg.full().clone() if rank == 0 else None
        hg.close()
        return out
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
