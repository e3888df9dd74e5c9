"""Host logic of the multi-GPU launcher on CPU: shard/chunk arithmetic, order-preserving
scatter/gather with fake workers, and the one-process-per-GPU gather over gloo (world 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tw_invoice_unet_ocr_llm_b200.launcher import (HostGather, MultiGpuSegmenter, chunk_bounds, gather_masks,
                                                   shard_bounds)


@pytest.mark.parametrize("total,world", [(512, 8), (512, 3), (5, 8), (0, 2), (64, 1), (1, 2)])
def test_shard_bounds_partition(total, world):
    spans = [shard_bounds(total, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and b >= a
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(total, world, world)


def test_chunk_bounds():
    assert chunk_bounds(0, 130, 64) == [(0, 64), (64, 128), (128, 130)]
    assert chunk_bounds(5, 5, 64) == []
    assert chunk_bounds(3, 10, 100) == [(3, 10)]
    with pytest.raises(ValueError):
        chunk_bounds(0, 1, 0)


class FakeWorker:
    """Stands in for GpuWorker: 'mask' = a deterministic function of each frame."""

    def __init__(self, dev):
        self.dev = dev
        self.seen = []

    def segment(self, frames, out):
        self.seen.append(frames.shape[0])
        f = frames.to(torch.int32)
        out.copy_((f.sum(dim=(1, 2))[:, :, None, None] % 251).to(torch.uint8).expand_as(out))


@pytest.mark.parametrize("b,ndev", [(10, 4), (3, 8), (64, 2), (1, 1)])
def test_multi_gpu_segmenter_preserves_order(b, ndev):
    rng = np.random.default_rng(b)
    frames = rng.integers(0, 256, (b, 16, 16, 3), dtype=np.uint8)
    seg = MultiGpuSegmenter(None, devices=list(range(ndev)), worker_factory=FakeWorker)
    out = seg.segment(frames)
    assert out.shape == (b, 3, 16, 16) and out.dtype == torch.uint8
    expect = (frames.astype(np.int64).sum(axis=(1, 2)) % 251).astype(np.uint8)
    assert np.array_equal(out[:, :, 0, 0].numpy(), expect)
    assert sum(sum(w.seen) for w in seg.workers) == b
    with pytest.raises(ValueError):
        seg.segment(np.zeros((2, 16, 16, 4), dtype=np.uint8))


def test_worker_error_propagates():
    class Boom(FakeWorker):
        def segment(self, frames, out):
            raise RuntimeError("device lost")
    seg = MultiGpuSegmenter(None, devices=[0, 1], worker_factory=Boom)
    with pytest.raises(RuntimeError, match="device lost"):
        seg.segment(np.zeros((4, 16, 16, 3), dtype=np.uint8))


def test_failed_worker_shard_is_requeued_on_survivors():
    """SURVEY section 5 (failure handling): a worker that raises loses nothing -- its shard is split over the
    workers that succeeded; the output is complete, in order, and the failure is counted."""
    class Flaky(FakeWorker):
        def segment(self, frames, out):
            if self.dev == 1:
                raise RuntimeError("device 1 lost")
            super().segment(frames, out)
    rng = np.random.default_rng(8)
    frames = rng.integers(0, 256, (13, 16, 16, 3), dtype=np.uint8)
    seg = MultiGpuSegmenter(None, devices=[0, 1, 2], worker_factory=Flaky)
    out = seg.segment(frames)
    expect = (frames.astype(np.int64).sum(axis=(1, 2)) % 251).astype(np.uint8)
    assert np.array_equal(out[:, :, 0, 0].numpy(), expect)
    assert seg.failures == [0, 1, 0]
    assert sum(sum(w.seen) for w in seg.workers) == 13           # every frame segmented exactly once
    assert seg.workers[1].seen == []


def test_packed_segmenter_allocates_bit_planes():
    class PackedFake:
        def __init__(self, dev):
            self.dev = dev

        def segment(self, frames, out):
            out.fill_(0xA5)
    seg = MultiGpuSegmenter(None, devices=[0, 1], worker_factory=PackedFake, packed=True)
    out = seg.segment(np.zeros((5, 16, 32, 3), dtype=np.uint8))
    assert out.shape == (5, 3, 16, 4) and bool((out == 0xA5).all())


def test_unpack_mask_bits_matches_numpy():
    from tw_invoice_unet_ocr_llm_b200.engine import unpack_mask_bits
    rng = np.random.default_rng(4)
    planes = rng.integers(0, 2, (2, 3, 8, 32), dtype=np.uint8)
    bits = np.packbits(planes, axis=-1, bitorder="little")
    assert np.array_equal(unpack_mask_bits(bits), planes)
    assert torch.equal(unpack_mask_bits(torch.from_numpy(bits)), torch.from_numpy(planes))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_rank(rank, world, port, total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(total, world, rank)
        local = torch.arange(lo, hi, dtype=torch.uint8).view(-1, 1, 1, 1).expand(-1, 3, 4, 4).contiguous()
        full = gather_masks(local, total)
        if rank == 0:
            q.put(full.numpy())
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def _gloo_rank_shm(rank, world, port, total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hg = HostGather(total, (3, 4, 2), pin=False)
        lo, hi = shard_bounds(total, world, rank)
        assert hg.local().shape == (hi - lo, 3, 4, 2)
        hg.local().copy_(torch.arange(lo, hi, dtype=torch.uint8).view(-1, 1, 1, 1).expand(-1, 3, 4, 2))
        hg.wait()
        if rank == 0:
            q.put(hg.full().numpy().copy())
        hg.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [7, 8])
def test_host_gather_shared_block_gloo_world2(total):
    """HostGather: every rank writes its shard into one shared host block in place (the device->host copy
    target of bench.py's sharded run); rank 0 reads the whole batch in input order."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_rank_shm, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert full.shape == (total, 3, 4, 2)
    assert np.array_equal(full[:, 0, 0, 0], np.arange(total, dtype=np.uint8))


@pytest.mark.parametrize("total", [7, 8])
def test_gather_masks_gloo_world2(total):
    """The N>1 layout of bench.py / torchrun: each rank owns a contiguous shard, rank 0 gathers
    the uint8 masks on the host; no GPU collective exists on the data path."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_rank, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert full.shape == (total, 3, 4, 4)
    assert np.array_equal(full[:, 0, 0, 0], np.arange(total, dtype=np.uint8))


def test_multi_gpu_segmenter_returns_boxes_in_order():
    """return_boxes=True: every shard writes its per-field extents next to its masks, input order kept."""
    class BoxWorker(FakeWorker):
        def segment(self, frames, out, boxes_out=None):
            super().segment(frames, out)
            if boxes_out is not None:
                tag = frames.to(torch.int32).sum(dim=(1, 2, 3)) % 1000
                boxes_out.copy_(tag[:, None, None].expand_as(boxes_out).to(torch.int32))
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (11, 8, 8, 3), dtype=np.uint8)
    seg = MultiGpuSegmenter(None, devices=list(range(4)), worker_factory=BoxWorker)
    masks, boxes = seg.segment(frames, return_boxes=True)
    assert masks.shape == (11, 3, 8, 8) and boxes.shape == (11, 3, 5) and boxes.dtype == torch.int32
    assert np.array_equal(boxes[:, 0, 0].numpy(), frames.astype(np.int64).sum(axis=(1, 2, 3)) % 1000)
    assert torch.equal(masks, seg.segment(frames))


def test_bind_to_gpu_numa_is_harmless_without_nvml():
    """No NVML / no GPU / one NUMA node: nothing changes and the call reports None."""
    from tw_invoice_unet_ocr_llm_b200.launcher import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    res = bind_to_gpu_numa(0)
    after = os.sched_getaffinity(0)
    assert res is None or set(res) == after
    if res is None:
        assert before == after
