"""GPU pre/post-processing around the forward, via ``libunetb200.so``:

* ``resize_u8``  -- ``PIL.Image.resize((512, 512))`` of reference inference.py:35,63 for uint8
  frames already on the device, bit-identical to Pillow's 8-bit bicubic resample;
* ``mask_bbox``  -- the ``np.where(mask)`` min/max of reference inference.py:85-93 as 5 ints per
  (image, class) instead of a 512x512 mask.
"""
from __future__ import annotations

import threading
from typing import Dict, Tuple

import numpy as np
import torch

from . import _native as nat

_table_cache: Dict[Tuple[int, int, str], Tuple[torch.Tensor, torch.Tensor, int]] = {}
_lock = threading.Lock()


def resize_tables_host(in_size: int, out_size: int):
    """Pillow's fixed-point coefficient table of one axis: (kk int32 [out, ksize], bounds int32 [out, 2])."""
    lib = nat.lib()
    ks = lib.unetb200_resize_ksize(in_size, out_size)
    if ks <= 0:
        raise ValueError(f"bad resize {in_size} -> {out_size}")
    kk = np.zeros((out_size, ks), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    nat.check(lib.unetb200_resize_coeffs(in_size, out_size, kk.ctypes.data, bounds.ctypes.data))
    return kk, bounds


def _tables(in_size: int, out_size: int, device: torch.device):
    key = (in_size, out_size, str(device))
    with _lock:
        t = _table_cache.get(key)
        if t is None:
            kk, bounds = resize_tables_host(in_size, out_size)
            t = (torch.from_numpy(kk).to(device), torch.from_numpy(bounds).to(device), kk.shape[1])
            if len(_table_cache) > 64:
                _table_cache.clear()
            _table_cache[key] = t
    return t


def resize_u8(frames: torch.Tensor, oh: int, ow: int, out: torch.Tensor = None, channels: int = None) -> torch.Tensor:
    """uint8 ``[N, H, W, C]`` on a CUDA device -> uint8 ``[N, oh, ow, C]``, bit-identical to
    ``PIL.Image.resize((ow, oh))`` (default BICUBIC) of each frame.  Enqueued on the current stream.
    ``out``: optional contiguous destination (e.g. a slot of a batch tensor).  ``channels`` < C reads only
    the leading channels of every pixel (``channels=3`` of a ``[N, H, W, 4]`` RGBX frame, Pillow's own storage
    of an RGB image) and returns ``[N, oh, ow, channels]``."""
    if frames.dtype != torch.uint8 or frames.dim() != 4 or not frames.is_cuda:
        raise RuntimeError("resize_u8 expects a CUDA uint8 [N,H,W,C] tensor")
    frames = frames.contiguous()
    n, h, w, ps = frames.shape
    c = ps if channels is None else int(channels)
    if not (1 <= c <= ps):
        raise RuntimeError(f"resize_u8: channels={c} of a {ps}-byte pixel")
    dev = frames.device
    if out is None:
        out = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=dev)
    elif (out.dtype != torch.uint8 or tuple(out.shape) != (n, oh, ow, c) or out.device != dev
          or not out.is_contiguous()):
        raise RuntimeError("resize_u8: out must be a contiguous uint8 [N,oh,ow,C] tensor on the same device")
    if ow == w and c != ps:          # no horizontal pass to drop the padding byte in: repack first
        frames, ps = frames[..., :c].contiguous(), c
    if oh == h and ow == w:
        out.copy_(frames)
        return out
    kx = bx = ky = by = tmp = None
    ksx = ksy = 0
    if ow != w:
        kx, bx, ksx = _tables(w, ow, dev)
    if oh != h:
        ky, by, ksy = _tables(h, oh, dev)
    if ow != w and oh != h:
        tmp = torch.empty((n, h, ow, c), dtype=torch.uint8, device=dev)
    p = lambda t: None if t is None else t.data_ptr()
    with torch.cuda.device(dev):
        nat.check(nat.lib().unetb200_resize_bicubic_u8_ps(
            frames.data_ptr(), n, h, w, c, ps, p(kx), p(bx), ksx, p(ky), p(by), ksy, p(tmp), out.data_ptr(),
            oh, ow, torch.cuda.current_stream(dev).cuda_stream))
    return out


def mask_bbox(mask: torch.Tensor) -> torch.Tensor:
    """uint8 ``[N, C, H, W]`` (non-zero = set) -> int32 ``[N, C, 5]`` = xmin, xmax, ymin, ymax, count
    (empty plane: W, -1, H, -1, 0).  Enqueued on the current stream."""
    if mask.dtype != torch.uint8 or mask.dim() != 4 or not mask.is_cuda:
        raise RuntimeError("mask_bbox expects a CUDA uint8 [N,C,H,W] tensor")
    mask = mask.contiguous()
    n, c, h, w = mask.shape
    out = torch.empty((n, c, 5), dtype=torch.int32, device=mask.device)
    with torch.cuda.device(mask.device):
        nat.check(nat.lib().unetb200_mask_bbox(mask.data_ptr(), n * c, h, w, out.data_ptr(),
                                               torch.cuda.current_stream(mask.device).cuda_stream))
    return out


def mask_bbox_bits(bits: torch.Tensor) -> torch.Tensor:
    """Bit-packed masks uint8 ``[N, C, H, W/8]`` (``Engine.run(mask_bits=True)``) -> int32 ``[N, C, 5]`` as
    :func:`mask_bbox`.  ``W`` must be a multiple of 32.  Enqueued on the current stream."""
    if bits.dtype != torch.uint8 or bits.dim() != 4 or not bits.is_cuda:
        raise RuntimeError("mask_bbox_bits expects a CUDA uint8 [N,C,H,W/8] tensor")
    bits = bits.contiguous()
    n, c, h, wb = bits.shape
    if wb % 4:
        raise RuntimeError("mask_bbox_bits needs a mask width that is a multiple of 32")
    out = torch.empty((n, c, 5), dtype=torch.int32, device=bits.device)
    with torch.cuda.device(bits.device):
        nat.check(nat.lib().unetb200_mask_bbox_bits(bits.data_ptr(), n * c, h, wb * 8, out.data_ptr(),
                                                    torch.cuda.current_stream(bits.device).cuda_stream))
    return out


def box_sums(frame: torch.Tensor, rects, channels: int = None) -> torch.Tensor:
    """uint8 ``[H, W, C]`` frame on a CUDA device + up to 16 half-open rectangles ``(x1, y1, x2, y2)``
    -> int64 ``[n]`` byte sums on the device (the ``np.array(crop).mean() < 3`` test of reference
    inference.py:121-125 becomes ``sum < 3 * area * channels``).  ``channels=3`` on a ``[H, W, 4]`` RGBX frame
    leaves the padding byte out.  Enqueued on the current stream."""
    if frame.dtype != torch.uint8 or frame.dim() != 3 or not frame.is_cuda or not frame.is_contiguous():
        raise RuntimeError("box_sums expects a contiguous CUDA uint8 [H,W,C] tensor")
    import ctypes as C
    h, w, c = frame.shape
    used = c if channels is None else int(channels)
    n = len(rects)
    flat = (C.c_int32 * (4 * n))(*[int(v) for r in rects for v in r])
    out = torch.empty((n,), dtype=torch.int64, device=frame.device)
    with torch.cuda.device(frame.device):
        nat.check(nat.lib().unetb200_box_sums_ps(frame.data_ptr(), h, w, c, used, flat, n, out.data_ptr(),
                                                 torch.cuda.current_stream(frame.device).cuda_stream))
    return out
