"""Synthetic stand-ins for what is not shipped with the reference.

* ``synthetic_invoices`` -- invoice-shaped images: light paper
  background, dark text-stroke rows, two QR-like module blocks, quantised to k/255
  like ``inference.preprocess`` output (reference inference.py:35-36).
* ``make_fixture_state`` -- a checkpoint in exactly the format of
  ``checkpoints/best_unet_model.pth`` (a plain fp32 ``state_dict`` with 136 keys,
  train.py:159).  The real file is a Git-LFS pointer (124 267 083 B) that cannot be
  fetched, so tests and benches use this deterministic fixture: seeded default init,
  random BatchNorm affine terms, BatchNorm statistics calibrated by a few train-mode
  passes, and an output head scaled so the logits straddle the thresholds of
  inference.py:76-78 with a background-dominated (mostly negative) distribution, as
  a trained segmenter has (``out_conv.bias`` starts at -4, unet_model.py:53).
"""
from __future__ import annotations

import numpy as np
import torch


def synthetic_invoices_u8(n: int, h: int = 512, w: int = 512, seed: int = 7) -> np.ndarray:
    """uint8 ``(n, h, w, 3)`` invoice-like frames."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, h, w, 3), dtype=np.uint8)
    for i in range(n):
        img = 0.92 + 0.03 * rng.standard_normal((h, w, 1)).astype(np.float32)
        img = np.repeat(img, 3, axis=2)
        img += 0.01 * rng.standard_normal((h, w, 3)).astype(np.float32)
        n_rows = max(4, (40 * h) // 512)
        for _ in range(n_rows):
            rh = int(rng.integers(max(2, h // 128), max(3, h // 42)))
            y = int(rng.integers(0, max(1, h - rh)))
            length = int(rng.integers(max(8, (40 * w) // 512), max(9, w // 2)))
            x = int(rng.integers(0, max(1, w - length)))
            ink = rng.random((rh, length, 1)) < 0.45
            img[y:y + rh, x:x + length] = np.where(ink, 0.12 + 0.05 * rng.random((rh, length, 3)),
                                                   img[y:y + rh, x:x + length])
        mod = max(1, (4 * w) // 512)
        for _ in range(2):
            side = 22 * mod
            if side >= min(h, w):
                break
            y = int(rng.integers(0, h - side))
            x = int(rng.integers(0, w - side))
            modules = (rng.random((22, 22)) < 0.5).astype(np.float32)
            block = np.kron(modules, np.ones((mod, mod), dtype=np.float32))[..., None]
            img[y:y + side, x:x + side] = 0.1 + 0.85 * block
        out[i] = np.clip(np.rint(np.clip(img, 0.0, 1.0) * 255.0), 0, 255).astype(np.uint8)
    return out


def synthetic_invoices(n: int, h: int = 512, w: int = 512, seed: int = 7) -> torch.Tensor:
    """float32 ``[n, 3, h, w]`` with values k/255 (what ``preprocess`` produces)."""
    u8 = synthetic_invoices_u8(n, h, w, seed)
    return torch.from_numpy(u8.astype(np.float32) / 255.0).permute(0, 3, 1, 2).contiguous()


@torch.no_grad()
def make_fixture_state(seed: int = 1234, calib_size: int = 64, calib_batches: int = 3,
                       model_cls=None) -> dict:
    """Deterministic fixture ``state_dict`` (CPU fp32, 136 keys).  ``model_cls`` defaults to the
    package's ``UNet``; pass the reference class to build the identical fixture with it."""
    if model_cls is None:
        from .unet_model import UNet as model_cls
    gen_state = torch.random.get_rng_state()
    try:
        torch.manual_seed(seed)
        model = model_cls(n_channels=3, n_classes=3)
        g = torch.Generator().manual_seed(seed + 1)
        bns = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d)]
        for m in bns:
            m.weight.copy_(0.5 + torch.rand(m.weight.shape, generator=g))
            m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=g))
            m.momentum = None            # cumulative average over the calibration passes
            m.reset_running_stats()
        # plain torch.nn graph of either class (the package's eval forward is CUDA-only)
        fwd = getattr(model, "_forward_train", None) or model.forward
        model.train()
        for b in range(calib_batches):
            fwd(synthetic_invoices(4, calib_size, calib_size, seed=100 + b))
        for m in bns:
            m.momentum = 0.1
        model.eval()                     # BatchNorm now uses the calibrated running statistics
        # head: spread the logits around the thresholds with a negative (background) mean
        feats = {}
        hook = model.conv1.register_forward_hook(lambda _m, _i, o: feats.__setitem__("c8", o))
        fwd(synthetic_invoices(2, calib_size, calib_size, seed=200))
        hook.remove()
        c8 = feats["c8"]
        wgt = torch.randn((3, 64, 1, 1), generator=g)
        z = torch.nn.functional.conv2d(c8, wgt)
        zs = z.std(dim=(0, 2, 3))
        zm = z.mean(dim=(0, 2, 3))
        scale = 1.6 / zs
        model.out_conv.weight.copy_(wgt * scale.view(3, 1, 1, 1))
        model.out_conv.bias.copy_(-3.6 - zm * scale)
        state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    finally:
        torch.random.set_rng_state(gen_state)
    return state


def synthetic_crops_u8(sizes, seed: int = 11) -> list:
    """uint8 ``(h, w, 3)`` field crops of the given ``(h, w)`` sizes, the kind ``run_unet`` hands to
    the OCR stage: a window of a synthetic invoice around dark strokes on light paper."""
    out = []
    for i, (h, w) in enumerate(sizes):
        page = synthetic_invoices_u8(1, max(64, 2 * h), max(64, 2 * w), seed=seed + i)[0]
        out.append(np.ascontiguousarray(page[:h, :w]))
    return out


def marked_invoices(n: int, size: int, seed: int):
    """Synthetic invoices with three marked fields (solid bar / stripes / checker = the three classes) and their
    masks: (float32 [n,3,size,size] with values k/255, float32 masks [n,3,size,size]).  Marks have the same
    absolute size at every resolution, so a model trained at 128x128 segments them at 512x512."""
    rng = np.random.default_rng(seed)
    img = synthetic_invoices_u8(n, size, size, seed=seed).astype(np.float32) / 255.0
    masks = np.zeros((n, 3, size, size), np.float32)
    for i in range(n):
        for c in range(3):
            w, h = int(rng.integers(24, 48)), int(rng.integers(8, 14))
            x, y = int(rng.integers(0, size - w)), int(rng.integers(0, size - h))
            yy, xx = np.mgrid[0:h, 0:w]
            if c == 0:
                patch = np.full((h, w), 0.05)                                  # solid dark bar
            elif c == 1:
                patch = 0.5 + 0.45 * (((xx // 3) % 2) * 2 - 1)                 # vertical stripes
            else:
                patch = 0.5 + 0.45 * ((((xx // 4) + (yy // 4)) % 2) * 2 - 1)   # checker
            img[i, y:y + h, x:x + w, :] = patch[..., None]
            masks[i, c, y:y + h, x:x + w] = 1.0
    x = torch.from_numpy(np.round(img * 255) / 255.0).float().permute(0, 3, 1, 2).contiguous()
    return x, torch.from_numpy(masks)


def invoice_loss(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """0.85 * multi-label Dice + 0.15 * focal on sigmoid probabilities (the reference's recipe, train.py:18-59)."""
    p = torch.sigmoid(logits)
    inter = (p * target).sum(dim=(2, 3))
    dice = 1 - ((2 * inter + 1) / (p.sum(dim=(2, 3)) + target.sum(dim=(2, 3)) + 1)).mean()
    bce = torch.nn.functional.binary_cross_entropy(p, target, reduction="none")
    pt = torch.where(target > 0.5, p, 1 - p)
    focal = (0.25 * (1 - pt) ** 2 * bce).mean()
    return 0.85 * dice + 0.15 * focal


def train_fixture_state(device, steps: int = 300, seed: int = 0, model_cls=None):
    """A TRAINED stand-in for ``checkpoints/best_unet_model.pth`` (the real file is a Git-LFS pointer): the
    UNet trained for ``steps`` AdamW steps (train.py:119-123) with ``invoice_loss`` on ``marked_invoices``.
    Plain torch train-mode ops -- this manufactures realistic weights (trained BatchNorm statistics, bimodal
    logits), it is not the product path.  Returns ``(state_dict on CPU, final loss)``."""
    if model_cls is None:
        from .unet_model import UNet as model_cls
    gen_state = torch.random.get_rng_state()
    try:
        torch.manual_seed(seed)
        model = model_cls(3, 3).to(device).train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
        xs, ms = marked_invoices(96, 128, seed=300)
        xs, ms = xs.to(device), ms.to(device)
        loss = None
        for _ in range(steps):
            idx = torch.randint(0, xs.shape[0], (8,), device=device)
            loss = invoice_loss(model(xs[idx]), ms[idx])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        model.eval()
        state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    finally:
        torch.random.set_rng_state(gen_state)
    return state, (float(loss.detach()) if loss is not None else float("nan"))
