"""CPU oracle: a restatement of the reference's U-Net forward and mask/crop logic.

TEST INFRASTRUCTURE ONLY.  Nothing in ``tw_invoice_unet_ocr_llm_b200/`` imports this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may, and only as the checker or the timed CPU baseline.

What it restates (file:line into the reference repository):

* ``oracle_forward``      -- ``UNet.forward``        unet_model.py:55-86
  (``_double_conv``       -- ``DoubleConv.forward``  unet_model.py:9-20)
* ``oracle_masks``        -- sigmoid + thresholds     inference.py:72-79
* ``oracle_crop_boxes``   -- mask -> crop rectangle   inference.py:84-112
* ``numpy_forward``       -- the same forward in plain numpy loops/einsum (tiny inputs only);
  an independent statement of the layer definitions (zero padding, cat order,
  ConvTranspose2d weight layout) that does not go through torch's conv kernels.

The arithmetic of the reference lives in a third-party dependency, PyTorch (pinned
``torch==2.1.0`` in the reference's requirements.txt:11; this image has 2.11.0): the
oracle calls the same ``torch.nn.functional`` operators on CPU in fp32, functionally,
straight from a ``state_dict`` -- no ``nn.Module`` and none of the reference's source.

Pinning: the reference ships no tests or golden vectors.  The
oracle is pinned instead against outputs of the reference itself, executed in the build
container from /root/reference by ``tests/golden/make_golden.py``; the resulting
vectors are committed under ``tests/golden/`` and checked by
``tests/test_oracle_golden.py`` (which also compares live against the imported reference
modules where /root/reference exists, i.e. in the build container).
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5                       # nn.BatchNorm2d default (unet_model.py:11,15)
IMG_SIZE = 512                      # inference.py:10
FIELDS = ["invoice_no", "date", "total_amount"]            # inference.py:12
THRESHOLDS = {"invoice_no": 0.25, "date": 0.40, "total_amount": 0.30}   # inference.py:76-78

_ENC = ["down1", "down2", "down3", "down4"]


def _double_conv(s: Mapping[str, torch.Tensor], prefix: str, x: torch.Tensor) -> torch.Tensor:
    """Conv3x3(pad 1) -> BatchNorm(eval) -> ReLU, twice (unet_model.py:9-17)."""
    for conv, bn in ((0, 1), (3, 4)):
        x = F.conv2d(x, s[f"{prefix}.net.{conv}.weight"], s[f"{prefix}.net.{conv}.bias"], padding=1)
        x = F.batch_norm(x, s[f"{prefix}.net.{bn}.running_mean"], s[f"{prefix}.net.{bn}.running_var"],
                         s[f"{prefix}.net.{bn}.weight"], s[f"{prefix}.net.{bn}.bias"],
                         training=False, eps=BN_EPS)
        x = F.relu(x)
    return x


@torch.no_grad()
def oracle_forward(state: Mapping[str, torch.Tensor], x: torch.Tensor,
                   taps: Optional[dict] = None) -> torch.Tensor:
    """fp32 CPU logits ``[N, n_classes, H, W]`` for ``x`` ``[N, C, H, W]`` (unet_model.py:55-86).

    ``taps``: optional dict that receives the intermediate activations by reference name."""
    x = x.detach().to("cpu", torch.float32)
    s = {k: v.detach().to("cpu") for k, v in state.items()}
    skips = []
    for name in _ENC:                                   # :56-66
        x = _double_conv(s, name, x)
        skips.append(x)
        if taps is not None:
            taps[name] = x
        x = F.max_pool2d(x, 2)
    x = _double_conv(s, "bottleneck", x)                # :68
    if taps is not None:
        taps["bottleneck"] = x
    for lvl in (4, 3, 2, 1):                            # :70-84
        x = F.conv_transpose2d(x, s[f"up{lvl}.weight"], s[f"up{lvl}.bias"], stride=2)
        if taps is not None:
            taps[f"up{lvl}"] = x
        x = torch.cat([x, skips[lvl - 1]], dim=1)       # upsampled first, then the skip (:71)
        x = _double_conv(s, f"conv{lvl}", x)
        if taps is not None:
            taps[f"conv{lvl}"] = x
    return F.conv2d(x, s["out_conv.weight"], s["out_conv.bias"])   # :86, raw logits


def oracle_masks(logits: torch.Tensor) -> np.ndarray:
    """bool ``(N, 3, H, W)``: ``sigmoid(logit) > threshold`` per field (inference.py:72-79)."""
    prob = torch.sigmoid(logits).cpu().numpy()
    # python-float scalars, as in the reference: numpy compares them in prob's dtype (fp32)
    return np.stack([prob[:, c] > THRESHOLDS[f] for c, f in enumerate(FIELDS)], axis=1)


def logit_thresholds() -> list:
    """The same thresholds in logit space: sigmoid(z) > t  <=>  z > ln(t/(1-t))."""
    return [math.log(THRESHOLDS[f] / (1.0 - THRESHOLDS[f])) for f in FIELDS]


def oracle_crop_boxes(masks: Dict[str, np.ndarray], ow: int, oh: int
                      ) -> Dict[str, Optional[Tuple[int, int, int, int]]]:
    """Crop rectangle in ORIGINAL-image pixels for each field, or None (inference.py:84-112)."""
    out: Dict[str, Optional[Tuple[int, int, int, int]]] = {}
    for key, mask in masks.items():
        ys, xs = np.where(mask)
        if len(xs) == 0 or len(ys) == 0:
            out[key] = None
            continue
        mx1, mx2, my1, my2 = xs.min(), xs.max(), ys.min(), ys.max()
        scale_x, scale_y = ow / IMG_SIZE, oh / IMG_SIZE
        x1, x2, y1, y2 = int(mx1 * scale_x), int(mx2 * scale_x), int(my1 * scale_y), int(my2 * scale_y)
        pad_x, pad_y = int((x2 - x1) * 0.15), int((y2 - y1) * 0.15)
        x1, y1 = max(0, x1 - pad_x), max(0, y1 - pad_y)
        x2, y2 = min(ow, x2 + pad_x), min(oh, y2 + pad_y)
        out[key] = None if (x2 <= x1 or y2 <= y1) else (x1, y1, x2, y2)
    return out


# --------------------------------------------------------------------------- numpy restatement
def _np_conv3x3(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    n, c, h, wd = x.shape
    xp = np.zeros((n, c, h + 2, wd + 2), dtype=np.float64)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((n, w.shape[0], h, wd), dtype=np.float64)
    for ky in range(3):
        for kx in range(3):
            out += np.einsum("nchw,oc->nohw", xp[:, :, ky:ky + h, kx:kx + wd], w[:, :, ky, kx])
    return out + b.reshape(1, -1, 1, 1)


def _np_double_conv(s, prefix, x):
    for conv, bn in ((0, 1), (3, 4)):
        x = _np_conv3x3(x, s[f"{prefix}.net.{conv}.weight"], s[f"{prefix}.net.{conv}.bias"])
        mu, var = s[f"{prefix}.net.{bn}.running_mean"], s[f"{prefix}.net.{bn}.running_var"]
        g, be = s[f"{prefix}.net.{bn}.weight"], s[f"{prefix}.net.{bn}.bias"]
        x = (x - mu.reshape(1, -1, 1, 1)) / np.sqrt(var.reshape(1, -1, 1, 1) + BN_EPS)
        x = np.maximum(x * g.reshape(1, -1, 1, 1) + be.reshape(1, -1, 1, 1), 0.0)
    return x


def numpy_forward(state: Mapping[str, torch.Tensor], x: np.ndarray) -> np.ndarray:
    """float64 numpy restatement of unet_model.py:55-86 for tiny inputs (e.g. 1x3x16x16)."""
    s = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in state.items() if v.dtype.is_floating_point}
    x = x.astype(np.float64)
    skips = []
    for name in _ENC:
        x = _np_double_conv(s, name, x)
        skips.append(x)
        n, c, h, w = x.shape
        x = x.reshape(n, c, h // 2, 2, w // 2, 2).max(axis=(3, 5))
    x = _np_double_conv(s, "bottleneck", x)
    for lvl in (4, 3, 2, 1):
        wt, bt = s[f"up{lvl}.weight"], s[f"up{lvl}.bias"]          # [Cin, Cout, 2, 2]
        n, c, h, w = x.shape
        up = np.zeros((n, wt.shape[1], 2 * h, 2 * w), dtype=np.float64)
        for a in range(2):
            for b in range(2):
                up[:, :, a::2, b::2] = np.einsum("nchw,co->nohw", x, wt[:, :, a, b])
        up += bt.reshape(1, -1, 1, 1)
        x = _np_double_conv(s, f"conv{lvl}", np.concatenate([up, skips[lvl - 1]], axis=1))
    w1 = s["out_conv.weight"][:, :, 0, 0]
    return np.einsum("nchw,oc->nohw", x, w1) + s["out_conv.bias"].reshape(1, -1, 1, 1)


# --------------------------------------------------------------------------- parity metrics
def parity_report(z_ref: torch.Tensor, z_new: torch.Tensor,
                  thr_logit: Optional[Sequence[float]] = None) -> dict:
    """Logit errors + mask agreement / IoU of ``z_new`` against the oracle's ``z_ref``."""
    z_ref = z_ref.detach().double().cpu()
    z_new = z_new.detach().double().cpu()
    thr = torch.tensor(thr_logit if thr_logit is not None else logit_thresholds(),
                       dtype=torch.float64).view(1, -1, 1, 1)
    d = (z_ref - z_new).abs()
    m_ref, m_new = z_ref > thr, z_new > thr
    agree = (m_ref == m_new)
    rep = {
        "max_abs": float(d.max()),
        "mean_abs": float(d.mean()),
        "logit_std": float(z_ref.std()),
        "max_abs_over_std": float(d.max() / z_ref.std()),
        "agreement": float(agree.double().mean()),
        "positive_frac": float(m_ref.double().mean()),
    }
    for band in (0.05, 0.1, 0.25):
        keep = (z_ref - thr).abs() > band
        rep[f"agreement_outside_{band}"] = float(agree[keep].double().mean()) if keep.any() else 1.0
    ious = []
    for c in range(z_ref.shape[1]):
        inter = float((m_ref[:, c] & m_new[:, c]).sum())
        union = float((m_ref[:, c] | m_new[:, c]).sum())
        ious.append(1.0 if union == 0 else inter / union)
    rep["iou"] = ious
    return rep
