"""GPU: the CUDA crop enhancement (csrc/enhance.cuh through unetb200_enhance_plan / _run) against the
OpenCV oracle (oracle/opencv_enhance.py) and the golden vectors produced by the reference's own
enhance_for_ocrspace / enhance_for_date_ocr (app_camera.py:572-598, 685-705).  Byte work: bit-exact."""
import os

import numpy as np
import pytest
from PIL import Image

from oracle import opencv_enhance as oe
from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_crops_u8

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "golden_enhance.npz"))


def _oracle(rgb, kind):
    from tw_invoice_unet_ocr_llm_b200 import enhance
    flags, clip = enhance._recipe(kind)
    return oe.enhance(rgb, flags, clip)


def _check(crops, kinds):
    from tw_invoice_unet_ocr_llm_b200 import enhance
    got = enhance.enhance_batch(crops, kinds)
    for i, (rgb, kind) in enumerate(zip(crops, kinds)):
        want = _oracle(rgb, kind)
        assert got[i].shape == want.shape and got[i].dtype == np.uint8
        bad = int((got[i] != want).sum())
        assert bad == 0, f"crop {i} {rgb.shape} {kind}: {bad} of {want.size} pixels differ"
    return got


def test_golden_vectors(cuda_dev):
    from tw_invoice_unet_ocr_llm_b200 import enhance
    n = sum(1 for k in GOLD.files if k.startswith("crop_"))
    crops = [GOLD[f"crop_{i}"] for i in range(n)]
    for kind in ("text", "amount", "date"):
        got = enhance.enhance_batch(crops, [kind] * n)
        for i in range(n):
            assert np.array_equal(got[i], GOLD[f"{kind}_{i}"]), (kind, i)


def test_ragged_batch_all_kinds(cuda_dev):
    sizes = [(24, 70), (31, 45), (40, 121), (17, 18), (56, 200), (9, 33), (64, 64), (8, 8), (100, 30), (33, 257)]
    crops = synthetic_crops_u8(sizes, seed=21)
    kinds = [("text", "amount", "date")[i % 3] for i in range(len(crops))]
    _check(crops, kinds)
    _check(crops, kinds[1:] + kinds[:1])


def test_edge_sizes(cuda_dev):
    """1-pixel crops, single rows / columns, widths whose 4x is not a multiple of 8 (integer tail of the
    vertical pass), heights below the CLAHE grid (the reflect-101 extension wraps more than once)."""
    rng = np.random.default_rng(3)
    sizes = [(1, 1), (1, 2), (2, 1), (1, 37), (41, 1), (2, 2), (3, 5), (5, 3), (1, 8), (8, 1), (7, 9), (16, 15)]
    crops = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    for kind in ("text", "amount", "date"):
        _check(crops, [kind] * len(crops))


def test_every_flag_combination_and_clip_limits(cuda_dev):
    rng = np.random.default_rng(4)
    crops, kinds = [], []
    for flags in range(8):
        for clip in (4.0, 3.0, 1.0, 40.0, 0.01):
            h, w = int(rng.integers(5, 70)), int(rng.integers(5, 140))
            crops.append(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
            kinds.append((flags, clip))
    _check(crops, kinds)


def test_degenerate_content(cuda_dev):
    """Constant crops (Otsu never finds a split: threshold 0), two-level crops, saturated noise."""
    rng = np.random.default_rng(5)
    white = np.full((20, 50, 3), 255, np.uint8)
    black = np.zeros((20, 50, 3), np.uint8)
    gray = np.full((13, 21, 3), 128, np.uint8)
    two = np.where(rng.random((30, 60, 1)) < 0.2, 20, 235).astype(np.uint8).repeat(3, axis=2)
    sat = (rng.integers(0, 2, (25, 40, 3)) * 255).astype(np.uint8)
    crops = [white, black, gray, two, sat]
    for kind in ("text", "amount", "date"):
        _check(crops, [kind] * len(crops))


def test_large_crops(cuda_dev):
    """A whole 720p frame as one crop (2880 x 5120 output) next to small ones."""
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    frame = synthetic_invoices_u8(1, 720, 1280, seed=31)[0]
    crops = [frame, frame[:301, :533], frame[100:140, 200:420]]
    _check(crops, ["text", "date", "amount"])


def test_batch_equals_single_and_is_deterministic(cuda_dev):
    from tw_invoice_unet_ocr_llm_b200 import enhance
    sizes = [(30, 90), (22, 47), (48, 160), (12, 12)] * 8
    crops = synthetic_crops_u8(sizes, seed=41)
    kinds = [("text", "date", "amount")[i % 3] for i in range(len(crops))]
    a = enhance.enhance_batch(crops, kinds)
    b = enhance.enhance_batch(crops, kinds)
    for i, (c, k) in enumerate(zip(crops, kinds)):
        assert np.array_equal(a[i], b[i])
        if i < 6:
            assert np.array_equal(a[i], enhance.enhance_batch([c], [k])[0])


def test_drop_in_functions(cuda_dev):
    """Names, arguments and return values of app_camera.py:572-598 / :685-705."""
    from tw_invoice_unet_ocr_llm_b200 import enhance
    rgb = synthetic_crops_u8([(28, 96)], seed=51)[0]
    pil = Image.fromarray(rgb)
    assert enhance.enhance_for_ocrspace(None) is None and enhance.enhance_for_date_ocr(None) is None
    t = enhance.enhance_for_ocrspace(pil)
    a = enhance.enhance_for_ocrspace(pil, mode="amount")
    d = enhance.enhance_for_date_ocr(pil)
    for im in (t, a, d):
        assert isinstance(im, Image.Image) and im.mode == "L" and im.size == (4 * 96, 4 * 28)
    assert np.array_equal(np.array(t), oe.enhance_for_ocrspace(rgb, "text"))
    assert np.array_equal(np.array(a), oe.enhance_for_ocrspace(rgb, "amount"))
    assert np.array_equal(np.array(d), oe.enhance_for_date_ocr(rgb))
    assert set(np.unique(np.array(t))) <= {0, 255}
    # non-RGB crops go through PIL's convert("RGB") like the reference (:581)
    g = enhance.enhance_for_ocrspace(pil.convert("L"), mode="amount")
    assert np.array_equal(np.array(g), oe.enhance_for_ocrspace(np.array(pil.convert("L").convert("RGB")), "amount"))
    # None entries pass through the batched entry point
    out = enhance.enhance_batch([None, pil, None], ["text", "date", "amount"])
    assert out[0] is None and out[2] is None and np.array_equal(out[1], np.array(d))


def test_bad_arguments(cuda_dev):
    from tw_invoice_unet_ocr_llm_b200 import enhance
    with pytest.raises(ValueError):
        enhance.enhance_batch([np.zeros((4, 4, 3), np.uint8)], ["bogus"])
    with pytest.raises(ValueError):
        enhance.enhance_batch([np.zeros((4, 4), np.uint8)], ["text"])
    with pytest.raises(ValueError):
        enhance.enhance_batch([np.zeros((4, 4, 3), np.uint8)], ["text", "date"])


def test_windows_of_a_device_frame_equal_packed_crops(cuda_dev):
    """enhance_windows reads the crops in place from a frame in HBM (row stride = frame width)."""
    import torch
    from tw_invoice_unet_ocr_llm_b200 import enhance
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    frame = synthetic_invoices_u8(1, 360, 640, seed=71)[0]
    rects = [(0, 0, 640, 360), (3, 5, 4, 6), (639, 359, 640, 360), (17, 40, 230, 91), (401, 100, 640, 133), (0, 300, 77, 360)]
    kinds = ["text", "amount", "date", "text", "date", "amount"]
    got = enhance.enhance_windows(torch.from_numpy(frame).to(cuda_dev), rects, kinds)
    packed = enhance.enhance_batch([frame[y1:y2, x1:x2] for x1, y1, x2, y2 in rects], kinds)
    for g, p, (x1, y1, x2, y2), k in zip(got, packed, rects, kinds):
        assert np.array_equal(g, p)
        assert np.array_equal(g, _oracle(np.ascontiguousarray(frame[y1:y2, x1:x2]), k))
    with pytest.raises(ValueError):
        enhance.enhance_windows(torch.from_numpy(frame).to(cuda_dev), [(0, 0, 641, 10)], ["text"])
    assert enhance.enhance_windows(torch.from_numpy(frame).to(cuda_dev), [], []) == []


def test_concurrent_threads(cuda_dev):
    """Streamlit runs every session in its own thread: concurrent enhance_batch / enhance_for_* calls (shared
    library state: none; pinned blocks from torch's caching host allocator) give the single-threaded results."""
    import threading
    from tw_invoice_unet_ocr_llm_b200 import enhance
    sizes = [(20 + 3 * i, 60 + 7 * i) for i in range(6)]
    crops = synthetic_crops_u8(sizes, seed=91)
    kinds = [("text", "date", "amount")[i % 3] for i in range(len(crops))]
    want = [_oracle(c, k) for c, k in zip(crops, kinds)]
    errors = []

    def work(tid):
        try:
            for rep in range(6):
                order = list(range(len(crops)))[tid % 3:] + list(range(len(crops)))[:tid % 3]
                got = enhance.enhance_batch([crops[i] for i in order], [kinds[i] for i in order])
                for g, i in zip(got, order):
                    assert np.array_equal(g, want[i]), (tid, rep, i)
                one = enhance.enhance_for_date_ocr(Image.fromarray(crops[tid % len(crops)]))
                assert np.array_equal(np.array(one), oe.enhance_for_date_ocr(crops[tid % len(crops)]))
        except BaseException as e:      # surfaced below
            errors.append(e)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[0]


def test_gpu_results_against_stock_ipp_opencv_are_bounded(cuda_dev):
    """The CUDA enhancement against the reference's calls run with the INSTALLED cv2 (IPP-routed resize, as in
    the reference's pinned stock wheel): not bit-exact by construction (the kernels follow OpenCV's own code
    path), but inside the bounds tests/test_enhance_oracle.py states for the final images."""
    pytest.importorskip("cv2")
    import sys
    sys.path.insert(0, HERE)
    from enhance_common import IPP_BOUNDS, stock_cv2_chain as _stock_cv2_chain
    from tw_invoice_unet_ocr_llm_b200 import enhance
    rng = np.random.default_rng(6)
    sizes = [(int(rng.integers(8, 72)), int(rng.integers(20, 320))) for _ in range(24)]
    crops = synthetic_crops_u8(sizes, seed=78)
    for kind, (overall, worst, maxdiff) in IPP_BOUNDS.items():
        got = enhance.enhance_batch(crops, [kind] * len(crops))
        bad = tot = 0
        for c, g in zip(crops, got):
            d = np.abs(_stock_cv2_chain(c, kind).astype(int) - g.astype(int))
            assert d.max() <= maxdiff and (d != 0).mean() <= worst, (kind, c.shape, int(d.max()), float((d != 0).mean()))
            bad += int((d != 0).sum())
            tot += d.size
        assert bad <= overall * tot, (kind, bad, tot)
