"""CPU: pins oracle/opencv_enhance.py (the restatement of app_camera.py:572-598, 685-705) against
(1) golden vectors produced by the reference's own two functions (tests/golden/make_golden_enhance.py)
and (2) OpenCV executed in this process, step by step."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import opencv_enhance as oe
from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_crops_u8

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from enhance_common import IPP_BOUNDS, stock_cv2_chain as _stock_cv2_chain  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "golden_enhance.npz"))



def _cv2():
    """OpenCV is only needed by the live comparisons; the golden-vector and plan tests run without it."""
    return pytest.importorskip("cv2")


def _images(rng, n, lo=3, hi=160):
    for t in range(n):
        h, w = int(rng.integers(lo, hi)), int(rng.integers(lo, hi * 2))
        if t % 3 == 0:
            yield rng.integers(0, 256, (h, w), dtype=np.uint8)
        elif t % 3 == 1:
            yield np.clip(np.where(rng.random((h, w)) < 0.3, rng.normal(60, 20, (h, w)),
                                   rng.normal(190, 25, (h, w))), 0, 255).astype(np.uint8)
        else:
            yield (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)


def test_golden_chains():
    n = sum(1 for k in GOLD.files if k.startswith("crop_"))
    assert n >= 6
    for i in range(n):
        rgb = GOLD[f"crop_{i}"]
        assert np.array_equal(oe.enhance_for_ocrspace(rgb, "text"), GOLD[f"text_{i}"])
        assert np.array_equal(oe.enhance_for_ocrspace(rgb, "amount"), GOLD[f"amount_{i}"])
        assert np.array_equal(oe.enhance_for_date_ocr(rgb), GOLD[f"date_{i}"])


def test_golden_inputs_are_the_seeded_crops():
    import json
    note = json.loads(str(GOLD["note"]))
    crops = synthetic_crops_u8([tuple(s) for s in note["sizes"]], seed=note["seed"])
    for i, c in enumerate(crops):
        assert np.array_equal(c, GOLD[f"crop_{i}"])


def test_gray_sharpen_blur_clahe_otsu_against_live_opencv():
    cv2 = _cv2()
    rng = np.random.default_rng(5)
    rgb = rng.integers(0, 256, (67, 131, 3), dtype=np.uint8)
    assert np.array_equal(oe.rgb_to_gray(rgb), cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY))
    kernel = np.array([[-1, -1, -1], [-1, 9, -1], [-1, -1, -1]])
    for g in _images(rng, 18, lo=1):
        assert np.array_equal(oe.sharpen(g), cv2.filter2D(g, -1, kernel))
        assert np.array_equal(oe.gaussian_blur3(g), cv2.GaussianBlur(g, (3, 3), 0))
        t, b = cv2.threshold(g, 0, 255, cv2.THRESH_OTSU)
        assert oe.otsu_threshold(g) == int(t)
    for t, g in enumerate(_images(rng, 12, lo=8)):
        if t % 4 == 0:
            g = g[:g.shape[0] - g.shape[0] % 8 or 8, :g.shape[1] - g.shape[1] % 8 or 8]
        for clip in (4.0, 3.0):
            assert np.array_equal(oe.clahe(g, clip), cv2.createCLAHE(clipLimit=clip, tileGridSize=(8, 8)).apply(g))


def test_clahe_tiny_images():
    """height/width below the 8x8 grid: the reflect-101 extension wraps more than once."""
    cv2 = _cv2()
    rng = np.random.default_rng(6)
    for h, w in [(4, 4), (4, 36), (12, 4), (8, 8), (4, 8)]:
        g = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(oe.clahe(g, 4.0), cv2.createCLAHE(clipLimit=4.0, tileGridSize=(8, 8)).apply(g)), (h, w)


_CHILD = r"""
import sys, numpy as np, cv2
rng = np.random.default_rng(int(sys.argv[2]))
out = {}
for t in range(int(sys.argv[3])):
    h, w = int(rng.integers(1, 90)), int(rng.integers(1, 150))
    g = rng.integers(0, 256, (h, w), dtype=np.uint8) if t % 2 else np.clip(rng.normal(200, 40, (h, w)), 0, 255).astype(np.uint8)
    out[f"g{t}"] = g
    out[f"r{t}"] = cv2.resize(g, None, fx=4, fy=4, interpolation=cv2.INTER_CUBIC)
np.savez(sys.argv[1], **out)
"""


def test_resize_against_opencv_native_path(tmp_path):
    """cv2.resize with Intel IPP disabled (OpenCV's own HResizeCubic / VResizeCubicVec path)."""
    _cv2()
    path = str(tmp_path / "r.npz")
    env = dict(os.environ, OPENCV_IPP="disabled")
    subprocess.run([sys.executable, "-c", _CHILD, path, "9", "40"], env=env, check=True, stderr=subprocess.DEVNULL)
    z = np.load(path)
    for t in range(40):
        assert np.array_equal(oe.resize_cubic_x4(z[f"g{t}"]), z[f"r{t}"]), z[f"g{t}"].shape


def test_resize_against_installed_opencv_is_within_one():
    """Whatever cv2.resize the installed wheel uses (IPP here): at most +-1 on a few ppm of pixels."""
    cv2 = _cv2()
    rng = np.random.default_rng(10)
    bad = tot = 0
    for g in _images(rng, 12, lo=2, hi=100):
        r, o = cv2.resize(g, None, fx=4, fy=4, interpolation=cv2.INTER_CUBIC), oe.resize_cubic_x4(g)
        d = np.abs(r.astype(int) - o.astype(int))
        assert d.max() <= 1
        bad += int((d != 0).sum())
        tot += d.size
    assert bad <= 1e-4 * tot


def test_final_images_against_stock_ipp_opencv_are_bounded():
    """app_camera.py:572-598 / 685-705 run with the installed (IPP-enabled) cv2 against the oracle chain:
    the post-CLAHE / post-Otsu disagreement stays inside IPP_BOUNDS.  With OPENCV_IPP=disabled it is zero
    (test_golden_chains, test_resize_against_opencv_native_path)."""
    _cv2()
    rng = np.random.default_rng(5)
    sizes = [(int(rng.integers(8, 72)), int(rng.integers(20, 320))) for _ in range(24)]
    crops = synthetic_crops_u8(sizes, seed=77)
    for kind, (overall, worst, maxdiff) in IPP_BOUNDS.items():
        bad = tot = 0
        for c in crops:
            want = _stock_cv2_chain(c, kind)
            got = oe.enhance_for_date_ocr(c) if kind == "date" else oe.enhance_for_ocrspace(c, kind)
            d = np.abs(want.astype(int) - got.astype(int))
            assert d.max() <= maxdiff, (kind, c.shape, int(d.max()))
            assert (d != 0).mean() <= worst, (kind, c.shape, float((d != 0).mean()))
            bad += int((d != 0).sum())
            tot += d.size
        assert bad <= overall * tot, (kind, bad, tot)


def test_c_abi_plan_matches_oracle_geometry():
    """unetb200_enhance_plan is host-only: CLAHE tile sizes / clip limits equal the oracle's, buffers are
    packed at 16-byte aligned, non-overlapping offsets, output blocks are numbered consecutively."""
    from tw_invoice_unet_ocr_llm_b200 import enhance
    rng = np.random.default_rng(7)
    sizes = [(int(rng.integers(1, 300)), int(rng.integers(1, 500))) for _ in range(40)] + [(1, 1), (2, 2), (8, 8)]
    kinds = [("text", "amount", "date")[i % 3] for i in range(len(sizes))]
    table, sb, ob, wb = enhance.plan(sizes, kinds)
    blocks = so = oo = wo = 0
    for (h, w), k, t in zip(sizes, kinds, table):
        flags, clip = enhance.KINDS[k]
        assert (flags, clip) == oe.MODES[k]
        _, _, th, tw = oe.clahe_geometry(4 * h, 4 * w)
        assert (t.tile_h, t.tile_w) == (th, tw)
        assert t.clip_count == oe.clahe_clip_limit(clip, th * tw)
        assert t.first_block == blocks and t.blocks_x == -(-4 * w // 32) and t.n_blocks == t.blocks_x * -(-4 * h // 32)
        blocks += t.n_blocks
        assert (t.src_off, t.out_off, t.ws_off) == (so, oo, wo)
        assert t.src_off % 16 == 0 and t.out_off % 16 == 0 and t.ws_off % 16 == 0
        so += -(-3 * h * w // 16) * 16
        oo += -(-16 * h * w // 16) * 16
        wo += -(-16 * h * w // 16) * 16 + 64 * 256 + 1024 + 16
    assert (sb, ob, wb) == (so, oo, wo)


def test_c_abi_plan_rejects_bad_crops():
    from tw_invoice_unet_ocr_llm_b200 import _native as nat, enhance
    for sizes, kinds in [([(0, 4)], ["text"]), ([(4, -1)], ["text"]), ([(9000, 4)], ["text"]),
                         ([(4, 4)], [(8, 4.0)]), ([(4, 4)], [(1, 0.0)])]:
        with pytest.raises(nat.UnetB200Error):
            enhance.plan(sizes, kinds)
    with pytest.raises(ValueError):
        enhance.plan([(4, 4)], ["bogus"])


def test_c_abi_plan_windows_of_a_frame():
    """Crops given as windows of a larger frame (src_stride != 0) keep their src_off, take no room in the
    packed source buffer, and may use 3- or 4-byte pixels; packed crops are always RGB."""
    import ctypes as C
    from tw_invoice_unet_ocr_llm_b200 import _native as nat
    lib = nat.lib()
    t = (nat.EnhCrop * 3)()
    t[0].h, t[0].w, t[0].flags, t[0].clip = 10, 20, 5, 4.0                                   # packed
    t[1].h, t[1].w, t[1].flags, t[1].clip, t[1].src_stride, t[1].src_pixel_bytes, t[1].src_off = 7, 9, 6, 3.0, 640, 4, 12345 * 4
    t[2].h, t[2].w, t[2].flags, t[2].clip, t[2].src_stride, t[2].src_off = 5, 5, 1, 4.0, 100, 999   # 3-byte default
    sb, ob, wb = C.c_uint64(), C.c_uint64(), C.c_uint64()
    nat.check(lib.unetb200_enhance_plan(t, 3, C.byref(sb), C.byref(ob), C.byref(wb)))
    assert sb.value == 608                               # only the packed crop: 10*20*3 = 600 -> 16-byte aligned
    assert (t[0].src_stride, t[0].src_pixel_bytes, t[0].src_off) == (20, 3, 0)
    assert (t[1].src_stride, t[1].src_pixel_bytes, t[1].src_off) == (640, 4, 12345 * 4)
    assert (t[2].src_stride, t[2].src_pixel_bytes, t[2].src_off) == (100, 3, 999)
    assert ob.value == 16 * (200 + 63 + 25)              # 16*h*w is always a multiple of 16
    for bad in ({"src_stride": 8}, {"src_pixel_bytes": 5}, {"src_pixel_bytes": 4}):      # stride < w, bad size, packed RGBX
        b = (nat.EnhCrop * 1)()
        b[0].h, b[0].w, b[0].flags, b[0].clip = 10, 20, 5, 4.0
        for k, v in bad.items():
            setattr(b[0], k, v)
        assert lib.unetb200_enhance_plan(b, 1, C.byref(sb), C.byref(ob), C.byref(wb)) == nat.EINVAL, bad
    # run() refuses a table that did not go through plan(), without touching the GPU
    raw = (nat.EnhCrop * 1)()
    raw[0].h, raw[0].w, raw[0].flags, raw[0].clip = 4, 4, 5, 4.0
    assert lib.unetb200_enhance_run(raw, C.c_void_p(16), 1, C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), None) == nat.EINVAL
