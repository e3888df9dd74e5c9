"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the OCR crop enhancement.

Restates, in numpy, what the reference's ``enhance_for_ocrspace`` (app_camera.py:572-598) and
``enhance_for_date_ocr`` (app_camera.py:685-705) compute.  The reference delegates every step to
OpenCV (third-party, pinned ``opencv-python-headless==4.8.1.78`` in the reference's requirements.txt:4,
absent from /root/reference; this image has 4.13.0),
so each function below restates the published OpenCV algorithm for 8-bit single-channel images:

=====================  ==========================================  ===============================
step                   reference call                              OpenCV routine restated
=====================  ==========================================  ===============================
``rgb_to_gray``        ``cv2.cvtColor(img, COLOR_RGB2GRAY)``       color_rgb: RGB2Gray<uchar>, 15-bit
``resize_cubic_x4``    ``cv2.resize(fx=4, fy=4, INTER_CUBIC)``     resize.cpp: HResizeCubic<uchar,int,short>
                                                                   + VResizeCubicVec_32s8u / VResizeCubic
``sharpen``            ``cv2.filter2D(gray, -1, [[-1..],[.9.]])``  filter.simd.hpp, BORDER_REFLECT_101
``clahe``              ``createCLAHE(clip, (8, 8)).apply``         clahe.cpp: CLAHE_CalcLut_Body,
                                                                   CLAHE_Interpolation_Body
``gaussian_blur3``     ``cv2.GaussianBlur(gray, (3, 3), 0)``       smooth.dispatch.cpp fixed-point 1-2-1
``otsu_threshold``     ``cv2.threshold(.., THRESH_OTSU)``          thresh.cpp: getThreshVal_Otsu_8u
=====================  ==========================================  ===============================

Pinning (tests/test_enhance_oracle.py): every function is compared bit for bit with OpenCV executed
in the test process; ``resize_cubic_x4`` and the two chains are compared with OpenCV's own code path
(``OPENCV_IPP=disabled``; the wheel in this image otherwise routes ``cv2.resize`` through Intel IPP,
whose cubic differs from OpenCV's by +-1 on ~5 ppm of the pixels), against golden vectors generated
by ``tests/golden/make_golden_enhance.py``.  Parity with the reference's pinned stock wheel (IPP enabled) is
therefore NOT bit-exact: ``test_final_images_against_stock_ipp_opencv_are_bounded`` bounds the difference
of the final (post-CLAHE / post-Otsu) images.

Nothing in the product path imports this module.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32

# cv::INTER_RESIZE_COEF_BITS
COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS
# lanes of one VResizeCubicVec_32s8u iteration in an SSE-baseline build (v_int16x8): the columns
# beyond the last full group of 8 take the integer tail loop (VResizeCubic + FixedPtCast)
VEC_LANES = 8

FLAG_SHARPEN, FLAG_BLUR, FLAG_OTSU = 1, 2, 4
MODES = {
    # app_camera.py:572-598
    "text": (FLAG_SHARPEN | FLAG_OTSU, 4.0),
    "amount": (FLAG_SHARPEN, 4.0),
    # app_camera.py:685-705
    "date": (FLAG_BLUR | FLAG_OTSU, 3.0),
}


def reflect101(p: np.ndarray, n: int) -> np.ndarray:
    """cv::borderInterpolate(p, n, BORDER_REFLECT_101) for an integer array."""
    p = np.asarray(p, dtype=np.int64).copy()
    if n == 1:
        return np.zeros_like(p)
    while True:
        lo, hi = p < 0, p >= n
        if not (lo.any() or hi.any()):
            return p
        p[lo] = -p[lo]
        p[hi] = 2 * (n - 1) - p[hi]


def rgb_to_gray(rgb: np.ndarray) -> np.ndarray:
    """RGB2Gray<uchar>: (R*9798 + G*19235 + B*3735 + 2^14) >> 15."""
    c = rgb.astype(np.int64)
    return ((c[..., 0] * 9798 + c[..., 1] * 19235 + c[..., 2] * 3735 + (1 << 14)) >> 15).astype(np.uint8)


def cubic_coeffs(x) -> np.ndarray:
    """cv::interpolateCubic (A = -0.75) in float32."""
    x, A, one = f32(x), f32(-0.75), f32(1)
    c0 = ((A * (x + one) - f32(5) * A) * (x + one) + f32(8) * A) * (x + one) - f32(4) * A
    c1 = ((A + f32(2)) * x - (A + f32(3))) * x * x + one
    c2 = ((A + f32(2)) * (one - x) - (A + f32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], dtype=np.float32)


def cubic_axis_tables(n_src: int, scale: int = 4):
    """Tap indices [n_dst, 4] (border-clamped) and 11-bit fixed-point taps [n_dst, 4] of one axis
    (resize.cpp: the xofs / ialpha loop; saturate_cast<short>(coeff * 2048) = round-half-even)."""
    n_dst = n_src * scale
    inv = 1.0 / scale
    idx = np.zeros((n_dst, 4), dtype=np.int64)
    taps = np.zeros((n_dst, 4), dtype=np.int64)
    for d in range(n_dst):
        fx = f32((d + 0.5) * inv - 0.5)
        s = int(np.floor(fx))
        fx = f32(fx - f32(s))
        taps[d] = np.rint(cubic_coeffs(fx) * f32(COEF_SCALE)).astype(np.int64)
        idx[d] = np.clip(np.arange(s - 1, s + 3), 0, n_src - 1)
    return idx, taps


def resize_cubic_x4(gray: np.ndarray) -> np.ndarray:
    """``cv2.resize(gray, None, fx=4, fy=4, interpolation=cv2.INTER_CUBIC)`` for uint8 [h, w]."""
    h, w = gray.shape
    dh, dw = 4 * h, 4 * w
    xi, xc = cubic_axis_tables(w)
    yi, yc = cubic_axis_tables(h)
    hor = (gray.astype(np.int64)[:, xi] * xc[None]).sum(-1)             # HResizeCubic: int rows [h, dw]
    rows = [hor[yi[:, k]] for k in range(4)]
    # integer tail: VResizeCubic with FixedPtCast<int, uchar, 22>
    acc = sum(rows[k] * yc[:, k][:, None] for k in range(4))
    out = np.clip((acc + (1 << (2 * COEF_BITS - 1))) >> (2 * COEF_BITS), 0, 255).astype(np.uint8)
    # vector body: float32, S0*b0 + (S1*b1 + (S2*b2 + S3*b3)), no fused multiply-add, v_round
    b = yc.astype(np.float32) * f32(1.0 / (COEF_SCALE * COEF_SCALE))
    r = [x.astype(np.float32) for x in rows]
    fl = r[0] * b[:, 0:1] + (r[1] * b[:, 1:2] + (r[2] * b[:, 2:3] + r[3] * b[:, 3:4]))
    nvec = dw - dw % VEC_LANES
    out[:, :nvec] = np.clip(np.rint(fl[:, :nvec]), 0, 255).astype(np.uint8)
    return out


def _window3(gray: np.ndarray):
    h, w = gray.shape
    ys = reflect101(np.arange(-1, h + 1), h)
    xs = reflect101(np.arange(-1, w + 1), w)
    return gray.astype(np.int64)[ys][:, xs]


def sharpen(gray: np.ndarray) -> np.ndarray:
    """``cv2.filter2D(gray, -1, [[-1,-1,-1],[-1,9,-1],[-1,-1,-1]])`` (BORDER_REFLECT_101, saturated)."""
    h, w = gray.shape
    p = _window3(gray)
    s = sum(p[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3))
    return np.clip(10 * gray.astype(np.int64) - s, 0, 255).astype(np.uint8)


def gaussian_blur3(gray: np.ndarray) -> np.ndarray:
    """``cv2.GaussianBlur(gray, (3, 3), 0)``: separable [1 2 1]/4 in 8.8 fixed point, one final
    round-half-up: (sum of the 16-weighted 3x3 window + 8) >> 4."""
    h, w = gray.shape
    p = _window3(gray)
    hor = p[:, 0:w] + 2 * p[:, 1:w + 1] + p[:, 2:w + 2]
    v = hor[0:h] + 2 * hor[1:h + 1] + hor[2:h + 2]
    return ((v + 8) >> 4).astype(np.uint8)


def clahe_geometry(h: int, w: int, tiles: int = 8):
    """(ext_h, ext_w, tile_h, tile_w): CLAHE_Impl::apply pads right/bottom with REFLECT_101 when
    either dimension is not a multiple of the grid (then BOTH get ``tiles - dim % tiles`` extra)."""
    if w % tiles == 0 and h % tiles == 0:
        eh, ew = h, w
    else:
        eh, ew = h + (tiles - h % tiles), w + (tiles - w % tiles)
    return eh, ew, eh // tiles, ew // tiles


def clahe_clip_limit(clip: float, tile_area: int) -> int:
    """``max((int)(clipLimit * tileSizeTotal / 256), 1)`` (double arithmetic), 0 = no clipping."""
    return max(int(clip * tile_area / 256), 1) if clip > 0 else 0


def clahe_luts(gray: np.ndarray, clip: float, tiles: int = 8) -> np.ndarray:
    """CLAHE_CalcLut_Body: uint8 [tiles, tiles, 256]."""
    h, w = gray.shape
    eh, ew, th, tw = clahe_geometry(h, w, tiles)
    ext = gray[reflect101(np.arange(eh), h)][:, reflect101(np.arange(ew), w)]
    area = th * tw
    lut_scale = f32(255) / f32(area)
    limit = clahe_clip_limit(clip, area)
    luts = np.zeros((tiles, tiles, 256), dtype=np.uint8)
    for ty in range(tiles):
        for tx in range(tiles):
            hist = np.bincount(ext[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw].ravel(), minlength=256).astype(np.int64)
            if limit > 0:
                clipped = int(np.maximum(hist - limit, 0).sum())
                hist = np.minimum(hist, limit)
                batch = clipped // 256
                resid = clipped - batch * 256
                hist += batch
                if resid:
                    step = max(256 // resid, 1)
                    i = 0
                    while i < 256 and resid > 0:
                        hist[i] += 1
                        i += step
                        resid -= 1
            luts[ty, tx] = np.clip(np.rint(np.cumsum(hist).astype(np.float32) * lut_scale), 0, 255).astype(np.uint8)
    return luts


def clahe(gray: np.ndarray, clip: float, tiles: int = 8) -> np.ndarray:
    """``cv2.createCLAHE(clipLimit=clip, tileGridSize=(8, 8)).apply(gray)`` for uint8 [h, w]."""
    h, w = gray.shape
    _, _, th, tw = clahe_geometry(h, w, tiles)
    lut = clahe_luts(gray, clip, tiles).astype(np.float32)

    def axis(n, tile):
        t = np.arange(n).astype(np.float32) * (f32(1) / f32(tile)) - f32(0.5)
        t1 = np.floor(t).astype(np.int64)
        a = (t - t1.astype(np.float32)).astype(np.float32)
        return np.maximum(t1, 0), np.minimum(t1 + 1, tiles - 1), a, (f32(1) - a).astype(np.float32)

    tx1, tx2, xa, xa1 = axis(w, tw)
    ty1, ty2, ya, ya1 = axis(h, th)
    v = gray.astype(np.int64)
    l11, l12 = lut[ty1[:, None], tx1[None, :], v], lut[ty1[:, None], tx2[None, :], v]
    l21, l22 = lut[ty2[:, None], tx1[None, :], v], lut[ty2[:, None], tx2[None, :], v]
    res = (l11 * xa1[None] + l12 * xa[None]) * ya1[:, None] + (l21 * xa1[None] + l22 * xa[None]) * ya[:, None]
    assert res.dtype == np.float32
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)


def otsu_threshold(gray: np.ndarray) -> int:
    """getThreshVal_Otsu_8u: double arithmetic, sequential over the 256 bins."""
    hist = np.bincount(gray.ravel(), minlength=256)
    scale = 1.0 / gray.size
    mu = 0.0
    for i in range(256):
        mu += i * float(hist[i])
    mu *= scale
    mu1 = q1 = max_sigma = 0.0
    max_val = 0
    eps = float(np.finfo(np.float32).eps)
    for i in range(256):
        p_i = float(hist[i]) * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < eps or max(q1, q2) > 1.0 - eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma = sigma
            max_val = i
    return max_val


def enhance(rgb: np.ndarray, flags: int, clip: float, stages: dict | None = None) -> np.ndarray:
    """The whole chain on a uint8 [h, w, 3] crop -> uint8 [4h, 4w]."""
    g = resize_cubic_x4(rgb_to_gray(rgb))
    if flags & FLAG_SHARPEN:
        g = sharpen(g)
    if stages is not None:
        stages["pre_clahe"] = g
    g = clahe(g, clip)
    if flags & FLAG_BLUR:
        g = gaussian_blur3(g)
    if stages is not None:
        stages["pre_threshold"] = g
    if flags & FLAG_OTSU:
        t = otsu_threshold(g)
        if stages is not None:
            stages["threshold"] = t
        g = np.where(g > t, 255, 0).astype(np.uint8)
    return g


def enhance_for_ocrspace(rgb: np.ndarray, mode: str = "text") -> np.ndarray:
    """app_camera.py:572-598 on the RGB array of the crop; any mode other than "text" skips Otsu."""
    flags, clip = MODES["text" if mode == "text" else "amount"]
    return enhance(rgb, flags, clip)


def enhance_for_date_ocr(rgb: np.ndarray) -> np.ndarray:
    """app_camera.py:685-705 on the RGB array of the crop."""
    flags, clip = MODES["date"]
    return enhance(rgb, flags, clip)
