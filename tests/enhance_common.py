"""Shared by the CPU and GPU enhancement tests: the reference's cv2 calls and the stated IPP bounds."""
import numpy as np


def stock_cv2_chain(rgb, kind):
    """The reference's own OpenCV calls (app_camera.py:581-598 / :689-703) with whatever cv2 is installed."""
    import cv2
    gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
    gray = cv2.resize(gray, None, fx=4, fy=4, interpolation=cv2.INTER_CUBIC)
    if kind == "date":
        gray = cv2.createCLAHE(clipLimit=3.0, tileGridSize=(8, 8)).apply(gray)
        gray = cv2.GaussianBlur(gray, (3, 3), 0)
        return cv2.threshold(gray, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)[1]
    gray = cv2.filter2D(gray, -1, np.array([[-1, -1, -1], [-1, 9, -1], [-1, -1, -1]]))
    enhanced = cv2.createCLAHE(clipLimit=4.0, tileGridSize=(8, 8)).apply(gray)
    return cv2.threshold(enhanced, 0, 255, cv2.THRESH_OTSU)[1] if kind == "text" else enhanced


# Bounds on the FINAL images (after sharpen / CLAHE / Otsu) between OpenCV's own code path -- what the oracle and
# the CUDA kernels reproduce bit for bit -- and the stock wheel's IPP-routed cv2.resize (the reference pins
# opencv-python-headless==4.8.1.78, requirements.txt:4, a stock wheel with IPP enabled).  IPP's cubic differs by
# +-1 on a few ppm of the upscaled pixels; the 3x3 sharpen multiplies that by up to 9, the CLAHE LUT by its slope,
# and Otsu turns a crossing into 0 <-> 255.  Measured here (60 crops, 5.8 Mpixel, cv2 4.13.0 + IPP): text 175 ppm
# of the binary pixels flipped (worst crop 0.63 %), amount 1150 ppm of the gray pixels differ (max 38 levels,
# worst crop 2.6 %), date 0.2 ppm.
IPP_BOUNDS = {            # kind -> (overall fraction, worst single crop, max gray-level difference)
    "text": (6e-4, 0.02, 255),
    "amount": (4e-3, 0.08, 96),
    "date": (5e-5, 0.005, 255),
}
