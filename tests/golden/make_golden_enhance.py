#!/usr/bin/env python
"""Golden vectors for the OCR crop enhancement -- BUILD container only.

``app_camera.py`` cannot be imported here (streamlit / easyocr / supabase are absent), so this script
parses the UNMODIFIED /root/reference/app_camera.py, compiles only the two function definitions
``enhance_for_ocrspace`` (:572-598) and ``enhance_for_date_ocr`` (:685-705) from it and executes
them on seeded synthetic crops.  It runs twice in subprocesses:

  OPENCV_IPP=disabled   OpenCV's own code path (what oracle/opencv_enhance.py restates)  -> golden
  default               this wheel routes cv2.resize through Intel IPP                   -> recorded
                        only as a mismatch count against the golden (+-1 before CLAHE on a few ppm)

    python tests/golden/make_golden_enhance.py
"""
import ast
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/app_camera.py"
SIZES = [(24, 70), (31, 45), (40, 121), (17, 18), (56, 200), (9, 33)]
SEED = 11


def reference_functions():
    import cv2
    from PIL import Image
    tree = ast.parse(open(REF, encoding="utf-8").read())
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef)
              and n.name in ("enhance_for_ocrspace", "enhance_for_date_ocr")]
    assert len(wanted) == 2
    ns = {"np": np, "cv2": cv2, "Image": Image}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), REF, "exec"), ns)
    return ns["enhance_for_ocrspace"], ns["enhance_for_date_ocr"]


def run_reference():
    from PIL import Image
    sys.path.insert(0, ROOT)
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_crops_u8
    ocrspace, date = reference_functions()
    out = {}
    for i, rgb in enumerate(synthetic_crops_u8(SIZES, seed=SEED)):
        pil = Image.fromarray(rgb)
        out[f"crop_{i}"] = rgb
        out[f"text_{i}"] = np.array(ocrspace(pil, mode="text"))
        out[f"amount_{i}"] = np.array(ocrspace(pil, mode="amount"))
        out[f"date_{i}"] = np.array(date(pil))
    assert ocrspace(None) is None and date(None) is None
    return out


def main():
    if len(sys.argv) > 1:                      # child: run and dump
        np.savez_compressed(sys.argv[1], **run_reference())
        return
    tmp = {k: os.path.join("/tmp", f"golden_enhance_{k}.npz") for k in ("native", "ipp")}
    for k, path in tmp.items():
        env = dict(os.environ)
        if k == "native":
            env["OPENCV_IPP"] = "disabled"
        else:
            env.pop("OPENCV_IPP", None)
        subprocess.run([sys.executable, os.path.abspath(__file__), path], env=env, check=True)
    native, ipp = np.load(tmp["native"]), np.load(tmp["ipp"])
    import cv2
    note = {"opencv": cv2.__version__, "sizes": SIZES, "seed": SEED, "ipp_vs_native_mismatching_pixels": {}}
    for k in native.files:
        if not k.startswith("crop_"):
            note["ipp_vs_native_mismatching_pixels"][k] = int((native[k] != ipp[k]).sum())
    rec = {k: native[k] for k in native.files}
    rec["note"] = np.array(json.dumps(note))
    np.savez_compressed(os.path.join(HERE, "golden_enhance.npz"), **rec)
    print(json.dumps(note, indent=1))


if __name__ == "__main__":
    main()
