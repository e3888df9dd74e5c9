"""In-tree build of ``libunetb200.so`` for sm_100a (``python -m tw_invoice_unet_ocr_llm_b200.build``)."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libunetb200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built")


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                  glob.glob(os.path.join(HERE, "..", "include", "*.h")))


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(s) <= t for s in sources())


def build(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    """Compile every CUDA source into the shared library; returns its path.  ``UNETB200_TEST_VARIANTS=1`` in the
    environment (or ``defines=["UNETB200_TEST_VARIANTS"]``) also compiles the measured-and-rejected kernel
    variants the cross-check tests can exercise (A_COL3 staging, the patch stem)."""
    defines = list(defines)
    if os.environ.get("UNETB200_TEST_VARIANTS") == "1" and "UNETB200_TEST_VARIANTS" not in defines:
        defines.append("UNETB200_TEST_VARIANTS")
    if out == OUT and not defines and not force and up_to_date():
        return OUT
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out + ".tmp",
           *sorted(glob.glob(os.path.join(CSRC, "*.cu")))]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    os.replace(out + ".tmp", out)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
