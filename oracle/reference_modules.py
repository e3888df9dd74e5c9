"""Loader for the UNMODIFIED reference modules (``unet_model.py``, ``inference.py``).

TEST INFRASTRUCTURE ONLY -- see ``oracle/unet_oracle.py``.  The modules are looked up in
``oracle/_ref/`` (written by ``oracle/make_ref.sh`` in the build container; git-ignored, but it
travels to the GPU box with the snapshot) and, failing that, in ``/root/reference`` (build
container only).  They are imported under private names so that they never shadow the
drop-in shims of the same names (``shims/unet_model.py``, ``shims/inference.py``).

Used by ``tests/`` (the oracle must equal the reference module bit for bit) and by
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs (the reference's own CPU forward,
``kind: "reference"``).
"""
from __future__ import annotations

import importlib.util
import os
import sys
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(HERE, "_ref"), "/root/reference")


def reference_dir() -> Optional[str]:
    """Directory holding the reference's ``unet_model.py`` and ``inference.py``, or ``None``."""
    for d in CANDIDATES:
        if os.path.isfile(os.path.join(d, "unet_model.py")) and os.path.isfile(os.path.join(d, "inference.py")):
            return d
    return None


def _load(path: str, private_name: str):
    spec = importlib.util.spec_from_file_location(private_name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_cache: dict = {}


def reference_unet_model():
    """The reference ``unet_model`` module (``UNet``, ``DoubleConv``), or ``None`` if unavailable."""
    if "unet_model" not in _cache:
        d = reference_dir()
        _cache["unet_model"] = _load(os.path.join(d, "unet_model.py"), "_reference_unet_model") if d else None
    return _cache["unet_model"]


def reference_inference():
    """The reference ``inference`` module (``load_model``, ``preprocess``, ``run_unet``), or ``None``.
    Its ``from unet_model import UNet`` is satisfied by the reference's own ``unet_model`` for the
    duration of the import only."""
    if "inference" not in _cache:
        um = reference_unet_model()
        if um is None:
            _cache["inference"] = None
        else:
            saved = sys.modules.get("unet_model")
            sys.modules["unet_model"] = um
            try:
                _cache["inference"] = _load(os.path.join(reference_dir(), "inference.py"), "_reference_inference")
            finally:
                if saved is None:
                    sys.modules.pop("unet_model", None)
                else:
                    sys.modules["unet_model"] = saved
    return _cache["inference"]
