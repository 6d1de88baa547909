"""Loader of the UNMODIFIED reference hot path staged under ``oracle/_ref`` (``make -C oracle ref``).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  ``oracle/_ref/wicca`` holds byte-for-byte
copies of ``/root/reference/wicca/{wavelet_coder,data_loader,validation,normalization}.py`` and
``config/`` (SHA-256 of each in ``oracle/_ref/SHA256SUMS``); the directory is git-ignored and is
rebuilt from the reference tree whenever ``__graft_entry__.build()`` runs where that tree exists.
``bench.py``'s ``cpu_baseline`` and ``--impl reference`` legs time ``HaarCoder.get_small_copy``
(``wicca/wavelet_coder.py:50-67``) from here; the NumPy port in ``haar_oracle.py`` is the
fallback only when this import fails.
"""
from __future__ import annotations

import importlib
import sys
from pathlib import Path

REF_DIR = Path(__file__).resolve().parent / "_ref"


def available() -> bool:
    return (REF_DIR / "wicca" / "wavelet_coder.py").exists()


def load_haar_coder():
    """Returns ``(HaarCoder class, kind)``: the staged reference (kind ``"reference"``)."""
    if not available():
        raise ImportError(f"{REF_DIR}/wicca is not staged (run `make -C oracle ref` where /root/reference exists)")
    clash = sys.modules.get("wicca")
    if clash is not None and not str(getattr(clash, "__file__", "")).startswith(str(REF_DIR)):
        raise ImportError("another `wicca` package is already imported")
    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    mod = importlib.import_module("wicca.wavelet_coder")
    return mod.HaarCoder, "reference"
