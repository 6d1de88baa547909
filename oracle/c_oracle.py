"""ctypes loader of the C restatement (oracle/haar_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from .haar_oracle import saturate_u8

HERE = Path(__file__).resolve().parent
_lib = None


def load():
    global _lib
    if _lib is None:
        so = HERE / "_build" / "liboracle_haar.so"
        src = HERE / "haar_oracle.c"
        if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
            subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
        _lib = C.CDLL(str(so))
        _lib.oracle_haar_icon_u8.restype = C.c_int
        _lib.oracle_haar_icon_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p]
        _lib.oracle_border_index.restype = C.c_int
        _lib.oracle_border_index.argtypes = [C.c_int, C.c_int, C.c_int]
    return _lib


def haar_icon(image: np.ndarray, depth: int, border_type: int = 1, border_constant=0) -> np.ndarray:
    """C-speed oracle for (H, W, C) uint8 images; same contract as haar_oracle.haar_icon_fp32."""
    img = np.ascontiguousarray(image)
    h, w, c = img.shape
    oh, ow = (h, w) if depth <= 0 else (-(-h // (1 << depth)), -(-w // (1 << depth)))
    out = np.empty((oh, ow, c), np.uint8)
    rc = load().oracle_haar_icon_u8(img.ctypes.data, h, w, c, w * c, int(depth), int(border_type),
                                    saturate_u8(border_constant), out.ctypes.data)
    if rc != 0:
        raise MemoryError("oracle_haar_icon_u8 failed")
    return out
