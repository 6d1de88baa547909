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
        _lib.oracle_haar_icons_multi_u8.restype = C.c_int
        _lib.oracle_haar_icons_multi_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int,
                                                    C.c_int, C.c_void_p]
        _lib.oracle_haar_forward_f32.restype = C.c_int
        _lib.oracle_haar_forward_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.oracle_haar_inverse_f32.restype = C.c_int
        _lib.oracle_haar_inverse_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
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


def haar_icons_multi(image: np.ndarray, depths, border_type: int = 1, border_constant=0) -> list[np.ndarray]:
    """All requested depths (1..8) of one image from ONE pass of exact integer block sums (the identity
    of SURVEY.md 8(a) row A3) - ~20x faster than calling :func:`haar_icon` per depth."""
    img = np.ascontiguousarray(image)
    h, w, c = img.shape
    depths = [int(d) for d in depths]
    outs = [np.empty((-(-h // (1 << d)), -(-w // (1 << d)), c), np.uint8) for d in depths]
    rc = load().oracle_haar_icons_multi_u8(img.ctypes.data, h, w, c, w * c, (C.c_int * len(depths))(*depths), len(depths),
                                           int(border_type), saturate_u8(border_constant),
                                           (C.c_void_p * len(depths))(*[o.ctypes.data for o in outs]))
    if rc != 0:
        raise ValueError(f"oracle_haar_icons_multi_u8 failed ({rc})")
    return outs


def haar_forward_plane(image: np.ndarray, depth: int, border_type: int = 1, border_constant=0) -> np.ndarray:
    """Mallat-ordered float32 coefficient plane (Hp, Wp, C) - the C twin of haar_oracle.haar_forward."""
    img = np.ascontiguousarray(image)
    h, w, c = img.shape
    r = 1 << depth
    plane = np.empty((-(-h // r) * r, -(-w // r) * r, c), np.float32)
    rc = load().oracle_haar_forward_f32(img.ctypes.data, h, w, c, w * c, int(depth), int(border_type),
                                        saturate_u8(border_constant), plane.ctypes.data)
    if rc != 0:
        raise ValueError(f"oracle_haar_forward_f32 failed ({rc})")
    return plane


def haar_inverse_plane(plane: np.ndarray, depth: int) -> np.ndarray:
    pl = np.ascontiguousarray(plane, dtype=np.float32)
    hp, wp, c = pl.shape
    out = np.empty_like(pl)
    rc = load().oracle_haar_inverse_f32(pl.ctypes.data, hp, wp, c, int(depth), out.ctypes.data)
    if rc != 0:
        raise ValueError(f"oracle_haar_inverse_f32 failed ({rc})")
    return out
