"""The C-ABI library builds, loads and exports every symbol include/wicca_b200.h declares."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "wicca_b200.h").read_text()
    return sorted(set(re.findall(r"WICCA_API\s+[\w\s\*]+?\b(wicca_\w+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("wicca_haar_icon_u8", "wicca_haar_icons_multi_u8", "wicca_haar_icons_multi_dev", "wicca_plan_create",
                 "wicca_batch_icons_u8", "wicca_haar_forward_f32", "wicca_haar_inverse_f32",
                 "wicca_icon_resize_norm_f32", "wicca_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from wicca_b200 import _capi
    lib = _capi.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _capi.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_capi.SIGNATURES) == set(declared_symbols())


def test_pure_helpers_need_no_gpu():
    from wicca_b200 import _capi
    lib = _capi.load()
    assert lib.wicca_version().startswith(b"wicca_b200")
    assert lib.wicca_pitch_bytes(8284, 3) == 24960 and lib.wicca_pitch_bytes(4096, 3) == 12288
    assert [lib.wicca_icon_dim(6393, d) for d in range(0, 7)] == [6393, 3197, 1599, 800, 400, 200, 100]
    assert [lib.wicca_icon_dim(8284, d) for d in range(1, 7)] == [4142, 2071, 1036, 518, 259, 130]


def test_argument_errors_are_reported_before_device_work():
    from wicca_b200 import _capi
    lib = _capi.load()
    img = np.zeros((5, 7, 3), np.uint8)
    out = np.zeros((3, 4, 3), np.uint8)
    t = _capi.Timing()
    assert lib.wicca_haar_icon_u8(None, 5, 7, 3, 0, 1, 1, 0.0, out.ctypes.data, 0, ctypes.byref(t)) == _capi.EINVAL
    assert lib.wicca_haar_icon_u8(img.ctypes.data, 0, 7, 3, 0, 1, 1, 0.0, out.ctypes.data, 0, None) == _capi.EINVAL
    assert lib.wicca_haar_icon_u8(img.ctypes.data, 5, 7, 3, 0, 1, 7, 0.0, out.ctypes.data, 0, None) == _capi.EBORDER
    assert b"border" in lib.wicca_last_error()
    assert lib.wicca_haar_icon_u8(img.ctypes.data, 5, 7, 3, 0, 99, 1, 0.0, out.ctypes.data, 0, None) == _capi.EDEPTH
    assert lib.wicca_haar_icon_u8(img.ctypes.data, 5, 7, 3, 4, 1, 1, 0.0, out.ctypes.data, 0, None) == _capi.EINVAL  # stride < W*C
    # depth <= 0 is a plain copy and needs no device
    cp = np.empty_like(img)
    assert lib.wicca_haar_icon_u8(img.ctypes.data, 5, 7, 3, 0, 0, 1, 0.0, cp.ctypes.data, 0, None) == 0
