"""ctypes binding of ``libwicca_b200.so`` (see ``include/wicca_b200.h``).

This is the whole Python<->CUDA boundary: plain pointers and sizes, no torch
types.  Loading fails loudly when the library is missing and cannot be built;
there is no CPU fallback behind it.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

from . import _build

c_u8p = C.POINTER(C.c_uint8)
c_intp = C.POINTER(C.c_int)
c_i64p = C.POINTER(C.c_int64)
c_f32p = C.POINTER(C.c_float)


class Timing(C.Structure):
    """``wicca_timing``: per-stage device times in milliseconds (CUDA events)."""
    _fields_ = [("h2d_ms", C.c_float), ("kernel_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float)]

    def as_dict(self) -> dict:
        return {k: float(getattr(self, k)) for k, _ in self._fields_}


class Target(C.Structure):
    """``wicca_target``: one classifier input, (out_h, out_w) and the ``preprocess_input`` family (0..3)."""
    _fields_ = [("out_h", C.c_int), ("out_w", C.c_int), ("norm_mode", C.c_int)]


# name -> (restype, argtypes); every symbol declared in include/wicca_b200.h
SIGNATURES = {
    "wicca_version": (C.c_char_p, []),
    "wicca_last_error": (C.c_char_p, []),
    "wicca_device_count": (C.c_int, []),
    "wicca_shutdown": (C.c_int, []),
    "wicca_pitch_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "wicca_icon_dim": (C.c_int, [C.c_int, C.c_int]),
    "wicca_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "wicca_host_alloc_near": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t, C.c_int]),
    "wicca_host_free": (C.c_int, [C.c_void_p]),
    "wicca_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "wicca_host_unregister": (C.c_int, [C.c_void_p]),
    "wicca_haar_icon_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_double,
                                     C.c_void_p, C.c_int, C.POINTER(Timing)]),
    "wicca_haar_icons_multi_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, c_intp, C.c_int, C.c_int,
                                            C.c_double, C.POINTER(C.c_void_p), C.c_int, C.POINTER(Timing)]),
    "wicca_haar_icons_multi_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, c_intp, C.c_int, C.c_int,
                                             C.c_double, C.POINTER(C.c_void_p), c_i64p, C.c_int, C.c_void_p]),
    "wicca_plan_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p), c_intp, c_intp, c_i64p, C.c_int, c_intp,
                                    C.c_int, C.c_int, C.c_double, C.POINTER(C.c_void_p)]),
    "wicca_plan_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "wicca_plan_icon": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), c_intp, c_intp, c_i64p]),
    "wicca_plan_read_icon": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "wicca_plan_resize_norm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wicca_plan_info": (C.c_int, [C.c_void_p, c_intp, c_i64p, c_i64p]),
    "wicca_plan_destroy": (C.c_int, [C.c_void_p]),
    "wicca_batch_icons_u8": (C.c_int, [C.POINTER(C.c_void_p), c_intp, c_intp, c_i64p, C.c_int, C.c_int, c_intp, C.c_int,
                                       C.c_int, C.c_double, C.POINTER(C.c_void_p), c_intp, C.c_int, C.POINTER(Timing)]),
    "wicca_haar_forward_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_double,
                                         C.c_void_p, C.c_int, C.POINTER(Timing)]),
    "wicca_haar_inverse_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                         C.POINTER(Timing)]),
    "wicca_haar_forward_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_double,
                                         C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "wicca_haar_inverse_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p]),
    "wicca_batch_classifier_inputs_f32": (C.c_int, [C.POINTER(C.c_void_p), c_intp, c_intp, c_i64p, C.c_int, C.c_int, C.c_int,
                                                    C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, c_intp,
                                                    C.c_int, C.POINTER(Timing)]),
    "wicca_batch_classifier_inputs_multi_f32": (C.c_int, [C.POINTER(C.c_void_p), c_intp, c_intp, c_i64p, C.c_int, c_intp, C.c_int,
                                                          C.c_int, C.c_double, C.POINTER(Target), C.c_int,
                                                          C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), c_intp, C.c_int,
                                                          C.POINTER(Timing)]),
    "wicca_wavelet_icon_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_double,
                                        C.POINTER(C.c_float), C.c_int, C.c_void_p, C.c_int, C.POINTER(Timing)]),
    "wicca_jpeg_probe": (C.c_int, [C.c_void_p, C.c_size_t, c_intp, c_intp, c_intp, c_intp, c_intp]),
    "wicca_jpeg_coeff_count": (C.c_int64, [C.c_void_p, C.c_size_t]),
    "wicca_jpeg_decode_coeffs_gpu": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int64, C.c_int, c_intp]),
    "wicca_jpeg_decode_u8": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int64, C.c_int, C.POINTER(Timing),
                                       C.POINTER(C.c_float)]),
    "wicca_jpeg_decode_dev": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "wicca_jpeg_icons_multi_u8": (C.c_int, [C.c_void_p, C.c_size_t, c_intp, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_void_p),
                                            C.c_int, C.POINTER(Timing), C.POINTER(C.c_float)]),
    "wicca_batch_icons_from_jpeg": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int, c_intp, C.c_int, C.c_int,
                                              C.c_double, C.POINTER(C.c_void_p), c_intp, C.c_int, C.c_int,
                                              C.POINTER(C.c_float)]),
    "wicca_batch_classifier_inputs_multi_from_jpeg": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int, c_intp, C.c_int,
                                                                C.c_int, C.c_double, C.POINTER(Target), C.c_int,
                                                                C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), c_intp, C.c_int,
                                                                C.POINTER(Timing)]),
    "wicca_resize_norm_dev": (C.c_int, [C.POINTER(C.c_void_p), c_intp, c_intp, c_i64p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "wicca_icon_resize_norm_f32": (C.c_int, [C.POINTER(C.c_void_p), c_intp, c_intp, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Timing)]),
}

# error codes of include/wicca_b200.h
EINVAL, ECHANNELS, EBORDER, EDEPTH, EDEVICE, EALIGN, ENOMEM, ESTATE, EUNSUPPORTED = -1, -2, -3, -4, -5, -6, -7, -8, -9

_lib = None
_lock = threading.Lock()


class WiccaLibraryError(RuntimeError):
    """The native library is missing / cannot be loaded.  No CPU fallback exists."""


def library_path() -> Path:
    env = os.environ.get("WICCA_B200_LIB")
    return Path(env) if env else _build.lib_path()


def load(auto_build: bool = True):
    """Load (building first if stale and nvcc is available) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if auto_build and not os.environ.get("WICCA_B200_LIB") and _build.is_stale():
            try:
                _build.build()
            except Exception as exc:  # noqa: BLE001
                if not path.exists():
                    raise WiccaLibraryError(
                        f"{path} is missing and could not be built ({exc}); wicca_b200 has no CPU fallback") from exc
        if not path.exists():
            raise WiccaLibraryError(f"{path} not found; run `python -m wicca_b200._build` (no CPU fallback exists)")
        try:
            lib = C.CDLL(str(path))
        except OSError as exc:
            raise WiccaLibraryError(f"cannot load {path}: {exc}") from exc
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error() -> str:
    return load().wicca_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    """Map a C return code to the Python exception the reference would raise."""
    if rc == 0:
        return
    msg = last_error() or what
    if rc > 0:
        raise RuntimeError(f"{what}: {msg}" if what else msg)
    if rc == EBORDER:
        raise border_error_type()(msg)
    if rc == EUNSUPPORTED:
        raise UnsupportedImageError(msg)
    raise ValueError(msg)


class UnsupportedImageError(ValueError):
    """A valid image file outside the subset the GPU ingest path decodes (PNG, TIFF, arithmetic-coded JPEG ...).
    Nothing is decoded approximately and there is no CPU fallback: read such a file with ``cv2.imread``."""


def border_error_type():
    """cv2.copyMakeBorder raises ``cv2.error`` for an unknown border type
    (data_loader.py:116); use the same class when OpenCV is importable."""
    try:
        import cv2  # noqa: PLC0415
        return cv2.error
    except Exception:  # noqa: BLE001
        return ValueError
