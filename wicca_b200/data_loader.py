"""GPU ingest for the image files the reference reads with OpenCV (row N2 of the hot-path table).

Mirror of ``wicca.data_loader.load_image`` (wicca/data_loader.py:27-63): ``cv2.imread(file_path)`` followed by
``cv2.cvtColor(image, cv2.COLOR_BGR2RGB)``.  For baseline (single-scan Huffman) JPEG files the whole decoder - entropy
decoding, dequantisation, inverse DCT, chroma upsampling, YCbCr -> RGB, EXIF orientation - runs on the GPU with
libjpeg-turbo's default arithmetic, so the returned array is bit-identical to the reference's; the host only parses
markers and strips the byte stuffing.  Files outside that subset (progressive JPEG, PNG, ...) raise
:class:`UnsupportedImageError` - there is no CPU fallback in this package; read those with ``cv2.imread``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

from . import _capi
from ._capi import UnsupportedImageError

__all__ = ["load_image", "decode_jpeg", "jpeg_info", "icons_from_jpeg", "icons_from_jpeg_files", "UnsupportedImageError"]


def _default_device() -> int:
    return int(os.environ.get("WICCA_B200_DEVICE", "0"))


def jpeg_info(data: bytes) -> dict:
    """Geometry of a JPEG stream the GPU path can decode: ``{"height", "width", "components", "h_max", "v_max"}``
    (``h_max`` x ``v_max`` is the chroma subsampling: 2 x 2 for 4:2:0).  Raises :class:`UnsupportedImageError` for
    anything else; needs no GPU."""
    lib = _capi.load()
    h, w, n, hm, vm = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _capi.check(lib.wicca_jpeg_probe(data, len(data), C.byref(h), C.byref(w), C.byref(n), C.byref(hm), C.byref(vm)),
                "wicca_jpeg_probe")
    return {"height": h.value, "width": w.value, "components": n.value, "h_max": hm.value, "v_max": vm.value}


def decode_jpeg(data: bytes, device: int | None = None, timing: dict | None = None) -> np.ndarray:
    """``cv2.cvtColor(cv2.imdecode(data, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)`` for a baseline JPEG: an
    ``(H, W, 3)`` uint8 RGB array (grey files come back with three equal channels, as IMREAD_COLOR does)."""
    info = jpeg_info(data)
    out = np.empty((info["height"], info["width"], 3), dtype=np.uint8)
    lib = _capi.load()
    t = _capi.Timing()
    host_ms = C.c_float()
    rc = lib.wicca_jpeg_decode_u8(data, len(data), out.ctypes.data, 0, _default_device() if device is None else int(device),
                                  C.byref(t), C.byref(host_ms))
    _capi.check(rc, "wicca_jpeg_decode_u8")
    if timing is not None:
        timing.update(t.as_dict(), host_decode_ms=float(host_ms.value))
    return out


def load_image(file_path: str, device: int | None = None) -> np.ndarray:
    """Same contract as the reference's ``load_image`` for the files this path decodes: RGB ``(H, W, 3)`` uint8.
    ``ValueError`` for an empty path (data_loader.py:47-48); unreadable files raise ``OSError`` and unsupported
    formats :class:`UnsupportedImageError` instead of the reference's print-and-return-``None``."""
    if not file_path:
        raise ValueError("File path cannot be empty")
    with open(file_path, "rb") as fh:
        data = fh.read()
    return decode_jpeg(data, device)


def icons_from_jpeg(data: bytes, transform_depths: Sequence[int], border_type: int = 1, border_constant: int = 0,
                    device: int | None = None, timing: dict | None = None) -> list:
    """``[get_small_copy(load_image(...), d) for d in transform_depths]`` straight from the JPEG bytes: the RGB image
    only ever exists in GPU memory.  Depths must be >= 1."""
    info = jpeg_info(data)
    depths = [int(d) for d in transform_depths]
    lib = _capi.load()
    outs = [np.empty((lib.wicca_icon_dim(info["height"], d), lib.wicca_icon_dim(info["width"], d), 3), dtype=np.uint8)
            for d in depths]
    t = _capi.Timing()
    host_ms = C.c_float()
    rc = lib.wicca_jpeg_icons_multi_u8(data, len(data), (C.c_int * len(depths))(*depths), len(depths), int(border_type),
                                       float(border_constant), (C.c_void_p * len(depths))(*[o.ctypes.data for o in outs]),
                                       _default_device() if device is None else int(device), C.byref(t), C.byref(host_ms))
    _capi.check(rc, "wicca_jpeg_icons_multi_u8")
    if timing is not None:
        timing.update(t.as_dict(), host_decode_ms=float(host_ms.value))
    return outs


def icons_from_jpeg_files(paths: Sequence[str], transform_depths: Sequence[int], border_type: int = 1,
                          border_constant: int = 0, devices: Sequence[int] | None = None, threads: int = 0,
                          timing: dict | None = None) -> list:
    """The reference's per-image loop (classifying_tools.py:312-321) from file paths: one list of icons per file.
    Files are Huffman-decoded by ``threads`` host threads (0 = one per core) and spread over ``devices``."""
    blobs = []
    for p in paths:
        if not p:
            raise ValueError("File path cannot be empty")
        with open(p, "rb") as fh:
            blobs.append(fh.read())
    depths = [int(d) for d in transform_depths]
    lib = _capi.load()
    n, nd = len(blobs), len(depths)
    if n == 0:
        return []
    infos = [jpeg_info(b) for b in blobs]
    outs = [[np.empty((lib.wicca_icon_dim(i["height"], d), lib.wicca_icon_dim(i["width"], d), 3), dtype=np.uint8) for d in depths]
            for i in infos]
    if devices is None:
        devices = list(range(max(1, lib.wicca_device_count())))
    datas = (C.c_void_p * n)(*[C.cast(C.c_char_p(b), C.c_void_p).value for b in blobs])
    lens = (C.c_size_t * n)(*[len(b) for b in blobs])
    dsts = (C.c_void_p * (n * nd))(*[o.ctypes.data for per in outs for o in per])
    host_ms = C.c_float()
    rc = lib.wicca_batch_icons_from_jpeg(datas, lens, n, (C.c_int * nd)(*depths), nd, int(border_type), float(border_constant),
                                         dsts, (C.c_int * len(devices))(*[int(x) for x in devices]), len(devices), int(threads),
                                         C.byref(host_ms))
    _capi.check(rc, "wicca_batch_icons_from_jpeg")
    if timing is not None:
        timing["host_decode_ms"] = float(host_ms.value)
    return outs
