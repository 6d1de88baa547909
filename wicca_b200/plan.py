"""Device-resident batches: thin wrapper over the ``wicca_plan_*`` C entry points.

A plan describes N images that already live in HBM (pitched uint8 HWC, see
``wicca_pitch_bytes``) plus the depths wanted; ``launch()`` produces every icon
of every image with one kernel launch (plus one tiny border pre-pass).  Used by
``bench.py`` for the device-resident number and by the torch bridge.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _capi


class IconPlan:
    def __init__(self, device: int, d_srcs: Sequence[int], hs: Sequence[int], ws: Sequence[int],
                 pitches: Sequence[int], depths: Sequence[int], channels: int = 3,
                 border_type: int = 1, border_constant: float = 0.0):
        self._lib = _capi.load()
        n, nd = len(d_srcs), len(depths)
        self.n, self.depths, self.channels, self.device = n, [int(d) for d in depths], channels, device
        self._h = C.c_void_p()
        rc = self._lib.wicca_plan_create(
            int(device), n, (C.c_void_p * n)(*d_srcs), (C.c_int * n)(*hs), (C.c_int * n)(*ws),
            (C.c_int64 * n)(*pitches), int(channels), (C.c_int * nd)(*self.depths), nd, int(border_type),
            float(border_constant), C.byref(self._h))
        _capi.check(rc, "wicca_plan_create")

    def launch(self, stream: int = 0) -> None:
        """Enqueue the whole batch on ``stream`` (a ``cudaStream_t`` handle; 0 = default stream)."""
        _capi.check(self._lib.wicca_plan_launch(self._h, C.c_void_p(stream)), "wicca_plan_launch")

    def icon_info(self, image: int, depth_index: int) -> tuple[int, int, int, int]:
        """(device pointer, h, w, pitch) of one icon."""
        p, h, w, pitch = C.c_void_p(), C.c_int(), C.c_int(), C.c_int64()
        _capi.check(self._lib.wicca_plan_icon(self._h, image, depth_index, C.byref(p), C.byref(h), C.byref(w),
                                              C.byref(pitch)), "wicca_plan_icon")
        return int(p.value), h.value, w.value, pitch.value

    def read_icon(self, image: int, depth_index: int) -> np.ndarray:
        """Synchronous copy of one icon to a tight host array."""
        _, h, w, _ = self.icon_info(image, depth_index)
        out = np.empty((h, w, self.channels), np.uint8)
        _capi.check(self._lib.wicca_plan_read_icon(self._h, image, depth_index, out.ctypes.data), "wicca_plan_read_icon")
        return out

    def resize_norm(self, depth_index: int, out_h: int, out_w: int, norm_mode: int, d_dst: int, d_dst_u8: int = 0,
                    stream: int = 0) -> None:
        """Fused epilogue on the plan's icons (device pointers in, nothing leaves the GPU)."""
        _capi.check(self._lib.wicca_plan_resize_norm(self._h, depth_index, out_h, out_w, norm_mode, C.c_void_p(d_dst),
                                                     C.c_void_p(d_dst_u8) if d_dst_u8 else None, C.c_void_p(stream)),
                    "wicca_plan_resize_norm")

    def info(self) -> dict:
        launches, br, bw = C.c_int(), C.c_int64(), C.c_int64()
        _capi.check(self._lib.wicca_plan_info(self._h, C.byref(launches), C.byref(br), C.byref(bw)), "wicca_plan_info")
        return {"launches": launches.value, "bytes_read": br.value, "bytes_written": bw.value}

    def close(self) -> None:
        if self._h:
            self._lib.wicca_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # noqa: D105
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def pitch_bytes(width: int, channels: int = 3) -> int:
    return int(_capi.load().wicca_pitch_bytes(int(width), int(channels)))


def to_device_pitched(image: np.ndarray, device: int = 0):
    """Upload a host ``(H, W, C)`` uint8 image into a pitched torch uint8 tensor ``(H, pitch)``
    (PyTorch is used only as the device allocator).  Returns the tensor; its ``data_ptr()`` is
    what the C ABI takes."""
    import torch  # noqa: PLC0415

    h, w, c = image.shape
    pitch = pitch_bytes(w, c)
    t = torch.zeros((h, pitch), dtype=torch.uint8, device=f"cuda:{device}")
    t[:, : w * c].copy_(torch.from_numpy(np.ascontiguousarray(image).reshape(h, w * c)), non_blocking=False)
    return t
