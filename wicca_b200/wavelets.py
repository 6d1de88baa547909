"""Other orthogonal wavelets behind the reference's ``WaveletCoder`` interface (row N4 of the hot-path table).

The reference implements only ``HaarCoder`` and lists Daubechies / Coiflet coders as its roadmap (README.md:25,
:222; abstract interface ``wicca/wavelet_coder.py:26-38``).  These coders keep the interface - ``get_small_copy(image,
transform_depth, border_type, border_constant)`` with the reference's defaults, return type and exceptions - and
replace the 2 x 2 mean by a longer low-pass filter: pad bottom/right to a multiple of ``2**depth`` exactly as
``get_padded_copy`` does, then ``depth`` levels of separable low-pass filtering (taps ``dec_lo / sqrt 2``, periodic
wrap-around) with decimation by 2, in float32, clipped and truncated to uint8.  There is no reference implementation
to be at parity with for these filters; ``OrthogonalWaveletCoder("haar")`` reproduces ``HaarCoder`` bit for bit.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading

import numpy as np

from . import _capi
from .wavelet_coder import (BORDER_REPLICATE, WaveletCoder, _as_depth, _check_layout, _row_major_view, validate_image)

__all__ = ["OrthogonalWaveletCoder", "DaubechiesCoder", "CoifletCoder", "DEC_LO"]

# decomposition low-pass filters (orthonormal: sum = sqrt 2, unit norm), as tabulated by Daubechies / PyWavelets
DEC_LO = {
    "haar": (0.7071067811865476, 0.7071067811865476),
    "db2": (-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025),
    "db3": (0.035226291882100656, -0.08544127388224149, -0.13501102001039084, 0.4598775021193313, 0.8068915093133388,
            0.3326705529509569),
    "db4": (-0.010597401784997278, 0.032883011666982945, 0.030841381835986965, -0.18703481171888114,
            -0.02798376941698385, 0.6308807679295904, 0.7148465705525415, 0.23037781330885523),
    "coif1": (-0.01565572813546454, -0.0727326195128539, 0.38486484686420286, 0.8525720202122554, 0.3378976624578092,
              -0.0727326195128539),
}
DEC_LO["db1"] = DEC_LO["haar"]


class OrthogonalWaveletCoder(WaveletCoder):
    """``get_small_copy`` with the low-pass filter of an orthogonal wavelet: ``"haar"``/``"db1"``, ``"db2"``, ``"db3"``,
    ``"db4"``, ``"coif1"``, or any even-length sequence of ``dec_lo`` coefficients (at most 16)."""

    def __init__(self, wavelet="db2", device: int | None = None) -> None:
        if isinstance(wavelet, str):
            if wavelet not in DEC_LO:
                raise ValueError(f"unknown wavelet {wavelet!r}; expected one of {sorted(DEC_LO)} or a sequence of taps")
            dec_lo = DEC_LO[wavelet]
        else:
            dec_lo = tuple(float(x) for x in wavelet)
        if len(dec_lo) < 2 or len(dec_lo) > 16 or len(dec_lo) % 2:
            raise ValueError("need an even number of taps between 2 and 16")
        self.taps = (np.asarray(dec_lo, dtype=np.float64) / math.sqrt(2.0)).astype(np.float32)
        self.device = int(os.environ.get("WICCA_B200_DEVICE", "0")) if device is None else int(device)
        self._tls = threading.local()

    def get_small_copy(self, image: np.ndarray, transform_depth: int, border_type: int = BORDER_REPLICATE,
                       border_constant: int = 0) -> np.ndarray:
        validate_image(image)
        depth = _as_depth(transform_depth)
        _check_layout(image, (depth,), border_type)
        lib = _capi.load()
        squeeze2d = image.ndim == 2
        view, stride = _row_major_view(image)
        h, w, c = view.shape
        oh, ow = (h, w) if depth <= 0 else (-(-h // (1 << depth)), -(-w // (1 << depth)))
        out = np.empty((oh, ow, c), dtype=np.uint8)
        t = _capi.Timing()
        rc = lib.wicca_wavelet_icon_u8(view.ctypes.data, h, w, c, stride, depth, int(border_type), float(border_constant),
                                       self.taps.ctypes.data_as(C.POINTER(C.c_float)), len(self.taps), out.ctypes.data,
                                       self.device, C.byref(t))
        _capi.check(rc, "wicca_wavelet_icon_u8")
        self._tls.timing = t.as_dict()
        return out[:, :, 0] if squeeze2d else out

    @property
    def last_timing(self) -> dict | None:
        return getattr(self._tls, "timing", None)


class DaubechiesCoder(OrthogonalWaveletCoder):
    """Daubechies wavelet with ``order`` vanishing moments (1 = Haar, 2, 3, 4)."""

    def __init__(self, order: int = 2, device: int | None = None) -> None:
        if order not in (1, 2, 3, 4):
            raise ValueError("Daubechies orders 1-4 are tabulated")
        super().__init__(f"db{order}", device)


class CoifletCoder(OrthogonalWaveletCoder):
    """Coiflet of order 1 (6 taps)."""

    def __init__(self, order: int = 1, device: int | None = None) -> None:
        if order != 1:
            raise ValueError("Coiflet order 1 is tabulated")
        super().__init__("coif1", device)
