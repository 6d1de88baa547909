"""NumPy restatement of ``cv2.resize(img_u8_hwc, (dw, dh), interpolation=cv2.INTER_AREA)``
and of Keras ``preprocess_input`` (CPU oracle for the epilogue rows A5/A6).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

Call sites restated: ``wicca/classifying_tools.py:318`` (icon -> classifier
input size; ``:315`` for the source image) and ``:286-287``
(``preprocess_input`` + cast to float32).

Both algorithms live in third-party dependencies absent from
``/root/reference``:

* OpenCV ``resize`` - ``opencv-python==4.12.0.88`` (``requirements.txt:91``).
  OpenCV 4.13.0 is installed in the build container, so the restatement is
  pinned against it (``tests/golden/make_golden.py`` writes cv2's outputs into
  ``tests/golden/resize_area_golden.npz``; ``tests/test_oracle_resize.py``
  re-checks live whenever cv2 is importable).
* Keras ``imagenet_utils.preprocess_input`` - ``keras==3.11.3``
  (``requirements.txt:58``), NOT installed and not vendored: PARITY UNPINNED.
  :func:`preprocess_input` restates the published
  ``_preprocess_numpy_input`` semantics.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


def _scales(ssize: int, dsize: int) -> tuple[float, float]:
    inv = dsize / ssize           # double, as OpenCV computes it
    return inv, 1.0 / inv         # scale MUST be 1.0/inv, not ssize/dsize


def regime(sw: int, sh: int, dw: int, dh: int) -> str:
    """Which OpenCV code path INTER_AREA takes for 8-bit input."""
    _, sx = _scales(sw, dw)
    _, sy = _scales(sh, dh)
    if sx >= 1 and sy >= 1:
        isx, isy = int(np.rint(sx)), int(np.rint(sy))   # saturate_cast<int>(double) == cvRound
        eps = np.finfo(np.float64).eps
        if abs(sx - isx) < eps and abs(sy - isy) < eps:
            return "fast"
        return "generic"
    return "bilinear"


def area_tab(ssize: int, dsize: int, scale: float):
    """``computeResizeAreaTab``: list of (dst_idx, src_idx, fp32 weight), in order."""
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1 = math.ceil(fsx1)
        sx2 = min(math.floor(fsx2), ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, F32((sx1 - fsx1) / cell)))
        for s in range(sx1, sx2):
            tab.append((dx, s, F32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, F32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def _rint_sat_u8(x: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def _resize_fast(src: np.ndarray, dw: int, dh: int, isx: int, isy: int) -> np.ndarray:
    sh, sw, c = src.shape
    blk = src[:dh * isy, :dw * isx].reshape(dh, isy, dw, isx, c).astype(np.int64)
    # OpenCV sums the area in the order given by its offset table (row-major
    # over the block); integer sums are order independent.
    s = blk.sum(axis=(1, 3))
    if isx == 2 and isy == 2:
        return ((s + 2) >> 2).astype(np.uint8)
    scale = F32(1.0 / (isx * isy))
    return _rint_sat_u8(s.astype(F32) * scale)


def _resize_generic(src: np.ndarray, dw: int, dh: int, sx: float, sy: float) -> np.ndarray:
    sh, sw, c = src.shape
    xtab = area_tab(sw, dw, sx)
    ytab = area_tab(sh, dh, sy)
    srcf = src.astype(F32)
    # horizontal pass for every source row: buf[sy, dx] += src[sy, sx] * alpha, taps in order
    hbuf = np.zeros((sh, dw, c), dtype=F32)
    for dx, s, a in xtab:
        hbuf[:, dx, :] = hbuf[:, dx, :] + srcf[:, s, :] * a
    out = np.zeros((dh, dw, c), dtype=F32)
    first = np.ones(dh, dtype=bool)
    for dy, s, b in ytab:
        if first[dy]:
            out[dy] = hbuf[s] * b
            first[dy] = False
        else:
            out[dy] = out[dy] + hbuf[s] * b
    return _rint_sat_u8(out)


def _bilinear_taps(ssize: int, dsize: int):
    inv, scale = _scales(ssize, dsize)
    idx0 = np.empty(dsize, np.int64)
    idx1 = np.empty(dsize, np.int64)
    c0 = np.empty(dsize, np.int64)
    c1 = np.empty(dsize, np.int64)
    for d in range(dsize):
        s = math.floor(d * scale)
        f = F32((d + 1) - (s + 1) * inv)
        f = F32(0) if f <= 0 else F32(f - F32(math.floor(f)))
        if s >= ssize - 1:
            s, f = ssize - 1, F32(0)
        idx0[d] = s
        idx1[d] = min(s + 1, ssize - 1)
        c0[d] = int(np.rint(F32(F32(1) - f) * F32(2048)))
        c1[d] = int(np.rint(F32(f * F32(2048))))
    return idx0, idx1, c0, c1


def _resize_bilinear_area(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw, c = src.shape
    x0, x1, cx0, cx1 = _bilinear_taps(sw, dw)
    y0, y1, cy0, cy1 = _bilinear_taps(sh, dh)
    s = src.astype(np.int64)
    rows = s[:, x0, :] * cx0[None, :, None] + s[:, x1, :] * cx1[None, :, None]   # int32 range, << 11
    r0 = rows[y0] >> 4
    r1 = rows[y1] >> 4
    out = (((cy0[:, None, None] * r0) >> 16) + ((cy1[:, None, None] * r1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def resize_area(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """uint8 HWC ``cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA)``."""
    assert src.dtype == np.uint8 and src.ndim == 3
    sh, sw, _ = src.shape
    if (sw, sh) == (dw, dh):
        return src.copy()
    _, sx = _scales(sw, dw)
    _, sy = _scales(sh, dh)
    reg = regime(sw, sh, dw, dh)
    if reg == "fast":
        return _resize_fast(src, dw, dh, int(np.rint(sx)), int(np.rint(sy)))
    if reg == "generic":
        return _resize_generic(src, dw, dh, sx, sy)
    return _resize_bilinear_area(src, dw, dh)


# ---------------------------------------------------------------------------
# Keras preprocess_input restatement (PARITY UNPINNED - keras not installed)
# ---------------------------------------------------------------------------
NORM_IDENTITY, NORM_TF, NORM_CAFFE, NORM_TORCH = 0, 1, 2, 3
NORM_MODES = {"identity": NORM_IDENTITY, "tf": NORM_TF, "caffe": NORM_CAFFE, "torch": NORM_TORCH}

_CAFFE_MEAN_BGR = np.array([103.939, 116.779, 123.68], dtype=F32)
_TORCH_MEAN = np.array([0.485, 0.456, 0.406], dtype=F32)
_TORCH_STD = np.array([0.229, 0.224, 0.225], dtype=F32)


def preprocess_input(batch_u8: np.ndarray, mode: str) -> np.ndarray:
    """``keras.applications.imagenet_utils.preprocess_input`` on a uint8
    ``(B,h,w,3)`` channels-last NumPy batch, then the float32 cast of
    ``wicca/classifying_tools.py:287``.

    ``_preprocess_numpy_input`` first does ``x = x.astype(floatx())`` (float32)
    for non-float input, then: "tf": ``x /= 127.5; x -= 1.``; "torch":
    ``x /= 255.`` then per channel ``x -= mean; x /= std``; "caffe": RGB->BGR
    then per channel ``x -= mean``; EfficientNet's ``preprocess_input`` is the
    identity.  All arithmetic in float32, true division.
    """
    x = batch_u8.astype(F32)
    if mode == "identity":
        return x
    if mode == "tf":
        x /= F32(127.5)
        x -= F32(1.0)
        return x
    if mode == "torch":
        x /= F32(255.0)
        x -= _TORCH_MEAN
        x /= _TORCH_STD
        return x
    if mode == "caffe":
        x = x[..., ::-1].copy()
        x -= _CAFFE_MEAN_BGR
        return x
    raise ValueError(f"unknown preprocess mode {mode!r}")
