#!/usr/bin/env python3
"""Generate the committed golden fixtures by running the LIVE reference.

Run in the build container only (needs ``/root/reference`` and cv2):

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

Writes next to this file:

* ``haar_icon_golden.npz`` - inputs + outputs of the reference's own
  ``wicca.wavelet_coder.HaarCoder().get_small_copy`` (``wavelet_coder.py:50-67``)
  over shapes x depths x border types, including the known-answer facts of
  SURVEY.md section 4 (truncation, all-0/all-255, depth<=0 identity).
* ``resize_area_golden.npz`` - ``cv2.resize(..., INTER_AREA)`` outputs (the
  call at ``classifying_tools.py:318``) for icons in all three OpenCV regimes.

Inputs are regenerated from seeds at test time (``np.random.default_rng``), so
only the seeds/parameters and the reference outputs are stored.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def gen_input(kind: str, seed: int, h: int, w: int, c: int) -> np.ndarray:
    """Deterministic inputs shared with the tests (tests import this)."""
    if kind == "noise":
        return np.random.default_rng(seed).integers(0, 256, (h, w, c), dtype=np.uint8)
    if kind == "zeros":
        return np.zeros((h, w, c), np.uint8)
    if kind == "full":
        return np.full((h, w, c), 255, np.uint8)
    if kind == "trunc":      # 255 with sparse 254s: sits on the truncation boundary
        a = np.full((h, w, c), 255, np.uint8)
        m = np.random.default_rng(seed).random((h, w, c)) < 0.05
        a[m] = 254
        return a
    if kind == "hramp":
        return np.broadcast_to((np.arange(w, dtype=np.int64) % 256).astype(np.uint8)[None, :, None], (h, w, c)).copy()
    if kind == "vramp":
        return np.broadcast_to((np.arange(h, dtype=np.int64) % 256).astype(np.uint8)[:, None, None], (h, w, c)).copy()
    raise ValueError(kind)


ICON_CASES = []
# (kind, seed, h, w, c, depth, border_type, border_constant)
_shapes = [(1, 1), (1, 2), (2, 1), (3, 5), (5, 7), (16, 16), (17, 33), (64, 64), (65, 63), (63, 129),
           (100, 130), (127, 255), (128, 256), (200, 259), (130, 517)]
_seed = 0
for (h, w) in _shapes:
    for d in (1, 2, 3, 4, 5, 6):
        for bt in (1, 0, 2, 3, 4):
            _seed += 1
            ICON_CASES.append(("noise", _seed, h, w, 3, d, bt, (_seed * 37) % 256))
for kind in ("zeros", "full", "trunc", "hramp", "vramp"):
    for d in (1, 3, 6, 8):
        _seed += 1
        ICON_CASES.append((kind, _seed, 75, 141, 3, d, 1, 0))
for c in (2, 4):
    for d in (1, 2, 5):
        for bt in (1, 2, 0):
            _seed += 1
            ICON_CASES.append(("noise", _seed, 45, 52, c, d, bt, 9))
for d in (0, -1, 7, 8, 9, 10):           # identity depths and the deep (fp32-rounding) ones
    _seed += 1
    ICON_CASES.append(("noise", _seed, 150, 260, 3, d, 1, 0))
ICON_CASES.append(("noise", 9001, 64, 128, 1, 3, 1, 0))     # C=1, no padding needed -> works in the reference
ICON_CASES.append(("noise", 9002, 64, 128, 5, 2, 1, 0))     # C=5, no padding needed
ICON_CASES.append(("noise", 9003, 50, 70, 3, 2, 17, 0))     # BORDER_ISOLATED flag is masked off
ICON_CASES.append(("noise", 9004, 50, 70, 3, 2, 0, 300))    # constant saturates to 255
ICON_CASES.append(("noise", 9005, 50, 70, 3, 2, 0, -5))     # constant saturates to 0

RESIZE_CASES = [
    # (seed, src_h, src_w, dst_w, dst_h)
    (1, 400, 518, 224, 224), (2, 400, 518, 331, 331),       # generic area
    (3, 200, 259, 224, 224), (4, 100, 130, 160, 160),       # bilinear "area mode"
    (5, 448, 448, 224, 224), (6, 336, 336, 112, 112),       # fast (2x2 and 3x3)
    (7, 300, 259, 112, 112), (8, 75, 141, 24, 33), (9, 33, 47, 40, 40),
]


# (seed, h, w, quality, sampling, restart interval, grey, EXIF orientation[, progressive])
JPEG_CASES = [
    (1, 17, 33, 90, "420", 0, 0, 1), (2, 37, 53, 75, "420", 3, 0, 1), (3, 70, 31, 95, "422", 0, 0, 1),
    (4, 96, 128, 85, "444", 0, 0, 1), (5, 64, 48, 35, "420", 0, 0, 1), (6, 33, 100, 100, "440", 0, 0, 1),
    (7, 50, 75, 90, "411", 5, 0, 1), (8, 40, 23, 85, "420", 0, 1, 1), (9, 61, 45, 90, "420", 0, 0, 6),
    (10, 45, 61, 80, "422", 0, 0, 3), (11, 1, 1, 90, "420", 0, 0, 1), (12, 2, 3, 90, "420", 0, 0, 1),
    (13, 128, 160, 92, "420", 0, 0, 8), (14, 99, 77, 60, "444", 7, 0, 5),
    (15, 70, 90, 90, "420", 0, 0, 1, 1), (16, 37, 53, 75, "444", 3, 0, 1, 1), (17, 64, 40, 95, "422", 0, 0, 6, 1),
    (18, 40, 23, 85, "420", 0, 1, 1, 1), (19, 120, 88, 100, "420", 0, 0, 1, 1), (20, 33, 33, 35, "411", 2, 0, 1, 1),
]


def jpeg_case_bytes(cv2, case) -> bytes:
    """The JPEG file of a case (encoded with OpenCV, EXIF orientation spliced in as an APP1 segment)."""
    seed, h, w, q, sampling, restart, grey, orientation = case[:8]
    progressive = len(case) > 8 and case[8]
    rng = np.random.default_rng(2000 + seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 100 * np.sin(xx / 17.0 + c) + 60 * np.cos(yy / 11.0 - c) for c in range(3)], -1)
    img = np.clip(img + rng.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)
    if q == 35:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    params = [cv2.IMWRITE_JPEG_QUALITY, q]
    if grey:
        img = img[:, :, 0]
    else:
        params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, getattr(cv2, f"IMWRITE_JPEG_SAMPLING_FACTOR_{sampling}")]
    if restart:
        params += [cv2.IMWRITE_JPEG_RST_INTERVAL, restart]
    if progressive:
        params += [cv2.IMWRITE_JPEG_PROGRESSIVE, 1]
    ok, enc = cv2.imencode(".jpg", img, params)
    assert ok
    data = bytes(enc)
    if orientation != 1:
        exif = (b"Exif\x00\x00MM\x00\x2a\x00\x00\x00\x08\x00\x01\x01\x12\x00\x03\x00\x00\x00\x01"
                + bytes([0, orientation]) + b"\x00\x00\x00\x00\x00\x00")
        data = data[:2] + b"\xff\xe1" + (len(exif) + 2).to_bytes(2, "big") + exif + data[2:]
    return data


def make_jpeg() -> int:
    """JPEG files + what the reference's own load_image (wicca/data_loader.py:27-63) returns for them."""
    import tempfile
    sys.path.insert(0, "/root/reference")
    import cv2
    from wicca.data_loader import load_image
    out = {"cases": np.array([json.dumps(JPEG_CASES)])}
    with tempfile.TemporaryDirectory() as tmp:
        for i, case in enumerate(JPEG_CASES):
            data = jpeg_case_bytes(cv2, case)
            path = os.path.join(tmp, f"case{i}.jpg")
            with open(path, "wb") as fh:
                fh.write(data)
            rgb = load_image(path)
            assert rgb is not None and rgb.dtype == np.uint8 and rgb.ndim == 3
            out[f"file_{i}"] = np.frombuffer(data, dtype=np.uint8)
            out[f"rgb_{i}"] = rgb
    np.savez_compressed(os.path.join(HERE, "jpeg_golden.npz"), **out)
    print(f"wrote {len(JPEG_CASES)} JPEG cases; cv2 {cv2.__version__} ({[l.strip() for l in cv2.getBuildInformation().splitlines() if 'JPEG:' in l][0]})")
    return 0


def main() -> int:
    if len(sys.argv) > 1 and sys.argv[1] == "jpeg":
        return make_jpeg()
    sys.path.insert(0, "/root/reference")
    import cv2
    from wicca.wavelet_coder import HaarCoder

    ref = HaarCoder()
    out = {"cases": np.array([json.dumps(ICON_CASES)])}
    for i, (kind, seed, h, w, c, d, bt, bc) in enumerate(ICON_CASES):
        img = gen_input(kind, seed, h, w, c)
        before = img.copy()
        icon = ref.get_small_copy(img, d, bt, bc)
        assert (img == before).all()
        out[f"icon_{i}"] = icon
    np.savez_compressed(os.path.join(HERE, "haar_icon_golden.npz"), **out)

    rz = {"cases": np.array(RESIZE_CASES, dtype=np.int64)}
    for i, (seed, sh, sw, dw, dh) in enumerate(RESIZE_CASES):
        src = gen_input("noise", 1000 + seed, sh, sw, 3)
        rz[f"out_{i}"] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA)
    np.savez_compressed(os.path.join(HERE, "resize_area_golden.npz"), **rz)
    print(f"wrote {len(ICON_CASES)} icon cases, {len(RESIZE_CASES)} resize cases; cv2 {cv2.__version__}, numpy {np.__version__}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
