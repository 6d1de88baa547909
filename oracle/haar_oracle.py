"""NumPy restatement of the reference HaarCoder hot path (CPU oracle).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  Not imported by the
product package.

Pinned against the live reference (``/root/reference``, imported in the build
container by ``tests/golden/make_golden.py``) through the committed fixtures in
``tests/golden/`` - the reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so "the reference run here" is the pin.

Every function cites the reference lines it restates (paths relative to
``/root/reference``).  ``cv2.copyMakeBorder`` (OpenCV, pinned
``opencv-python==4.12.0.88`` in ``requirements.txt:91``) is a third-party
dependency of the path; its published ``borderInterpolate`` rule is restated in
:func:`border_index` so the oracle itself needs only NumPy.
"""
from __future__ import annotations

import numpy as np

# OpenCV border type codes (cv2.BORDER_*); only these five are accepted by
# cv2.copyMakeBorder (BORDER_TRANSPARENT=5 and anything else raise cv2.error).
BORDER_CONSTANT = 0
BORDER_REPLICATE = 1
BORDER_REFLECT = 2
BORDER_WRAP = 3
BORDER_REFLECT_101 = 4
BORDER_ISOLATED = 16  # flag bit, masked off by copyMakeBorder

VALID_BORDERS = (BORDER_CONSTANT, BORDER_REPLICATE, BORDER_REFLECT, BORDER_WRAP,
                 BORDER_REFLECT_101)


def validate_image(image) -> None:
    """Restates ``wicca/validation.py:80-101`` (``validate_image``).

    The reference's final ``np.max(image) > 255`` check (``:100``) cannot fire
    for uint8 input and is omitted.
    """
    if image is None:
        raise ValueError("Image didn't found. Please check your input.")
    if image.shape[0] == 0 or image.shape[1] == 0 or image.size == 0:
        raise ValueError("Image is empty")
    if image.dtype != np.uint8:
        raise ValueError("Image must be of type uint8")


def border_index(p: int, n: int, border_type: int) -> int:
    """Source index for padded index ``p`` on an axis of length ``n``.

    Restates OpenCV ``cv::borderInterpolate`` as used by ``cv2.copyMakeBorder``
    at ``wicca/data_loader.py:116-117``.  Returns -1 for BORDER_CONSTANT
    (caller substitutes the constant).  Only ``p >= 0`` occurs on this path
    (padding is bottom/right only, ``data_loader.py:107-110``).
    """
    border_type &= ~BORDER_ISOLATED
    if 0 <= p < n:
        return p
    if border_type == BORDER_REPLICATE:
        return 0 if p < 0 else n - 1
    if border_type in (BORDER_REFLECT, BORDER_REFLECT_101):
        delta = 1 if border_type == BORDER_REFLECT_101 else 0
        if n == 1:
            return 0
        while not (0 <= p < n):
            if p < 0:
                p = -p - 1 + delta
            else:
                p = n - 1 - (p - n) - delta
        return p
    if border_type == BORDER_WRAP:
        return p % n
    if border_type == BORDER_CONSTANT:
        return -1
    raise ValueError(f"Unknown/unsupported border type {border_type}")


def saturate_u8(value) -> int:
    """``cv::saturate_cast<uchar>(double)``: round half to even, clamp to 0..255.

    This is what ``copyMakeBorder`` does with each entry of the ``border_value``
    list built at ``wicca/data_loader.py:115``.
    """
    v = int(np.rint(float(value)))
    return 0 if v < 0 else 255 if v > 255 else v


def padded_shape(h: int, w: int, ratio: int) -> tuple[int, int]:
    """Padded extents, ``wicca/data_loader.py:107-110`` (bottom/right only)."""
    hp = -(-h // ratio) * ratio
    wp = -(-w // ratio) * ratio
    return hp, wp


def get_padded_copy(image: np.ndarray, ratio: int, border_type: int = BORDER_REPLICATE,
                    border_constant=0) -> np.ndarray:
    """Restates ``wicca/data_loader.py:66-117`` (``get_padded_copy``) for 3-D input.

    Returns the same object when no padding is needed (``:112-113``), else a
    padded copy built from the :func:`border_index` map (``:116-117``).
    """
    if not isinstance(image, np.ndarray):
        raise ValueError("Image must be a numpy array")
    if ratio <= 0:
        raise ValueError("Ratio must be positive")
    if image.ndim == 2:
        rows, cols = image.shape
    elif image.ndim == 3:
        rows, cols, _ = image.shape
    else:
        raise ValueError("Image must be 2D or 3D array")
    hp, wp = padded_shape(rows, cols, ratio)
    if hp == rows and wp == cols:
        return image
    bt = border_type & ~BORDER_ISOLATED
    if bt not in VALID_BORDERS:
        raise ValueError(f"Unknown/unsupported border type {border_type}")
    if bt == BORDER_CONSTANT:
        out = np.full((hp, wp) + image.shape[2:], saturate_u8(border_constant), dtype=image.dtype)
        out[:rows, :cols] = image
        return out
    # one full copy plus the two edge strips - the cost profile of cv2.copyMakeBorder (a fancy-index
    # gather of the whole image, as an earlier version did, is ~10x slower than the reference's pad)
    out = np.empty((hp, wp) + image.shape[2:], dtype=image.dtype)
    out[:rows, :cols] = image
    if wp > cols:
        xmap = np.array([border_index(p, cols, bt) for p in range(cols, wp)], dtype=np.intp)
        out[:rows, cols:] = image[:, xmap]
    if hp > rows:
        ymap = np.array([border_index(p, rows, bt) for p in range(rows, hp)], dtype=np.intp)
        out[rows:] = out[ymap]
    return out


def haar_icon_fp32(image: np.ndarray, transform_depth: int,
                   border_type: int = BORDER_REPLICATE, border_constant=0) -> np.ndarray:
    """Step-for-step restatement of ``HaarCoder.get_small_copy``
    (``wicca/wavelet_coder.py:50-67``): validate, pad to a multiple of
    ``2**depth``, upcast to float32, ``depth`` times (row-pair sum, column-pair
    sum, times 0.25), clip, truncate to uint8.  Same NumPy operations in the
    same order, so its CPU cost profile is the reference's.
    """
    validate_image(image)
    ratio = 2 ** transform_depth
    low_left = get_padded_copy(image, ratio, border_type, border_constant).astype(np.float32)
    for _ in range(transform_depth):
        sums = low_left[::2, :, :] + low_left[1::2, :, :]
        low_left = (sums[:, ::2, :] + sums[:, 1::2, :]) * 0.25
    return np.clip(low_left, 0, 255).astype(np.uint8)


def haar_icon_blocksum(image: np.ndarray, transform_depth: int,
                       border_type: int = BORDER_REPLICATE, border_constant=0) -> np.ndarray:
    """Integer identity for the same result:
    ``icon = (sum of each 2^d x 2^d block of the padded image) >> 2d``.

    Exact for depth <= 8 because every fp32 intermediate of
    ``wavelet_coder.py:61-65`` is ``k / 4**level`` with ``k < 2**24``.
    """
    validate_image(image)
    d = int(transform_depth)
    if d <= 0:
        return image.copy()
    if d > 8:
        raise ValueError("blocksum identity only holds for depth <= 8")
    r = 1 << d
    p = get_padded_copy(image, r, border_type, border_constant)
    hp, wp, c = p.shape
    s = p.reshape(hp // r, r, wp // r, r, c).astype(np.uint32).sum(axis=(1, 3), dtype=np.uint32)
    return (s >> (2 * d)).astype(np.uint8)


# --------------------------------------------------------------------------
# Extension: full-subband forward / inverse transform.
# Not in the reference (it keeps LL only); defined so that LL at every level is
# exactly the reference's ``low_left`` (same row-then-column order, same 0.25).
# --------------------------------------------------------------------------

def haar_forward(image: np.ndarray, transform_depth: int,
                 border_type: int = BORDER_REPLICATE, border_constant=0):
    """Full 2-D Haar analysis, fp32.  Returns ``[LL_d, (LH_d, HL_d, HH_d), ...,
    (LH_1, HL_1, HH_1)]`` (pywt ``wavedec2`` ordering).  With the 2x2 block
    ``a b / c d``: ``LL=(a+b+c+d)/4  HL=(a-b+c-d)/4  LH=(a+b-c-d)/4
    HH=(a-b-c+d)/4``; LL follows ``wavelet_coder.py:62-65`` literally."""
    validate_image(image)
    ratio = 2 ** transform_depth
    ll = get_padded_copy(image, ratio, border_type, border_constant).astype(np.float32)
    details = []
    for _ in range(transform_depth):
        ev, od = ll[::2, :, :], ll[1::2, :, :]
        rs, rd = ev + od, ev - od              # row-pair sum / difference
        q = np.float32(0.25)
        new_ll = (rs[:, ::2, :] + rs[:, 1::2, :]) * q
        hl = (rs[:, ::2, :] - rs[:, 1::2, :]) * q   # high-pass along x
        lh = (rd[:, ::2, :] + rd[:, 1::2, :]) * q   # high-pass along y
        hh = (rd[:, ::2, :] - rd[:, 1::2, :]) * q
        details.append((lh, hl, hh))
        ll = new_ll
    return [ll] + details[::-1]


def haar_inverse(coeffs) -> np.ndarray:
    """Synthesis for :func:`haar_forward`: ``a=LL+HL+LH+HH  b=LL-HL+LH-HH
    c=LL+HL-LH-HH  d=LL-HL-LH+HH``.  Returns the fp32 padded image (exact for
    depth <= 8)."""
    ll = np.asarray(coeffs[0], dtype=np.float32)
    for lh, hl, hh in coeffs[1:]:
        h, w, c = ll.shape
        out = np.empty((2 * h, 2 * w, c), dtype=np.float32)
        out[0::2, 0::2] = (ll + hl) + (lh + hh)
        out[0::2, 1::2] = (ll - hl) + (lh - hh)
        out[1::2, 0::2] = (ll + hl) - (lh + hh)
        out[1::2, 1::2] = (ll - hl) - (lh - hh)
        ll = out
    return ll


def synthetic_image(seed: int, h: int, w: int, c: int = 3) -> np.ndarray:
    """Synthetic input generator of SURVEY.md 8(d): uniform uint8 noise."""
    return np.random.default_rng(seed).integers(0, 256, (h, w, c), dtype=np.uint8)
