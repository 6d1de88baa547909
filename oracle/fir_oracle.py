"""TEST INFRASTRUCTURE ONLY - never imported by the product (wicca_b200/).

Oracle for row N4 of the hot-path table: other orthogonal wavelets (Daubechies, Coiflet) behind the reference's
``WaveletCoder`` interface (wicca/wavelet_coder.py:26-38).  PARITY UNPINNED for the long filters: the reference
implements only Haar (README.md:25 and :222 list the others as a roadmap) and ``pywt`` is not installed, so there is
nothing to be at parity with.  What *is* pinned:

  * with the Haar taps ``[1/2, 1/2]`` this construction computes exactly the reference's icon (same float32
    operations up to exact reassociation), which ``tests/test_oracle_fir.py`` checks against the reference goldens;
  * the filter banks are checked mathematically (sum = sqrt 2, unit norm, orthogonal to their even shifts, vanishing
    moments), which fixes them up to the sign / reversal convention stated here.

Definition.  Pad bottom/right to a multiple of 2^depth exactly as ``get_padded_copy`` does (data_loader.py:66-117),
convert to float32, and repeat ``depth`` times on the current (h, w, C) plane:

    rows:     t[y, j]  = sum_n g[n] * x[y, (2 j + n - c) mod w]        n = 0 .. L-1, accumulated left to right
    columns:  ll[i, j] = sum_m g[m] * t[(2 i + m - c) mod h, j]

with ``g = dec_lo / sqrt(2)`` rounded to float32 (so that a constant image keeps its value, like the reference's mean
of 2 x 2), ``c = L/2 - 1``, every product and every sum rounded to float32 separately (no FMA).  Finally clip to
[0, 255] and truncate to uint8 like ``HaarCoder.get_small_copy`` (wavelet_coder.py:66-67).
"""
from __future__ import annotations

import numpy as np

from .haar_oracle import get_padded_copy, validate_image

# decomposition low-pass filters (orthonormal, sum = sqrt 2), as tabulated by Daubechies / PyWavelets
DEC_LO = {
    "haar": [0.7071067811865476, 0.7071067811865476],
    "db2": [-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025],
    "db3": [0.035226291882100656, -0.08544127388224149, -0.13501102001039084, 0.4598775021193313, 0.8068915093133388,
            0.3326705529509569],
    "db4": [-0.010597401784997278, 0.032883011666982945, 0.030841381835986965, -0.18703481171888114,
            -0.02798376941698385, 0.6308807679295904, 0.7148465705525415, 0.23037781330885523],
    "coif1": [-0.01565572813546454, -0.0727326195128539, 0.38486484686420286, 0.8525720202122554, 0.3378976624578092,
              -0.0727326195128539],
}


def taps_f32(name: str) -> np.ndarray:
    """The float32 taps the kernels use: dec_lo / sqrt(2); exactly [0.5, 0.5] for Haar."""
    return (np.asarray(DEC_LO[name], dtype=np.float64) / np.sqrt(2.0)).astype(np.float32)


def lowpass_level(x: np.ndarray, g: np.ndarray) -> np.ndarray:
    """One level on a float32 (h, w, C) plane with even h, w."""
    h, w, _ = x.shape
    n_taps = len(g)
    c = n_taps // 2 - 1
    cols = (2 * np.arange(w // 2)[:, None] + np.arange(n_taps)[None, :] - c) % w          # (w/2, L)
    t = np.zeros((h, w // 2, x.shape[2]), dtype=np.float32)
    for n in range(n_taps):
        prod = (x[:, cols[:, n], :] * g[n]).astype(np.float32)
        t = prod if n == 0 else (t + prod).astype(np.float32)
    rows = (2 * np.arange(h // 2)[:, None] + np.arange(n_taps)[None, :] - c) % h
    out = np.zeros((h // 2, w // 2, x.shape[2]), dtype=np.float32)
    for m in range(n_taps):
        prod = (t[rows[:, m], :, :] * g[m]).astype(np.float32)
        out = prod if m == 0 else (out + prod).astype(np.float32)
    return out


def wavelet_icon(image: np.ndarray, transform_depth: int, name: str, border_type: int = 1, border_constant=0) -> np.ndarray:
    validate_image(image)
    depth = int(transform_depth)
    if depth <= 0:
        return image.copy()
    x = get_padded_copy(image, 2 ** depth, border_type, border_constant).astype(np.float32)
    g = taps_f32(name)
    for _ in range(depth):
        x = lowpass_level(x, g)
    return np.clip(x, 0, 255).astype(np.uint8)
