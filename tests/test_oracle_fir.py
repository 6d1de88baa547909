"""CPU tests for row N4 (other orthogonal wavelets): the filter banks are what they claim to be, and the oracle's
construction reproduces the reference bit for bit when it is given the Haar taps."""
import json

import numpy as np
import pytest

from oracle import fir_oracle as fo
from tests.golden.make_golden import gen_input


@pytest.mark.parametrize("name,moments", [("haar", 1), ("db2", 2), ("db3", 3), ("db4", 4), ("coif1", 2)])
def test_filter_bank_is_orthonormal_with_vanishing_moments(name, moments):
    h = np.asarray(fo.DEC_LO[name], dtype=np.float64)
    assert abs(h.sum() - np.sqrt(2)) < 1e-9 and abs((h * h).sum() - 1) < 1e-9
    for k in range(1, len(h) // 2):
        assert abs((h[2 * k:] * h[:-2 * k]).sum()) < 1e-9                     # orthogonal to its even shifts
    hi = h[::-1] * (-1.0) ** np.arange(len(h))                                 # quadrature mirror high-pass
    for p in range(moments):
        assert abs((hi * np.arange(len(h)) ** p).sum()) < 1e-7                  # vanishing moments
    assert abs(float(fo.taps_f32(name).sum()) - 1) < 1e-6


def test_haar_taps_reproduce_the_reference_goldens(icon_golden):
    cases, icons = icon_golden
    checked = 0
    for i, (kind, seed, h, w, c, d, bt, bc) in enumerate(cases):
        if not 1 <= d <= 8 or h * w > 600 * 600:
            continue
        img = gen_input(kind, seed, h, w, c)
        assert np.array_equal(fo.wavelet_icon(img, d, "haar", bt, bc), icons[i]), (kind, seed, h, w, c, d, bt, bc)
        checked += 1
    assert checked > 100


def test_constant_image_keeps_its_value():
    img = np.full((64, 96, 3), 177, np.uint8)
    for name in ("db2", "db3", "db4", "coif1"):
        icon = fo.wavelet_icon(img, 3, name)
        assert icon.shape == (8, 12, 3) and np.all((icon == 176) | (icon == 177))      # float32 taps sum to 1 +- 1 ulp


@pytest.mark.parametrize("name", ["db2", "db3", "db4", "coif1"])
def test_level_is_the_periodised_orthogonal_analysis_low_pass(name):
    """An independent float64 formulation of one level - the orthonormal analysis operator built as a matrix from the
    shifted, periodically wrapped filter - must agree with the oracle's gather-and-accumulate loop (to float32 rounding),
    preserve energy together with its quadrature-mirror high-pass (Parseval), and commute with the decimation grid: row
    2k of the operator is the filter placed at 2k - c."""
    rng = np.random.default_rng(7)
    h = np.asarray(fo.DEC_LO[name], dtype=np.float64)
    taps, c, n = len(h), len(h) // 2 - 1, 64
    lo = np.zeros((n // 2, n))
    hi = np.zeros((n // 2, n))
    g_hi = h[::-1] * (-1.0) ** np.arange(taps)
    for k in range(n // 2):
        for m in range(taps):
            lo[k, (2 * k + m - c) % n] += h[m]
            hi[k, (2 * k + m - c) % n] += g_hi[m]
    full = np.vstack([lo, hi])
    assert np.allclose(full @ full.T, np.eye(n), atol=1e-9)                       # an orthonormal transform
    x = rng.integers(0, 256, (n, n, 3)).astype(np.float32)
    want = np.einsum("ik,kjc->ijc", lo / np.sqrt(2), np.einsum("jk,ikc->ijc", lo / np.sqrt(2), x.astype(np.float64)))
    got = fo.lowpass_level(x, fo.taps_f32(name))
    assert got.shape == (n // 2, n // 2, 3) and np.allclose(got, want, rtol=0, atol=2e-3)
