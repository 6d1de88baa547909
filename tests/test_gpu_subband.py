"""Full sub-band forward / inverse (extension row A4) against the NumPy restatement; exact."""
import numpy as np
import pytest

from oracle import haar_oracle as ho
from tests.golden.make_golden import gen_input
from wicca_b200 import HaarCoder

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def coder():
    return HaarCoder()


@pytest.mark.parametrize("shape,depth,border", [((37, 53, 3), 1, 1), ((37, 53, 3), 3, 4), ((64, 128, 3), 6, 1),
                                                ((200, 259, 3), 5, 2), ((130, 517, 4), 2, 0), ((96, 64, 1), 4, 1),
                                                ((300, 500, 3), 8, 3)])
def test_forward_matches_oracle_and_icon(coder, shape, depth, border):
    img = gen_input("noise", sum(shape) + depth, *shape)
    got = coder.forward(img, depth, border, 33)
    exp = ho.haar_forward(img, depth, border, 33)
    assert len(got) == len(exp) == depth + 1
    assert got[0].dtype == np.float32 and np.array_equal(got[0], exp[0])
    for (glh, ghl, ghh), (elh, ehl, ehh) in zip(got[1:], exp[1:]):
        assert np.array_equal(glh, elh) and np.array_equal(ghl, ehl) and np.array_equal(ghh, ehh)
    # LL truncated is the reference's icon
    assert np.array_equal(got[0].astype(np.uint8), ho.haar_icon_fp32(img, depth, border, 33))


@pytest.mark.parametrize("depth", [1, 3, 6, 8])
def test_round_trip_is_exact(coder, depth):
    img = gen_input("noise", 40 + depth, 517, 771, 3)
    rec = coder.inverse(coder.forward(img, depth))
    pad = ho.get_padded_copy(img, 2 ** depth).astype(np.float32)
    assert rec.shape == pad.shape and np.max(np.abs(rec - pad)) == 0.0


def test_inverse_matches_oracle_on_arbitrary_coefficients(coder):
    rng = np.random.default_rng(7)
    ll = rng.integers(-512, 512, (5, 7, 3)).astype(np.float32) / 4
    co = [ll]
    for lvl in range(3):
        shp = (5 << lvl, 7 << lvl, 3)
        co.append(tuple(rng.integers(-512, 512, shp).astype(np.float32) / 4 for _ in range(3)))
    assert np.array_equal(coder.inverse(co), ho.haar_inverse(co))


def test_config3_large_round_trip(coder):
    """BASELINE.json configs[2] (scaled to fit the test budget: 8192x8192x3; bench.py runs 16384^2)."""
    img = gen_input("noise", 3, 8192, 8192, 3)
    for depth in (1, 6):
        co = coder.forward(img, depth)
        assert np.array_equal(co[0].astype(np.uint8), coder.get_small_copy(img, depth))
        rec = coder.inverse(co)
        assert np.max(np.abs(rec - img.astype(np.float32))) == 0.0
        # linearity / energy: LL mean equals image mean (Haar LL is a block average)
        assert abs(float(co[0].mean(dtype=np.float64)) - float(img.mean(dtype=np.float64))) < 1e-6
