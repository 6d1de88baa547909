"""Property tests of the icon path (no oracle needed at full size) and hypothesis-driven random
parity against the oracle at small sizes."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import c_oracle
from oracle import haar_oracle as ho
from wicca_b200 import HaarCoder

pytestmark = pytest.mark.gpu

coder = HaarCoder()


@settings(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(h=st.integers(1, 300), w=st.integers(1, 300), c=st.sampled_from([2, 3, 3, 3, 4]), depth=st.integers(1, 8),
       border=st.sampled_from([0, 1, 2, 3, 4]), bconst=st.integers(-10, 300), seed=st.integers(0, 2 ** 31 - 1))
def test_random_parity_against_oracle(h, w, c, depth, border, bconst, seed):
    img = np.random.default_rng(seed).integers(0, 256, (h, w, c), dtype=np.uint8)
    got = coder.get_small_copy(img, depth, border, bconst)
    exp = ho.haar_icon_blocksum(img, depth, border, bconst)
    assert got.shape == exp.shape and np.array_equal(got, exp)


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(h=st.integers(1, 700), w=st.integers(1, 1100), seed=st.integers(0, 2 ** 31 - 1),
       depths=st.lists(st.integers(1, 6), min_size=1, max_size=6, unique=True), border=st.sampled_from([0, 1, 2, 3, 4]))
def test_random_multi_depth_sets(h, w, seed, depths, border):
    img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
    outs = coder.get_small_copies(img, depths, border, 77)
    for d, o in zip(depths, outs):
        assert np.array_equal(o, ho.haar_icon_blocksum(img, d, border, 77)), (h, w, d, border)


def test_composition_and_mass_conservation_at_16384():
    """configs[2] size.  With pixel values that are multiples of 64 no level truncates for depth <= 3,
    so icon_3 == icon_1(icon_1(icon_1(x))) and sum(icon_d) * 4^d == sum(x) exactly."""
    rng = np.random.default_rng(5)
    img = (rng.integers(0, 4, (16384, 16384, 3), dtype=np.uint8) * 64).astype(np.uint8)
    i1, i2, i3 = coder.get_small_copies(img, [1, 2, 3])
    total = int(img.sum(dtype=np.uint64))
    assert int(i1.sum(dtype=np.uint64)) * 4 == total
    assert int(i2.sum(dtype=np.uint64)) * 16 == total
    assert int(i3.sum(dtype=np.uint64)) * 64 == total
    assert np.array_equal(coder.get_small_copy(i1, 1), i2)
    assert np.array_equal(coder.get_small_copy(i2, 1), i3)
    # idempotence on constants and monotonicity under a global shift
    assert (coder.get_small_copy(np.full((4096, 6000, 3), 201, np.uint8), 6) == 201).all()
    base = rng.integers(0, 200, (2048, 3000, 3), dtype=np.uint8)
    a, b = coder.get_small_copy(base, 4), coder.get_small_copy(base + 55, 4)
    assert np.array_equal(b, a + 55)          # adding a constant to every pixel shifts every exact block mean


def test_full_size_against_c_oracle():
    """configs[1] shape, all depths, all border types, against the C restatement (fast at 53 MP)."""
    img = np.random.default_rng(11).integers(0, 256, (6393, 8284, 3), dtype=np.uint8)
    for border in (1, 4, 3):
        icons = coder.get_small_copies(img, [1, 2, 3, 4, 5, 6], border, 0)
        for d, ic in zip(range(1, 7), icons):
            if d in (1, 3, 6):
                assert np.array_equal(ic, c_oracle.haar_icon(img, d, border, 0)), (d, border)
