"""The C restatement of the oracle against the reference goldens and the NumPy restatement."""
import numpy as np

from oracle import c_oracle
from oracle import haar_oracle as ho
from tests.golden.make_golden import gen_input


def test_c_oracle_matches_reference_goldens(icon_golden):
    cases, outs = icon_golden
    for (kind, seed, h, w, c, d, bt, bc), exp in zip(cases, outs):
        img = gen_input(kind, seed, h, w, c)
        got = c_oracle.haar_icon(img, d, bt, bc)
        assert got.shape == exp.shape and np.array_equal(got, exp), (kind, seed, h, w, c, d, bt, bc)


def test_c_border_index_matches_numpy():
    lib = c_oracle.load()
    for bt in (1, 2, 3, 4, 0, 17):
        for n in (1, 2, 3, 7, 64):
            for p in range(0, 4 * n + 5):
                assert lib.oracle_border_index(p, n, bt) == ho.border_index(p, n, bt), (p, n, bt)


def test_c_oracle_medium_image_all_depths():
    img = gen_input("noise", 1, 1599, 2071, 3)
    for d in range(1, 11):
        assert np.array_equal(c_oracle.haar_icon(img, d), ho.haar_icon_fp32(img, d)), d
