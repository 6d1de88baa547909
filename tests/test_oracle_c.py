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


def test_c_multi_depth_blocksum_matches_fp32_restatement_and_goldens(icon_golden):
    """oracle_haar_icons_multi_u8 (one pass of exact integer block sums, used by bench.py to check whole
    batches) against the float32 restatement and the reference goldens."""
    cases, outs = icon_golden
    n = 0
    for (kind, seed, h, w, c, d, bt, bc), exp in zip(cases, outs):
        if not (1 <= d <= 8) or bt not in (0, 1, 2, 3, 4) or not (1 <= c <= 4):
            continue
        img = gen_input(kind, seed, h, w, c)
        got = c_oracle.haar_icons_multi(img, [d], bt, bc)[0]
        assert got.shape == exp.shape and np.array_equal(got, exp), (kind, seed, h, w, c, d, bt, bc)
        n += 1
    assert n > 100
    img = gen_input("noise", 2, 333, 517, 3)
    for bt in range(5):
        ds = [1, 2, 3, 4, 5, 6, 7, 8]
        for d, got in zip(ds, c_oracle.haar_icons_multi(img, ds, bt, 200)):
            assert np.array_equal(got, c_oracle.haar_icon(img, d, bt, 200)), (bt, d)


def test_c_subband_forward_inverse_match_numpy_definition():
    from wicca_b200.wavelet_coder import list_to_mallat
    rng = np.random.default_rng(11)
    for (h, w, c) in ((37, 53, 3), (64, 128, 3), (1, 1, 3), (9, 200, 4), (130, 7, 2)):
        img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        for bt in range(5):
            for d in (1, 2, 5):
                plane = c_oracle.haar_forward_plane(img, d, bt, 31)
                exp, _ = list_to_mallat(ho.haar_forward(img, d, bt, 31))
                assert np.array_equal(plane, exp), (h, w, c, bt, d)
                rec = c_oracle.haar_inverse_plane(plane, d)
                assert np.array_equal(rec, ho.get_padded_copy(img, 1 << d, bt, 31).astype(np.float32))
                ll = plane[: plane.shape[0] >> d, : plane.shape[1] >> d]
                assert np.array_equal(ll.astype(np.uint8)[: -(-h // (1 << d)), : -(-w // (1 << d))], ho.haar_icon_fp32(img, d, bt, 31))
