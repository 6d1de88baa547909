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


@pytest.mark.parametrize("w", [769, 770, 772, 776])
def test_round_trip_row_alignments(coder, w):
    """Depth 1 on widths whose sub-band rows start 4-, 8- and 16-byte aligned: the inverse kernel fetches its level-1
    details with cp.async in chunks of that size."""
    img = gen_input("noise", w, 200, w, 3)
    co = coder.forward(img, 1)
    exp = ho.haar_forward(img, 1)
    assert all(np.array_equal(a, b) for a, b in zip(co[1], exp[1])) and np.array_equal(co[0], exp[0])
    rec = coder.inverse(co)
    pad = ho.get_padded_copy(img, 2).astype(np.float32)
    assert rec.shape == pad.shape and np.array_equal(rec, pad)


def test_inverse_matches_oracle_on_arbitrary_coefficients(coder):
    rng = np.random.default_rng(7)
    ll = rng.integers(-512, 512, (5, 7, 3)).astype(np.float32) / 4
    co = [ll]
    for lvl in range(3):
        shp = (5 << lvl, 7 << lvl, 3)
        co.append(tuple(rng.integers(-512, 512, shp).astype(np.float32) / 4 for _ in range(3)))
    assert np.array_equal(coder.inverse(co), ho.haar_inverse(co))


def test_config3_large_round_trip(coder):
    """Host-buffer entry points on a large image (8192x8192x3; the 16384^2 case of BASELINE.json configs[2]
    runs device-resident in test_config3_device_resident_16384 below)."""
    img = gen_input("noise", 3, 8192, 8192, 3)
    for depth in (1, 6):
        co = coder.forward(img, depth)
        assert np.array_equal(co[0].astype(np.uint8), coder.get_small_copy(img, depth))
        rec = coder.inverse(co)
        assert np.max(np.abs(rec - img.astype(np.float32))) == 0.0
        # linearity / energy: LL mean equals image mean (Haar LL is a block average)
        assert abs(float(co[0].mean(dtype=np.float64)) - float(img.mean(dtype=np.float64))) < 1e-6


def test_config3_device_resident_16384():
    """BASELINE.json configs[2]: forward DWT + inverse IDWT round trip, all sub-bands kept, on a
    16384 x 16384 x 3 image, reconstruction error check (device-resident entry points)."""
    import ctypes as C
    import torch
    from wicca_b200 import _capi
    from wicca_b200.plan import pitch_bytes
    lib = _capi.load()
    S = 16384
    pitch = pitch_bytes(S, 3)
    g = torch.Generator(device="cuda:0"); g.manual_seed(7)
    img = torch.randint(0, 256, (S, pitch), dtype=torch.uint8, device="cuda:0", generator=g)
    coeffs = torch.empty((S, S, 3), dtype=torch.float32, device="cuda:0")
    work = torch.empty((S * S * 3 * 5 // 16 + 64,), dtype=torch.float32, device="cuda:0")
    rec = torch.empty((S, S, 3), dtype=torch.float32, device="cuda:0")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ref = img[:, : S * 3].reshape(S, S, 3)
    for depth in (1, 3, 6):
        _capi.check(lib.wicca_haar_forward_dev(img.data_ptr(), S, S, 3, pitch, depth, 1, 0.0, coeffs.data_ptr(),
                                               work.data_ptr(), 0, st), "forward_dev")
        _capi.check(lib.wicca_haar_inverse_dev(coeffs.data_ptr(), S, S, 3, depth, rec.data_ptr(), work.data_ptr(), 0, st),
                    "inverse_dev")
        torch.cuda.synchronize()
        assert float((rec - ref.float()).abs().max().item()) == 0.0, depth
        # LL of the plane, truncated, is the reference's icon (spot check against the oracle on a corner crop)
        n = S >> depth
        ll = coeffs[: 64, : 64].cpu().numpy()
        crop = ref[: 64 << depth, : 64 << depth].cpu().numpy()
        assert np.array_equal(ll.astype(np.uint8), ho.haar_icon_blocksum(crop, depth)), depth
        assert n * (1 << depth) == S


@pytest.mark.parametrize("depth", [1, 6])
def test_every_subband_of_a_large_image_against_the_c_oracle(coder, depth):
    """configs[2] scale: EVERY coefficient of every sub-band of an 8192 x 8192 x 3 image against the C oracle
    (oracle/haar_oracle.c: oracle_haar_forward_f32), not just the round trip - a forward and an inverse that agree with
    each other but mis-place a detail band at a tile seam would fail here."""
    from oracle import c_oracle
    from wicca_b200.wavelet_coder import list_to_mallat
    img = gen_input("noise", 17 + depth, 8192, 8192, 3)
    got, _ = list_to_mallat(coder.forward(img, depth))
    exp = c_oracle.haar_forward_plane(img, depth)
    assert got.shape == exp.shape and np.array_equal(got, exp)
    rec = coder.inverse(coder.forward(img, depth))
    assert np.array_equal(rec, c_oracle.haar_inverse_plane(exp, depth))


@pytest.mark.parametrize("border", [0, 1, 2, 3, 4])
def test_every_subband_of_a_ragged_53mp_image_all_borders(coder, border):
    """The headline shape (6393, 8284, 3) needs padding in both axes at depth 6 (-> 6400 x 8320): all five cv2 border
    rules, every sub-band of every level against the C oracle."""
    from oracle import c_oracle
    from wicca_b200.wavelet_coder import list_to_mallat
    img = gen_input("noise", 90 + border, 6393, 8284, 3)
    got, _ = list_to_mallat(coder.forward(img, 6, border, 201))
    exp = c_oracle.haar_forward_plane(img, 6, border, 201)
    assert got.shape == exp.shape == (6400, 8320, 3)
    assert np.array_equal(got, exp)
