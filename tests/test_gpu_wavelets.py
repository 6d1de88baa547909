"""GPU tests of row N4: the longer orthogonal wavelets behind the WaveletCoder interface.  The kernels must agree
bit for bit with the CPU restatement (oracle/fir_oracle.py); with the Haar taps they must agree with HaarCoder,
i.e. with the reference."""
import numpy as np
import pytest

from oracle import fir_oracle as fo
from oracle import haar_oracle as ho
from tests.golden.make_golden import gen_input

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["db2", "db3", "db4", "coif1"])
def test_matches_oracle(name):
    from wicca_b200 import OrthogonalWaveletCoder
    coder = OrthogonalWaveletCoder(name)
    assert np.array_equal(coder.taps, fo.taps_f32(name))
    for (h, w, c, d, bt, bc) in [(64, 64, 3, 1, 1, 0), (77, 131, 3, 2, 1, 0), (300, 201, 3, 3, 4, 0), (129, 67, 3, 4, 0, 37), (128, 64, 1, 3, 1, 0),
                                 (500, 333, 3, 5, 2, 0), (40, 56, 4, 2, 3, 0), (1, 1, 3, 3, 1, 0), (257, 1024, 3, 6, 1, 0),
                                 # float planes whose rows are only 8-byte aligned (interior tiles fetched in 8-byte chunks)
                                 (200, 516, 3, 2, 1, 0), (260, 520, 3, 3, 4, 0)]:
        img = gen_input("noise", 31 * h + w, h, w, c)
        got = coder.get_small_copy(img, d, bt, bc)
        exp = fo.wavelet_icon(img, d, name, bt, bc)
        assert got.shape == exp.shape and got.dtype == np.uint8
        assert np.array_equal(got, exp), (name, h, w, c, d, bt, bc)


def test_haar_taps_equal_haarcoder_and_the_reference_formula():
    from wicca_b200 import DaubechiesCoder, HaarCoder
    fir, haar = DaubechiesCoder(1), HaarCoder()
    for (h, w, d, bt) in [(96, 128, 3, 1), (77, 131, 2, 4), (301, 200, 5, 2), (513, 255, 1, 0)]:
        img = gen_input("noise", h + w, h, w, 3)
        a = fir.get_small_copy(img, d, bt, 9)
        assert np.array_equal(a, haar.get_small_copy(img, d, bt, 9))
        assert np.array_equal(a, ho.haar_icon_fp32(img, d, bt, 9))


def test_interface_and_errors():
    from wicca_b200 import CoifletCoder, DaubechiesCoder, OrthogonalWaveletCoder, WaveletCoder
    coder = CoifletCoder()
    assert isinstance(coder, WaveletCoder)
    img = gen_input("noise", 5, 50, 70, 3)
    assert np.array_equal(coder.get_small_copy(img, 0), img)                  # depth 0 returns the image, like the reference
    assert coder.get_small_copy(image=img, transform_depth=2).shape == (13, 18, 3)
    grey = gen_input("noise", 6, 40, 40, 1)[:, :, 0]
    with pytest.raises(IndexError):                                           # 2-D input fails in the reference's loop too
        coder.get_small_copy(grey, 2)
    with pytest.raises(ValueError):
        coder.get_small_copy(img.astype(np.float32), 2)
    with pytest.raises(ValueError):
        coder.get_small_copy(None, 2)
    with pytest.raises(TypeError):
        coder.get_small_copy(img, (2,))
    with pytest.raises(ValueError):
        OrthogonalWaveletCoder("db9")
    with pytest.raises(ValueError):
        OrthogonalWaveletCoder([0.5, 0.25, 0.25])
    with pytest.raises(ValueError):
        DaubechiesCoder(7)
    custom = OrthogonalWaveletCoder(fo.DEC_LO["db2"])
    assert np.array_equal(custom.get_small_copy(img, 2), DaubechiesCoder(2).get_small_copy(img, 2))


def test_long_custom_filter_and_two_channels_take_the_general_kernels():
    """10 taps (not instantiated in the tiled kernel) and C = 2 (neither): same definition, same oracle."""
    from wicca_b200 import OrthogonalWaveletCoder
    rng = np.random.default_rng(3)
    taps10 = rng.normal(size=10)
    taps10 = taps10 / taps10.sum() * np.sqrt(2.0)
    fo.DEC_LO["_custom10"] = list(taps10)
    try:
        coder = OrthogonalWaveletCoder(taps10)
        for (h, w, c, d) in [(96, 130, 3, 2), (64, 64, 2, 3), (33, 200, 4, 1)]:
            img = gen_input("noise", h * 7 + w, h, w, c)
            assert np.array_equal(coder.get_small_copy(img, d), fo.wavelet_icon(img, d, "_custom10")), (h, w, c, d)
        db2 = OrthogonalWaveletCoder("db2")
        img2 = gen_input("noise", 77, 100, 60, 2)
        assert np.array_equal(db2.get_small_copy(img2, 2), fo.wavelet_icon(img2, 2, "db2"))
    finally:
        del fo.DEC_LO["_custom10"]


@pytest.mark.parametrize("name,side", [("db2", 2048), ("db4", 1024), ("coif1", 1536)])
def test_millions_of_elements_catch_rare_rounding_events(name, side):
    """A fused multiply-add in place of a rounded product followed by a rounded sum changes a few results per million
    (ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 whatever the modifiers say; the column pass once did that):
    small images pass by luck, so compare every byte of images with millions of elements, at depths 1 and 3."""
    from wicca_b200 import OrthogonalWaveletCoder
    coder = OrthogonalWaveletCoder(name)
    img = np.random.default_rng(5).integers(0, 256, (side, side, 3), dtype=np.uint8)
    for d in (1, 3):
        assert np.array_equal(coder.get_small_copy(img, d), fo.wavelet_icon(img, d, name)), (name, side, d)
