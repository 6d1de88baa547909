"""The oracle is pinned against outputs of the live reference (committed fixtures)."""
import numpy as np
import pytest

from oracle import haar_oracle as ho
from oracle import resize_oracle as ro
from tests.golden.make_golden import gen_input


def test_icon_oracle_matches_reference_goldens(icon_golden):
    cases, outs = icon_golden
    assert len(cases) > 400
    for (kind, seed, h, w, c, d, bt, bc), exp in zip(cases, outs):
        img = gen_input(kind, seed, h, w, c)
        got = ho.haar_icon_fp32(img, d, bt, bc)
        assert got.shape == exp.shape and got.dtype == np.uint8
        assert np.array_equal(got, exp), (kind, seed, h, w, c, d, bt, bc)
        if 0 < d <= 8:
            assert np.array_equal(ho.haar_icon_blocksum(img, d, bt, bc), exp), ("blocksum", h, w, c, d, bt)


def test_known_answers():
    one = np.zeros((2, 2, 3), np.uint8); one[0, 1] = 1
    assert ho.haar_icon_fp32(one, 1).ravel().tolist() == [0, 0, 0]          # truncation, not rounding
    t = np.full((2, 2, 3), 255, np.uint8); t[1, 1] = 254
    assert ho.haar_icon_fp32(t, 1).ravel().tolist() == [254, 254, 254]
    for d in range(1, 9):
        assert (ho.haar_icon_fp32(np.full((40, 56, 3), 255, np.uint8), d) == 255).all()
        assert (ho.haar_icon_fp32(np.zeros((40, 56, 3), np.uint8), d) == 0).all()
    img = gen_input("noise", 3, 9, 11, 3)
    assert np.array_equal(ho.haar_icon_fp32(img, 0), img) and np.array_equal(ho.haar_icon_fp32(img, -2), img)


def test_forward_inverse_oracle_exact():
    img = gen_input("noise", 11, 37, 53, 3)
    for d in (1, 2, 3, 6, 8):
        co = ho.haar_forward(img, d)
        assert np.array_equal(co[0].astype(np.uint8), ho.haar_icon_fp32(img, d))
        rec = ho.haar_inverse(co)
        assert np.array_equal(rec, ho.get_padded_copy(img, 2 ** d).astype(np.float32))


def test_resize_oracle_matches_cv2_goldens(resize_golden):
    cases, outs = resize_golden
    regimes = set()
    for (seed, sh, sw, dw, dh), exp in zip(cases, outs):
        src = gen_input("noise", 1000 + seed, sh, sw, 3)
        regimes.add(ro.regime(sw, sh, dw, dh))
        assert np.array_equal(ro.resize_area(src, dw, dh), exp), (seed, sh, sw, dw, dh)
    assert regimes == {"fast", "generic", "bilinear"}


def test_resize_oracle_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for (sh, sw) in [(400, 518), (200, 259), (100, 130), (448, 448), (57, 91)]:
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        for t in (224, 331, 96):
            assert np.array_equal(ro.resize_area(src, t, t), cv2.resize(src, (t, t), interpolation=cv2.INTER_AREA))


def test_resize_oracle_live_cv2_headline_sizes_and_random_pairs():
    """The pin of row A5 at scale: the depth-1 icon of the headline image (3197 x 4142) to both classifier sizes, and 60
    random (source, target) pairs over all three regimes incl. non-square targets - the restatement must equal the
    installed cv2 bit for bit (SURVEY.md Appendix A probed 88 pairs the same way)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2026)
    big = rng.integers(0, 256, (3197, 4142, 3), dtype=np.uint8)
    for t in (224, 331):
        assert np.array_equal(ro.resize_area(big, t, t), cv2.resize(big, (t, t), interpolation=cv2.INTER_AREA)), t
    seen = set()
    for _ in range(60):
        sh, sw = int(rng.integers(2, 1700)), int(rng.integers(2, 2200))
        tw = int(rng.choice([224, 240, 299, 331, int(rng.integers(3, 400))]))
        th = int(rng.choice([224, 299, 331, int(rng.integers(3, 400))]))
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        seen.add(ro.regime(sw, sh, tw, th))
        assert np.array_equal(ro.resize_area(src, tw, th), cv2.resize(src, (tw, th), interpolation=cv2.INTER_AREA)), (sh, sw, tw, th)
    assert {"generic", "bilinear"} <= seen


def test_preprocess_modes():
    x = gen_input("noise", 2, 8, 8, 3)[None]
    assert np.array_equal(ro.preprocess_input(x, "identity"), x.astype(np.float32))
    tf = ro.preprocess_input(x, "tf")
    assert tf.dtype == np.float32 and tf.min() >= -1 and tf.max() <= 1
    caffe = ro.preprocess_input(x, "caffe")
    assert np.allclose(caffe[..., 0], x[..., 2].astype(np.float32) - 103.939)
    torch_ = ro.preprocess_input(x, "torch")
    assert np.allclose(torch_[..., 1], (x[..., 1] / 255.0 - 0.456) / 0.224, atol=1e-5)


def test_torch_mode_agrees_with_torchvision_normalize():
    """Row A6 stays UNPINNED (Keras 3.11.3 is not installed, vendored or in the wheelhouse).  The one independent
    implementation this image holds is torchvision's `normalize`, the convention Keras' "torch" mode copies
    (x / 255, then (x - mean) / std per channel, float32): corroboration of one of the four modes, not a pin."""
    import pytest
    torch = pytest.importorskip("torch")
    tvf = pytest.importorskip("torchvision.transforms.functional")
    from oracle import resize_oracle as ro
    batch = np.random.default_rng(4).integers(0, 256, (3, 17, 23, 3), dtype=np.uint8)
    batch[0, :16, :16, 0] = np.arange(256, dtype=np.uint8).reshape(16, 16)          # every byte value at least once
    exp = ro.preprocess_input(batch, "torch")
    x = torch.from_numpy(batch).permute(0, 3, 1, 2).to(torch.float32) / 255.0          # to_tensor's scaling, NCHW
    got = tvf.normalize(x, mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]).permute(0, 2, 3, 1).numpy()
    assert got.dtype == np.float32 and np.array_equal(got, exp)
