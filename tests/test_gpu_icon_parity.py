"""GPU parity of the icon path through the C ABI (HaarCoder -> ctypes -> CUDA) against the
oracle and the reference goldens.  Bit-exact: the icon is integer arithmetic."""
import threading

import numpy as np
import pytest

from oracle import haar_oracle as ho
from tests.golden.make_golden import gen_input
from wicca_b200 import HaarCoder

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def coder():
    return HaarCoder()


def test_reference_goldens_bit_exact(coder, icon_golden):
    cases, outs = icon_golden
    for (kind, seed, h, w, c, d, bt, bc), exp in zip(cases, outs):
        img = gen_input(kind, seed, h, w, c)
        keep = img.copy()
        got = coder.get_small_copy(img, d, bt, bc)
        assert got.dtype == np.uint8 and got.shape == exp.shape, (h, w, c, d, bt)
        assert got.flags.c_contiguous and got.flags.writeable and got.base is None
        assert np.array_equal(got, exp), (kind, seed, h, w, c, d, bt, bc)
        assert np.array_equal(img, keep)                        # input never mutated


def test_known_answers(coder):
    one = np.zeros((2, 2, 3), np.uint8); one[0, 1] = 1
    assert coder.get_small_copy(one, 1).ravel().tolist() == [0, 0, 0]
    t = np.full((2, 2, 3), 255, np.uint8); t[1, 1] = 254
    assert coder.get_small_copy(t, 1).ravel().tolist() == [254, 254, 254]
    for d in range(1, 9):
        assert (coder.get_small_copy(np.full((70, 130, 3), 255, np.uint8), d) == 255).all()
        assert (coder.get_small_copy(np.zeros((70, 130, 3), np.uint8), d) == 0).all()


@pytest.mark.parametrize("border", [1, 0, 2, 3, 4])
def test_random_shapes_vs_oracle(coder, border):
    rng = np.random.default_rng(100 + border)
    for _ in range(40):
        h, w = int(rng.integers(1, 400)), int(rng.integers(1, 700))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        d = int(rng.integers(1, 9))
        bc = int(rng.integers(0, 256))
        got = coder.get_small_copy(img, d, border, bc)
        assert np.array_equal(got, ho.haar_icon_blocksum(img, d, border, bc)), (h, w, d, border, bc)


@pytest.mark.parametrize("c", [1, 2, 4])
def test_other_channel_counts(coder, c):
    rng = np.random.default_rng(c)
    img = rng.integers(0, 256, (96, 160, c), dtype=np.uint8)       # divisible: C=1 works in the reference too
    for d in (1, 3, 5):
        assert np.array_equal(coder.get_small_copy(img, d), ho.haar_icon_fp32(img, d))
    if c > 1:
        img = rng.integers(0, 256, (97, 161, c), dtype=np.uint8)
        for d in (1, 2, 4):
            for bt in (1, 2, 0):
                assert np.array_equal(coder.get_small_copy(img, d, bt, 17), ho.haar_icon_fp32(img, d, bt, 17))


def test_deep_levels_replay_fp32(coder):
    img = gen_input("noise", 77, 700, 1100, 3)
    for d in (7, 8, 9, 10, 11):
        assert np.array_equal(coder.get_small_copy(img, d), ho.haar_icon_fp32(img, d)), d


def test_multi_depth_one_pass_equals_single_depth(coder):
    img = gen_input("noise", 5, 1237, 2011, 3)
    depths = [1, 2, 3, 4, 5, 6]
    multi = coder.get_small_copies(img, depths)
    for d, m in zip(depths, multi):
        assert np.array_equal(m, ho.haar_icon_blocksum(img, d)), d
        assert np.array_equal(m, coder.get_small_copy(img, d)), d
    mixed = coder.get_small_copies(img, [3, 0, 8, 3, 6])             # copies, deep, duplicate depths
    assert np.array_equal(mixed[1], img)
    assert np.array_equal(mixed[0], mixed[3]) and np.array_equal(mixed[0], ho.haar_icon_blocksum(img, 3))
    assert np.array_equal(mixed[2], ho.haar_icon_blocksum(img, 8))
    assert np.array_equal(mixed[4], ho.haar_icon_blocksum(img, 6))


def test_non_contiguous_and_keyword_call(coder):
    big = gen_input("noise", 9, 300, 500, 3)
    view = big[10:201, 20:333]                                     # row-strided view
    assert np.array_equal(coder.get_small_copy(image=view, transform_depth=3),
                          ho.haar_icon_blocksum(np.ascontiguousarray(view), 3))
    fort = np.asfortranarray(big)
    assert np.array_equal(coder.get_small_copy(fort, 2), ho.haar_icon_blocksum(big, 2))
    assert np.array_equal(coder.get_small_copy(big, np.int64(2)), ho.haar_icon_blocksum(big, 2))
    assert np.array_equal(coder.get_small_copy(big, True), ho.haar_icon_blocksum(big, 1))


def test_config1_4096_depth3(coder):
    """BASELINE.json configs[0]: 4096x4096x3, depth 3."""
    img = gen_input("noise", 0, 4096, 4096, 3)
    got = coder.get_small_copy(img, 3)
    assert got.shape == (512, 512, 3)
    assert np.array_equal(got, ho.haar_icon_blocksum(img, 3))


def test_config2_headline_shape_all_depths(coder):
    """BASELINE.json configs[1] shape: (6393, 8284, 3), depths 1..6, ragged on both axes."""
    img = gen_input("noise", 1, 6393, 8284, 3)
    icons = coder.get_small_copies(img, [1, 2, 3, 4, 5, 6])
    shapes = [(3197, 4142), (1599, 2071), (800, 1036), (400, 518), (200, 259), (100, 130)]
    for d, (ic, shp) in enumerate(zip(icons, shapes), start=1):
        assert ic.shape == shp + (3,)
        assert np.array_equal(ic, ho.haar_icon_blocksum(img, d)), d
    # size-independent property: block sums nest, so icon_{d+1} is within 1 LSB-sum of the 2x2
    # mean of icon_d only up to truncation; check the exact relation on the sums instead via zeros/full
    t = coder.last_timing
    assert t and t["kernel_ms"] > 0 and t["h2d_ms"] > 0


def test_thread_pool_reentrancy(coder):
    """The reference calls one shared coder from a ThreadPoolExecutor (classifying_tools.py:414-418)."""
    imgs = [gen_input("noise", 200 + i, 500 + 37 * i, 700 + 11 * i, 3) for i in range(8)]
    exp = [ho.haar_icon_blocksum(im, 1 + i % 6) for i, im in enumerate(imgs)]
    errs = []

    def work(i):
        try:
            for _ in range(5):
                got = coder.get_small_copy(imgs[i], 1 + i % 6)
                if not np.array_equal(got, exp[i]):
                    errs.append(i)
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs


def test_plan_batch_device_resident():
    import torch
    from wicca_b200.plan import IconPlan, to_device_pitched
    rng = np.random.default_rng(3)
    shapes = [(640, 829), (6393 // 4, 8284 // 4), (64, 128), (333, 1000), (65, 129)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    dev = [to_device_pitched(im) for im in imgs]
    depths = [1, 2, 3, 4, 5, 6]
    for border in (1, 4, 0):
        plan = IconPlan(0, [t.data_ptr() for t in dev], [s[0] for s in shapes], [s[1] for s in shapes],
                        [t.shape[1] for t in dev], depths, border_type=border, border_constant=99)
        plan.launch(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        for i, im in enumerate(imgs):
            for k, d in enumerate(depths):
                assert np.array_equal(plan.read_icon(i, k), ho.haar_icon_blocksum(im, d, border, 99)), (i, d, border)
        info = plan.info()
        assert info["launches"] >= 1 and info["bytes_read"] == sum(h * w * 3 for h, w in shapes)
        plan.close()
