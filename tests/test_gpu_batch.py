"""Sharded host batch (row (e)): image i -> devices[i % n]; output order == input order and every
icon equals the single-call result bit for bit, for however many GPUs the box has."""
import numpy as np
import pytest

from oracle import haar_oracle as ho
from tests.golden.make_golden import gen_input
from wicca_b200 import HaarCoder, _capi

pytestmark = pytest.mark.gpu


def ragged_images(n, seed=0):
    rng = np.random.default_rng(seed)
    return [gen_input("noise", 300 + i, 900 + int(rng.integers(-256, 257)), 1300 + int(rng.integers(-256, 257)), 3)
            for i in range(n)]


def test_batch_matches_single_calls_in_order():
    coder = HaarCoder()
    imgs = ragged_images(13)
    depths = [2, 3, 4, 5, 6]
    ndev = _capi.load().wicca_device_count()
    for devices in ([0], list(range(ndev)), [0, 0]):         # [0,0]: two workers sharing one GPU
        out = coder.get_small_copies_batch(imgs, depths, devices=devices)
        assert len(out) == len(imgs)
        for im, row in zip(imgs, out):
            for d, ic in zip(depths, row):
                assert np.array_equal(ic, ho.haar_icon_blocksum(im, d)), (im.shape, d, devices)
    t = coder.last_timing
    assert t["h2d_ms"] > 0 and t["kernel_ms"] > 0


def test_batch_mixed_depths_and_borders():
    coder = HaarCoder()
    imgs = ragged_images(5, seed=4)
    out = coder.get_small_copies_batch(imgs, [0, 1, 7], border_type=4, border_constant=0)
    for im, row in zip(imgs, out):
        assert np.array_equal(row[0], im)
        assert np.array_equal(row[1], ho.haar_icon_blocksum(im, 1, 4))
        assert np.array_equal(row[2], ho.haar_icon_blocksum(im, 7, 4))


def test_pinned_host_buffers_round_trip():
    import ctypes as C
    lib = _capi.load()
    h, w = 1000, 1500
    p = C.c_void_p()
    _capi.check(lib.wicca_host_alloc(C.byref(p), h * w * 3))
    buf = np.ctypeslib.as_array((C.c_uint8 * (h * w * 3)).from_address(p.value)).reshape(h, w, 3)
    buf[:] = gen_input("noise", 8, h, w, 3)
    got = HaarCoder().get_small_copy(buf, 3)
    assert np.array_equal(got, ho.haar_icon_blocksum(np.array(buf), 3))
    del buf
    _capi.check(lib.wicca_host_free(p))
