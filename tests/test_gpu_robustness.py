"""Error paths and lifecycle of the C ABI on a real device."""
import ctypes as C

import numpy as np
import pytest

from oracle import haar_oracle as ho
from tests.golden.make_golden import gen_input
from wicca_b200 import HaarCoder, _capi

pytestmark = pytest.mark.gpu


def test_bad_device_is_an_argument_error():
    lib = _capi.load()
    img = np.zeros((8, 8, 3), np.uint8)
    out = np.zeros((4, 4, 3), np.uint8)
    n = lib.wicca_device_count()
    assert n >= 1
    rc = lib.wicca_haar_icon_u8(img.ctypes.data, 8, 8, 3, 0, 1, 1, 0.0, out.ctypes.data, n + 3, None)
    assert rc == _capi.EDEVICE and b"device" in lib.wicca_last_error()
    c = HaarCoder()
    c.device = n + 3
    with pytest.raises(ValueError):
        c.get_small_copy(img, 1)


def test_shutdown_then_reuse():
    lib = _capi.load()
    c = HaarCoder()
    img = gen_input("noise", 5, 300, 400, 3)
    exp = ho.haar_icon_blocksum(img, 2)
    assert np.array_equal(c.get_small_copy(img, 2), exp)
    assert lib.wicca_shutdown() == 0            # frees every cached stream / buffer
    assert np.array_equal(c.get_small_copy(img, 2), exp)
    assert np.array_equal(c.get_small_copies(img, [1, 2])[1], exp)


def test_growing_and_shrinking_images_reuse_scratch():
    c = HaarCoder()
    rng = np.random.default_rng(0)
    for (h, w) in [(64, 64), (3000, 4000), (10, 10), (4096, 4096), (1, 7), (2000, 17), (17, 2000)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for d in (1, 4, 6):
            assert np.array_equal(c.get_small_copy(img, d), ho.haar_icon_blocksum(img, d)), (h, w, d)


def test_plan_argument_errors():
    lib = _capi.load()
    h = C.c_void_p()
    one = (C.c_void_p * 1)(0)
    rc = lib.wicca_plan_create(0, 1, one, (C.c_int * 1)(8), (C.c_int * 1)(8), (C.c_int64 * 1)(128), 3, (C.c_int * 1)(3), 1, 1,
                               0.0, C.byref(h))
    assert rc == _capi.EINVAL                   # NULL image pointer
    import torch
    t = torch.zeros((8, 128), dtype=torch.uint8, device="cuda:0")
    ptr = (C.c_void_p * 1)(t.data_ptr())
    rc = lib.wicca_plan_create(0, 1, ptr, (C.c_int * 1)(8), (C.c_int * 1)(8), (C.c_int64 * 1)(128), 3, (C.c_int * 1)(9), 1, 1,
                               0.0, C.byref(h))
    assert rc == _capi.EDEPTH                   # plans take depths 1..8
    rc = lib.wicca_plan_create(0, 1, ptr, (C.c_int * 1)(8), (C.c_int * 1)(8), (C.c_int64 * 1)(8), 3, (C.c_int * 1)(2), 1, 1,
                               0.0, C.byref(h))
    assert rc == _capi.EINVAL                   # pitch < W*C
