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


def test_batch_over_every_gpu_of_the_box():
    """The multi-GPU code path proper: image i -> GPU i % n for every GPU the box has (skips only on a
    one-GPU box), ragged shapes, compared image by image with the single-GPU result and the oracle."""
    ndev = _capi.load().wicca_device_count()
    if ndev < 2:
        pytest.skip("needs at least two GPUs")
    coder = HaarCoder()
    imgs = ragged_images(4 * ndev + 3, seed=9)
    depths = [1, 2, 3, 4, 5, 6]
    multi = coder.get_small_copies_batch(imgs, depths, devices=list(range(ndev)))
    single = coder.get_small_copies_batch(imgs, depths, devices=[0])
    for im, rm, rs in zip(imgs, multi, single):
        for d, a, b in zip(depths, rm, rs):
            assert np.array_equal(a, b) and np.array_equal(a, ho.haar_icon_blocksum(im, d)), (im.shape, d)
    # every GPU can also be addressed on its own
    for dev in range(ndev):
        coder.device = dev
        assert np.array_equal(coder.get_small_copy(imgs[dev], 3), ho.haar_icon_blocksum(imgs[dev], 3))


def test_batch_writes_into_caller_arrays_and_registered_memory():
    """`out=`: icons land in arrays the caller owns - here views of an IconArena segment page-locked with
    wicca_host_register, as the one-node gather of wicca_b200/sharding.py uses them."""
    from wicca_b200.sharding import IconArena, sharded_small_copies
    coder = HaarCoder()
    imgs = ragged_images(7, seed=2)
    depths = [0, 2, 5]
    arena = IconArena([im.shape for im in imgs], depths, pin=True)
    out = sharded_small_copies(lambda i: imgs[i], len(imgs), depths,
                               lambda images, ds, out=None: coder.get_small_copies_batch(images, ds, out=out), arena=arena)
    for i, (im, row) in enumerate(zip(imgs, out)):
        for d, ic, view in zip(depths, row, arena.views(i)):
            assert ic.shape == view.shape and np.shares_memory(ic, view)
            assert np.array_equal(ic, ho.haar_icon_blocksum(im, d)), (im.shape, d)
    with pytest.raises(ValueError):
        coder.get_small_copies_batch(imgs[:1], depths, out=[[np.empty((3, 3, 3), np.uint8)] * 3])
    del out, row, ic, view
    arena.close()


def test_batch_page_locked_buffers_take_the_flat_link_copies():
    """Batch workers move large page-locked images and icons over the link as flat copies and re-pitch them on the
    device (copy_rows_kernel): rows whose byte count is odd, a multiple of 2 or of 4, page-locked outputs through
    ``out=``, a strided page-locked source (which keeps the 2-D copy), all against the oracle."""
    import ctypes as C
    lib = _capi.load()
    coder = HaarCoder()
    blocks = []

    def pinned(shape):
        n = int(np.prod(shape))
        p = C.c_void_p()
        _capi.check(lib.wicca_host_alloc(C.byref(p), max(1, n)), "wicca_host_alloc")
        blocks.append(p)
        return np.ctypeslib.as_array((C.c_uint8 * max(1, n)).from_address(p.value)).reshape(shape)

    try:
        rng = np.random.default_rng(8)
        imgs = []
        for (h, w) in [(700, 1001), (640, 1002), (900, 1004), (513, 2047), (300, 299)]:      # the last one is below the flat threshold
            a = pinned((h, w, 3))
            a[:] = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            imgs.append(a)
        wide = pinned((600, 1500, 3))
        wide[:] = rng.integers(0, 256, wide.shape, dtype=np.uint8)
        imgs.append(wide[:, 100:1101])                                                       # strided rows, still page-locked
        depths = [1, 2, 3, 6]
        outs = [[pinned((-(-im.shape[0] // (1 << d)), -(-im.shape[1] // (1 << d)), 3)) for d in depths] for im in imgs]
        for row in outs:
            for o in row:
                o[:] = 0xAB
        got = coder.get_small_copies_batch(imgs, depths, devices=[0], out=outs)
        for im, row, row_out in zip(imgs, got, outs):
            for d, ic, o in zip(depths, row, row_out):
                exp = ho.haar_icon_blocksum(np.ascontiguousarray(im), d)
                assert np.array_equal(ic, exp), (im.shape, d)
                assert np.array_equal(o, exp), (im.shape, d, "page-locked destination")
    finally:
        for p in blocks:
            lib.wicca_host_free(p)
