"""Epilogue rows A5/A6: icon -> INTER_AREA resize -> preprocess_input, against the oracle
(which is pinned on cv2 outputs) and, when cv2 is importable, against cv2 itself."""
import numpy as np
import pytest

from oracle import resize_oracle as ro
from tests.golden.make_golden import gen_input
from wicca_b200 import HaarCoder

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def coder():
    return HaarCoder()


def test_resize_goldens_bit_exact(coder, resize_golden):
    cases, outs = resize_golden
    for (seed, sh, sw, dw, dh), exp in zip(cases, outs):
        src = gen_input("noise", 1000 + seed, sh, sw, 3)
        f32, u8 = coder.icons_to_batch([src], (dw, dh), "identity", return_uint8=True)
        assert u8.shape == (1, dh, dw, 3) and np.array_equal(u8[0], exp), (seed, sh, sw, dw, dh)
        assert np.array_equal(f32[0], exp.astype(np.float32))


@pytest.mark.parametrize("target", [224, 331])
def test_headline_icon_sizes_all_regimes(coder, target):
    """Icon sizes of the (6393, 8284) image at depths 2..6 -> 224 / 331 (depth 1 is in the slow test)."""
    sizes = [(1599, 2071), (800, 1036), (400, 518), (200, 259), (100, 130)]
    icons = [gen_input("noise", 50 + i, h, w, 3) for i, (h, w) in enumerate(sizes)]
    f32, u8 = coder.icons_to_batch(icons, (target, target), "tf", return_uint8=True)
    assert f32.shape == (len(sizes), target, target, 3) and f32.dtype == np.float32
    for i, ic in enumerate(icons):
        exp = ro.resize_area(ic, target, target)
        assert np.array_equal(u8[i], exp), (sizes[i], target, ro.regime(ic.shape[1], ic.shape[0], target, target))
    exp_f = ro.preprocess_input(u8, "tf")
    assert np.allclose(f32, exp_f, rtol=1e-5, atol=0) and np.array_equal(f32, exp_f)


@pytest.mark.parametrize("mode", ["identity", "tf", "caffe", "torch"])
def test_preprocess_modes(coder, mode):
    icons = [gen_input("noise", 70 + i, 300 + 10 * i, 410 + 7 * i, 3) for i in range(4)]
    f32, u8 = coder.icons_to_batch(icons, (240, 240), mode, return_uint8=True)
    exp = ro.preprocess_input(u8, mode)
    # tolerance stated by north_star: <= 1e-5 relative; the implementation is in fact bit-exact
    assert np.allclose(f32, exp, rtol=1e-5, atol=1e-6)
    assert np.array_equal(f32, exp)


def test_live_cv2_when_available(coder):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    icons = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in [(448, 448), (672, 672), (224, 224), (57, 91),
                                                                           (3197 // 2, 4142 // 2), (299, 299)]]
    for t in (224, 299):
        _, u8 = coder.icons_to_batch(icons, (t, t), "identity", return_uint8=True)
        for i, ic in enumerate(icons):
            assert np.array_equal(u8[i], cv2.resize(ic, (t, t), interpolation=cv2.INTER_AREA)), (ic.shape, t)


def test_end_to_end_icon_then_batch(coder):
    """configs[3]: HaarCoder icons -> resize -> normalise, batch of images."""
    from oracle import haar_oracle as ho
    imgs = [gen_input("noise", 90 + i, 1200 + 13 * i, 1600 + 7 * i, 3) for i in range(6)]
    for depth in (1, 3):
        icons = [coder.get_small_copy(im, depth) for im in imgs]
        batch = coder.icons_to_batch(icons, (224, 224), "tf")
        exp = ro.preprocess_input(np.stack([ro.resize_area(ho.haar_icon_blocksum(im, depth), 224, 224) for im in imgs]), "tf")
        assert np.array_equal(batch, exp)


def test_source_image_branch_n1(coder):
    """Row N1: the reference also resizes the full source image (classifying_tools.py:315); the same
    kernel handles it (scale ~ 9-14 here), host and device-resident entry points."""
    import ctypes as C
    import torch
    from wicca_b200 import _capi
    from wicca_b200.plan import to_device_pitched
    imgs = [gen_input("noise", 600 + i, 2000 + 31 * i, 3100 - 17 * i, 3) for i in range(3)]
    exp = np.stack([ro.resize_area(im, 224, 224) for im in imgs])
    f32, u8 = coder.icons_to_batch(imgs, (224, 224), "caffe", return_uint8=True)
    assert np.array_equal(u8, exp) and np.array_equal(f32, ro.preprocess_input(exp, "caffe"))
    dev = [to_device_pitched(im) for im in imgs]
    out = torch.empty((3, 224, 224, 3), dtype=torch.float32, device="cuda:0")
    out8 = torch.empty((3, 224, 224, 3), dtype=torch.uint8, device="cuda:0")
    n = len(imgs)
    rc = _capi.load().wicca_resize_norm_dev((C.c_void_p * n)(*[t.data_ptr() for t in dev]), (C.c_int * n)(*[im.shape[0] for im in imgs]),
                                            (C.c_int * n)(*[im.shape[1] for im in imgs]), (C.c_int64 * n)(*[t.shape[1] for t in dev]),
                                            n, 224, 224, 1, out.data_ptr(), out8.data_ptr(), 0,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _capi.check(rc, "wicca_resize_norm_dev")
    torch.cuda.synchronize()
    assert np.array_equal(out8.cpu().numpy(), exp)
    assert np.array_equal(out.cpu().numpy(), ro.preprocess_input(exp, "tf"))
    # The tap tables of a resident batch are cached on the device (keyed by pointers and geometry, never by content):
    # new pixels behind the same pointers, another target in between, a second stream - every call must see the data.
    side = torch.cuda.Stream()
    for rep, target in enumerate((224, 331, 224, 224)):
        new = [gen_input("noise", 900 + 10 * rep + i, *im.shape) for i, im in enumerate(imgs)]
        for t, im in zip(dev, new):
            t[:, : im.shape[1] * 3].copy_(torch.from_numpy(im.reshape(im.shape[0], -1)))
        torch.cuda.synchronize()
        o8 = torch.empty((3, target, target, 3), dtype=torch.uint8, device="cuda:0")
        of = torch.empty((3, target, target, 3), dtype=torch.float32, device="cuda:0")
        st = side if rep % 2 else torch.cuda.current_stream()
        rc = _capi.load().wicca_resize_norm_dev((C.c_void_p * n)(*[t.data_ptr() for t in dev]), (C.c_int * n)(*[im.shape[0] for im in imgs]),
                                                (C.c_int * n)(*[im.shape[1] for im in imgs]), (C.c_int64 * n)(*[t.shape[1] for t in dev]),
                                                n, target, target, 0, of.data_ptr(), o8.data_ptr(), 0, C.c_void_p(st.cuda_stream))
        _capi.check(rc, "wicca_resize_norm_dev")
        torch.cuda.synchronize()
        assert np.array_equal(o8.cpu().numpy(), np.stack([ro.resize_area(im, target, target) for im in new])), (rep, target)
    assert _capi.load().wicca_shutdown() == 0           # drops the cached tables; the next call rebuilds them


@pytest.mark.parametrize("depth,shape,mode", [(2, (224, 224), "tf"), (3, (331, 331), "caffe"), (5, (224, 224), "torch"),
                                              (1, (240, 299), "identity")])
def test_classifier_batches_one_call(coder, depth, shape, mode):
    """The fused _get_img_batch + preprocess_input call: both branches, ragged images, all regimes."""
    from oracle import haar_oracle as ho
    rng = np.random.default_rng(depth)
    imgs = [gen_input("noise", 700 + i, 1500 + int(rng.integers(-200, 200)), 2100 + int(rng.integers(-300, 300)), 3)
            for i in range(5)]
    batch_images, batch_icons = coder.classifier_batches(imgs, depth, shape, mode)
    ow, oh = shape
    assert batch_images.shape == batch_icons.shape == (5, oh, ow, 3) and batch_icons.dtype == np.float32
    exp_icons = ro.preprocess_input(np.stack([ro.resize_area(ho.haar_icon_blocksum(im, depth), ow, oh) for im in imgs]), mode)
    exp_images = ro.preprocess_input(np.stack([ro.resize_area(im, ow, oh) for im in imgs]), mode)
    assert np.array_equal(batch_icons, exp_icons)
    assert np.array_equal(batch_images, exp_images)
    only_icons = coder.classifier_batches(imgs, depth, shape, mode, with_source=False)
    assert only_icons[0] is None and np.array_equal(only_icons[1], exp_icons)


def test_classifier_batches_multi_one_upload(coder):
    """Row N3: every (target, depth) batch of the reference's classifier x depth loops from one upload per
    image, bit-identical to the oracle and to the one-target call."""
    from oracle import haar_oracle as ho
    rng = np.random.default_rng(11)
    imgs = [gen_input("noise", 900 + i, 1300 + int(rng.integers(-150, 150)), 1900 + int(rng.integers(-250, 250)), 3)
            for i in range(4)]
    depths = [2, 3, 4, 5, 6]
    targets = [((224, 224), "tf"), ((224, 224), "caffe"), ((299, 299), "tf"), ((240, 240), "identity"), ((331, 331), "torch")]
    out = coder.classifier_batches_multi(imgs, depths, targets)
    assert len(out) == len(targets)
    icons = {d: [ho.haar_icon_blocksum(im, d) for im in imgs] for d in depths}
    for (shape, mode), (batch_images, by_depth) in zip(targets, out):
        ow, oh = shape
        exp_images = ro.preprocess_input(np.stack([ro.resize_area(im, ow, oh) for im in imgs]), mode)
        assert batch_images.shape == (4, oh, ow, 3) and np.array_equal(batch_images, exp_images)
        assert sorted(by_depth) == depths
        for d in depths:
            exp = ro.preprocess_input(np.stack([ro.resize_area(ic, ow, oh) for ic in icons[d]]), mode)
            assert np.array_equal(by_depth[d], exp), (shape, mode, d)
    one_images, one_icons = coder.classifier_batches(imgs, 3, (299, 299), "tf")
    assert np.array_equal(one_images, out[2][0]) and np.array_equal(one_icons, out[2][1][3])
    no_src = coder.classifier_batches_multi(imgs, [2], [((224, 224), "tf")], with_source=False)
    assert no_src[0][0] is None and np.array_equal(no_src[0][1][2], out[0][1][2])


def test_classifier_batches_multi_rejects_bad_arguments(coder):
    img = gen_input("noise", 1, 64, 64, 3)
    with pytest.raises(ValueError):
        coder.classifier_batches_multi([img], [], [((224, 224), "tf")])
    with pytest.raises(ValueError):
        coder.classifier_batches_multi([img], [2, 2], [((224, 224), "tf")])
    with pytest.raises(ValueError):
        coder.classifier_batches_multi([img], [2], [((224, 224), "keras")])
    with pytest.raises(ValueError):
        coder.classifier_batches_multi([img], [0], [((224, 224), "tf")])
    with pytest.raises(ValueError):
        coder.classifier_batches_multi([img], [2], [])


@pytest.mark.parametrize("target", [64, 224, 299, 331, 500])
def test_area_tap_run_shapes_sweep(coder, target):
    """The row kernel walks a tap run as [first group] + whole groups + [whole or half group], decided per warp, with one
    of three register-capped instantiations by CTA size: sweep the x scale from just above 1 (runs of two or three taps)
    to 21 (runs of 23) so that every tail form, mixed warps included, and every instantiation meets the oracle."""
    rng = np.random.default_rng(target)
    widths = sorted({int(target * s) + int(rng.integers(0, 7)) for s in np.linspace(1.02, 21.0, 36)})
    icons = [rng.integers(0, 256, (int(rng.integers(40, 90)), w, 3), dtype=np.uint8) for w in widths]
    _, u8 = coder.icons_to_batch(icons, (target, 17), "identity", return_uint8=True)
    for i, ic in enumerate(icons):
        assert np.array_equal(u8[i], ro.resize_area(ic, target, 17)), (ic.shape, target)


def test_area_rows_unaligned_pitch_device_api(coder):
    """Rows that cannot be bulk-copied (pitch and base not multiples of 16 bytes) take the cooperative copy path."""
    import ctypes as C
    import torch
    from wicca_b200 import _capi
    rng = np.random.default_rng(5)
    H, W, pitch = 301, 1777, 1777 * 3 + 5
    imgs = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(2)]
    flat = torch.zeros(2 * H * pitch + 3, dtype=torch.uint8, device="cuda:0")
    ptrs = []
    for k, im in enumerate(imgs):
        view = flat[3 + k * H * pitch: 3 + (k + 1) * H * pitch].view(H, pitch)
        view[:, : W * 3].copy_(torch.from_numpy(im.reshape(H, -1)))
        ptrs.append(view.data_ptr())
    for target in (224, 331):
        o8 = torch.empty((2, target, target, 3), dtype=torch.uint8, device="cuda:0")
        of = torch.empty((2, target, target, 3), dtype=torch.float32, device="cuda:0")
        rc = _capi.load().wicca_resize_norm_dev((C.c_void_p * 2)(*ptrs), (C.c_int * 2)(H, H), (C.c_int * 2)(W, W), (C.c_int64 * 2)(pitch, pitch),
                                                2, target, target, 0, of.data_ptr(), o8.data_ptr(), 0,
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _capi.check(rc, "wicca_resize_norm_dev")
        torch.cuda.synchronize()
        assert np.array_equal(o8.cpu().numpy(), np.stack([ro.resize_area(im, target, target) for im in imgs])), target
