"""Outside the one-pass kernel's own domain (C == 3, depths 1..6): depths 7, 8 and beyond finished from the level-6
sum plane (haar_tail_kernel), and every other channel count / alignment through the row-streaming kernel
(haar_icon_rows_kernel).  The reference accepts any positive depth (wavelet_coder.py:58-65) and any channel count
cv2.copyMakeBorder takes (data_loader.py:115-117); parity is bit-exact against the NumPy restatement of the float32 path."""
import ctypes as C

import numpy as np
import pytest

from oracle import c_oracle
from oracle import haar_oracle as ho
from tests.golden.make_golden import gen_input
from wicca_b200 import HaarCoder, _capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def coder():
    return HaarCoder()


@pytest.mark.parametrize("border", [1, 0, 2, 3, 4])
def test_depths_7_and_8_rgb_all_borders_ragged(coder, border):
    """Shapes whose depth-8 padding reaches past the depth-6 extents in x, in y, in both, and not at all."""
    for (h, w) in [(517, 771), (64, 64), (65, 63), (255, 257), (256, 512), (700, 130), (1, 1), (300, 1301)]:
        img = gen_input("noise", h + w + border, h, w, 3)
        got = coder.get_small_copies(img, [7, 8, 3, 6], border, 77)
        for d, g in zip([7, 8, 3, 6], got):
            exp = ho.haar_icon_fp32(img, d, border, 77)
            assert g.shape == exp.shape and np.array_equal(g, exp), (h, w, border, d)
        assert np.array_equal(coder.get_small_copy(img, 8, border, 77), ho.haar_icon_fp32(img, 8, border, 77))


def test_depths_1_to_8_one_call_full_size(coder):
    img = gen_input("noise", 5, 6393, 8284, 3)
    ds = [1, 2, 3, 4, 5, 6, 7, 8]
    got = coder.get_small_copies(img, ds)
    exp = c_oracle.haar_icons_multi(img, ds)
    for d, g, e in zip(ds, got, exp):
        assert g.shape == e.shape and np.array_equal(g, e), d
    assert got[7].shape == (25, 33, 3)


@pytest.mark.parametrize("depth", [9, 10, 12])
def test_depths_beyond_8_replay_float32_like_the_reference(coder, depth):
    for (h, w) in [(517, 771), (1500, 2100)]:
        img = gen_input("noise", depth, h, w, 3)
        for border in (1, 4):
            assert np.array_equal(coder.get_small_copy(img, depth, border), ho.haar_icon_fp32(img, depth, border)), (h, w, depth, border)


@pytest.mark.parametrize("channels", [1, 2, 4, 5])
def test_other_channel_counts_every_depth(coder, channels):
    rng = np.random.default_rng(channels)
    for (h, w) in [(256, 512), (517, 771), (33, 1000), (1024, 96)]:
        img = rng.integers(0, 256, (h, w, channels), dtype=np.uint8)
        for d in range(1, 9):
            r = 1 << d
            pad = (h % r) or (w % r)
            if pad and channels in (1, 5):
                continue            # the reference itself fails there (IndexError / cv2.error), covered in test_gpu_robustness
            for border in ((1, 0, 2, 3, 4) if channels == 4 else (1,)):
                got = coder.get_small_copy(img, d, border, 9)
                exp = ho.haar_icon_fp32(img, d, border, 9)
                assert got.shape == exp.shape and np.array_equal(got, exp), (h, w, channels, d, border)


def test_rgba_full_size(coder):
    img = gen_input("noise", 6, 3001, 4003, 4)
    for d in (1, 3, 6, 8):
        assert np.array_equal(coder.get_small_copy(img, d), ho.haar_icon_blocksum(img, d)), d


def test_unaligned_device_pointers_and_pitches():
    """Device-pointer ABI with a source that is neither 16- nor 4-byte aligned and an odd pitch."""
    import torch
    lib = _capi.load()
    h, w = 301, 517
    img = gen_input("noise", 1, h, w, 3)
    pitch = w * 3 + 5
    buf = torch.zeros(h * pitch + 64, dtype=torch.uint8, device="cuda:0")
    view = buf[3:3 + h * pitch].view(h, pitch)
    view[:, : w * 3] = torch.from_numpy(img.reshape(h, w * 3)).cuda()
    for ds in ([1], [4], [7], [2, 8]):
        outs = [torch.full((-(-h // (1 << d)) * (-(-w // (1 << d)) * 3 + 7) + 8,), 0xEE, dtype=torch.uint8, device="cuda:0") for d in ds]
        pitches = [-(-w // (1 << d)) * 3 + 7 for d in ds]
        n = len(ds)
        rc = lib.wicca_haar_icons_multi_dev(view.data_ptr(), h, w, 3, pitch, (C.c_int * n)(*ds), n, 1, 0.0,
                                            (C.c_void_p * n)(*[o.data_ptr() + 1 for o in outs]), (C.c_int64 * n)(*pitches), 0, None)
        _capi.check(rc, "icons_multi_dev")
        torch.cuda.synchronize()
        for d, o, p in zip(ds, outs, pitches):
            oh, ow = -(-h // (1 << d)), -(-w // (1 << d))
            host = o.cpu().numpy()
            got = host[1:1 + oh * p].reshape(oh, p)
            assert np.array_equal(got[:, : ow * 3].reshape(oh, ow, 3), ho.haar_icon_blocksum(img, d)), d
            assert host[0] == 0xEE and (got[:, ow * 3:][:-1] == 0xEE).all()          # nothing outside the rows is touched


def test_plan_with_depths_1_to_8():
    import torch
    from wicca_b200.plan import IconPlan, pitch_bytes
    shapes = [(517, 771), (1300, 900), (64, 64)]
    imgs, tens = [], []
    for k, (h, w) in enumerate(shapes):
        im = gen_input("noise", 50 + k, h, w, 3)
        p = pitch_bytes(w, 3)
        t = torch.zeros((h, p), dtype=torch.uint8, device="cuda:0")
        t[:, : w * 3] = torch.from_numpy(im.reshape(h, w * 3)).cuda()
        imgs.append(im); tens.append(t)
    ds = [1, 2, 3, 4, 5, 6, 7, 8]
    for border in (1, 4):
        plan = IconPlan(0, [t.data_ptr() for t in tens], [s[0] for s in shapes], [s[1] for s in shapes],
                        [t.shape[1] for t in tens], ds, border_type=border)
        assert plan.info()["launches"] <= 2 + 3 * len(shapes)        # one pass (+ strip pre-pass) + the tiny tail launches
        plan.launch(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        for i, im in enumerate(imgs):
            for k, d in enumerate(ds):
                assert np.array_equal(plan.read_icon(i, k), ho.haar_icon_fp32(im, d, border)), (i, d, border)
        plan.close()


def test_concurrent_callers_share_one_coder_across_all_paths():
    """The reference calls one coder instance from a ThreadPoolExecutor (classifying_tools.py:414-418).  Twelve
    threads hammer the three kernel families at once - one-pass (depths <= 6), level-6 plane + tail (depths 7-10) and the
    row-streaming kernel (RGBA, grey) - each call must still equal the oracle (per-call streams, planes and scratch)."""
    from concurrent.futures import ThreadPoolExecutor
    coder = HaarCoder()
    rng = np.random.default_rng(77)
    jobs = []
    for k in range(36):
        c = (3, 3, 4, 1)[k % 4]
        h, w = int(rng.integers(200, 900)), int(rng.integers(200, 1300))
        depths = ([1, 4, 6], [8, 3], [7, 9], [2, 5])[k % 4] if c == 3 else ([3, 6], [8])[k % 2]
        if c == 1:                                   # grey images must not need padding (the reference fails there)
            h, w = h // 256 * 256 + 256, w // 256 * 256 + 256
        jobs.append((rng.integers(0, 256, (h, w, c), dtype=np.uint8), depths, (1, 4, 2)[k % 3]))

    def run(job):
        img, depths, border = job
        got = coder.get_small_copies(img, depths, border, 5)
        return all(np.array_equal(g, ho.haar_icon_fp32(img, d, border, 5)) for g, d in zip(got, depths))

    with ThreadPoolExecutor(12) as ex:
        for rep in range(3):
            assert all(ex.map(run, jobs)), rep
