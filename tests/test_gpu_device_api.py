"""Device-pointer entry points: wicca_haar_icons_multi_dev with guard bytes around caller-owned
buffers (the pool has no compute-sanitizer), and the plan's fused resize/normalise epilogue."""
import ctypes as C

import numpy as np
import pytest

from oracle import haar_oracle as ho
from oracle import resize_oracle as ro
from tests.golden.make_golden import gen_input
from wicca_b200 import _capi
from wicca_b200.plan import IconPlan, pitch_bytes, to_device_pitched

pytestmark = pytest.mark.gpu

GUARD = 4096


def _run_dev(img, depths, border, bconst, align_ok=True):
    import torch
    lib = _capi.load()
    h, w, c = img.shape
    src = to_device_pitched(img)
    bufs, ptrs, pitches, shapes = [], [], [], []
    for d in depths:
        oh, ow = -(-h // (1 << d)), -(-w // (1 << d))
        pitch = (ow * c + 127) // 128 * 128 if (align_ok and d <= 8) else ow * c
        t = torch.full((GUARD + oh * pitch + GUARD,), 0xEE, dtype=torch.uint8, device="cuda:0")
        bufs.append(t); shapes.append((oh, ow, pitch))
        ptrs.append(t.data_ptr() + GUARD); pitches.append(pitch)
    n = len(depths)
    rc = lib.wicca_haar_icons_multi_dev(src.data_ptr(), h, w, c, src.shape[1], (C.c_int * n)(*depths), n, border,
                                        float(bconst), (C.c_void_p * n)(*ptrs), (C.c_int64 * n)(*pitches), 0,
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _capi.check(rc, "wicca_haar_icons_multi_dev")
    torch.cuda.synchronize()
    outs = []
    for t, (oh, ow, pitch) in zip(bufs, shapes):
        host = t.cpu().numpy()
        assert (host[:GUARD] == 0xEE).all() and (host[-GUARD:] == 0xEE).all(), "guard bytes overwritten"
        body = host[GUARD:GUARD + oh * pitch].reshape(oh, pitch)
        # TMA store clips at 16-byte granularity: the bytes between w*C and the next 16-byte boundary of a
        # row (always inside the row's pitch) may be overwritten, nothing beyond (documented in the header)
        pad = body[:, (ow * c + 15) // 16 * 16:]
        if not (pad == 0xEE).all():
            bad = np.argwhere(pad != 0xEE)
            raise AssertionError(f"row padding overwritten: depth {depths[len(outs)]} icon {oh}x{ow} pitch {pitch}: "
                                 f"{len(bad)} bytes, first rows/cols {bad[:6].tolist()}, values "
                                 f"{[int(pad[r, q]) for r, q in bad[:6]]}, img {h}x{w}")
        outs.append(body[:, :ow * c].reshape(oh, ow, c).copy())
    return outs


@pytest.mark.parametrize("border", [1, 0, 2, 3, 4])
def test_dev_api_with_guards(border):
    for (h, w) in [(64, 128), (777, 1301), (65, 129), (1000, 8284 // 4), (333, 17)]:
        img = gen_input("noise", h + w + border, h, w, 3)
        depths = [1, 2, 3, 4, 5, 6]
        outs = _run_dev(img, depths, border, 41)
        for d, o in zip(depths, outs):
            assert np.array_equal(o, ho.haar_icon_blocksum(img, d, border, 41)), (h, w, d, border)


def test_dev_api_general_path_and_deep_levels():
    img = gen_input("noise", 12, 300, 520, 4)                   # C = 4 -> general kernel
    outs = _run_dev(img, [1, 3], 2, 0)
    for d, o in zip([1, 3], outs):
        assert np.array_equal(o, ho.haar_icon_fp32(img, d, 2, 0))
    img = gen_input("noise", 13, 700, 1100, 3)
    outs = _run_dev(img, [7, 9], 1, 0, align_ok=False)          # depth > 6 -> general kernel, depth 9 -> fp32 levels
    assert np.array_equal(outs[0], ho.haar_icon_blocksum(img, 7))
    assert np.array_equal(outs[1], ho.haar_icon_fp32(img, 9))


def test_plan_fused_epilogue_config4():
    """configs[3]: icons -> INTER_AREA 224 / 331 -> preprocess_input, without leaving the device."""
    import torch
    rng = np.random.default_rng(21)
    shapes = [(1599, 2071), (1600, 2048), (801, 1037), (1234, 999)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    dev = [to_device_pitched(im) for im in imgs]
    depths = [1, 2, 3]
    plan = IconPlan(0, [t.data_ptr() for t in dev], [s[0] for s in shapes], [s[1] for s in shapes],
                    [t.shape[1] for t in dev], depths)
    st = torch.cuda.current_stream().cuda_stream
    plan.launch(st)
    for target in (224, 331):
        for k, d in enumerate(depths):
            for mode_name, mode in (("tf", 1), ("caffe", 2), ("torch", 3), ("identity", 0)):
                out = torch.empty((len(imgs), target, target, 3), dtype=torch.float32, device="cuda:0")
                out8 = torch.empty((len(imgs), target, target, 3), dtype=torch.uint8, device="cuda:0")
                plan.resize_norm(k, target, target, mode, out.data_ptr(), out8.data_ptr(), st)
                torch.cuda.synchronize()
                exp8 = np.stack([ro.resize_area(ho.haar_icon_blocksum(im, d), target, target) for im in imgs])
                assert np.array_equal(out8.cpu().numpy(), exp8), (target, d)
                assert np.array_equal(out.cpu().numpy(), ro.preprocess_input(exp8, mode_name)), (target, d, mode_name)
    plan.close()


def test_torch_bridge_cuda_tensor():
    import torch
    from wicca_b200.torch_bridge import icons_from_cuda_tensor
    for (h, w) in [(640, 1024), (777, 1301)]:                 # aligned rows (direct) and unaligned rows (staged)
        img = gen_input("noise", h * 3 + w, h, w, 3)
        t = torch.from_numpy(img).cuda()
        outs = icons_from_cuda_tensor(t, [1, 3, 6, 7], border_type=4)
        torch.cuda.synchronize()
        for d, o in zip([1, 3, 6, 7], outs):
            assert o.shape == (-(-h // 2 ** d), -(-w // 2 ** d), 3)
            assert np.array_equal(o.cpu().numpy(), ho.haar_icon_blocksum(img, d, 4)), (h, w, d)
