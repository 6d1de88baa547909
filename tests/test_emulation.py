"""CPU replay of the device code path (wicca_b200/csrc/haar_math.cuh compiled as C++):
every lane of every work item of the one-pass icon kernel, against the oracle."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import haar_oracle as ho
from tests.golden.make_golden import gen_input

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def emul():
    out = ROOT / "tests" / "cpu_emul" / "_build"
    out.mkdir(parents=True, exist_ok=True)
    so = out / "libemul_icon.so"
    src = ROOT / "tests" / "cpu_emul" / "emul_icon.cpp"
    hdrs = [ROOT / "wicca_b200" / "csrc" / n for n in ("haar_math.cuh", "icon_types.h")]
    if not so.exists() or so.stat().st_mtime < max(p.stat().st_mtime for p in [src, *hdrs]):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include",
                        f"-I{ROOT / 'wicca_b200' / 'csrc'}", str(src), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    lib.emul_fused_icons.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_uint, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]

    def run(img, bt, bc, depths, sum6=None):
        h, w, _ = img.shape
        mask, outs, arr = 0, {}, (ctypes.c_void_p * 6)()
        for d in depths:
            mask |= 1 << (d - 1)
            outs[d] = np.full((-(-h // 2 ** d), -(-w // 2 ** d), 3), 0x77, np.uint8)
            arr[d - 1] = outs[d].ctypes.data
        rc = lib.emul_fused_icons(img.ctypes.data, h, w, bt, bc, mask, arr, None if sum6 is None else sum6.ctypes.data)
        assert rc == 0, f"guard bytes overwritten (code {rc})"
        return outs
    return run


SHAPES = [(1, 1), (1, 2), (2, 1), (3, 5), (5, 7), (16, 16), (17, 33), (64, 64), (65, 63), (63, 129), (100, 130),
          (127, 255), (128, 256), (200, 259), (130, 517), (129, 257), (70, 300)]


@pytest.mark.parametrize("border", [1, 0, 2, 3, 4])
def test_emulated_kernel_matches_oracle(emul, border):
    rng = np.random.default_rng(border)
    for (h, w) in SHAPES:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        bc = int(rng.integers(0, 256))
        for depths in ([1, 2, 3, 4, 5, 6], [1], [3], [6], [2, 5]):
            outs = emul(img, border, bc, depths)
            for d in depths:
                assert np.array_equal(outs[d], ho.haar_icon_blocksum(img, d, border, bc)), (h, w, border, depths, d)


@pytest.mark.parametrize("kind", ["full", "trunc", "zeros", "hramp", "vramp"])
def test_emulated_kernel_adversarial(emul, kind):
    img = gen_input(kind, 5, 139, 301, 3)
    outs = emul(img, 1, 0, [1, 2, 3, 4, 5, 6])
    for d, got in outs.items():
        assert np.array_equal(got, ho.haar_icon_blocksum(img, d, 1, 0)), (kind, d)


@pytest.mark.parametrize("border", [1, 0, 2, 3, 4])
def test_emulated_level6_sum_plane(emul, border):
    """The plane of exact 64 x 64 block sums the kernel leaves for depths 7 and 8 (haar_tail_kernel)."""
    rng = np.random.default_rng(100 + border)
    for (h, w) in [(1, 1), (64, 64), (65, 63), (63, 129), (130, 517), (200, 259), (129, 257)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        bc = int(rng.integers(0, 256))
        for depths in ([], [6], [1, 2, 3, 4, 5, 6]):
            s6 = np.full((-(-h // 64), -(-w // 64), 3), 0xDEADBEEF, np.uint32)
            emul(img, border, bc, depths, sum6=s6)
            pad = ho.get_padded_copy(img, 64, border, bc).astype(np.uint32)
            exp = pad.reshape(pad.shape[0] // 64, 64, pad.shape[1] // 64, 64, 3).sum(axis=(1, 3), dtype=np.uint32)
            assert np.array_equal(s6, exp), (h, w, border, depths)
