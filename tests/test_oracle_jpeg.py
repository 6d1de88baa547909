"""CPU tests of the JPEG ingest path (row N2): the oracle restatement of libjpeg-turbo's decoder is pinned on
cv2.imdecode (the reference's own dependency, run live), and the library's host-side Huffman decoder is
checked against the oracle.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import jpeg_oracle as jo  # noqa: E402
from wicca_b200 import _capi  # noqa: E402

SAMPLING = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
            "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440,
            "411": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411}


def photo_like(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(xx / 17.0 + c) + 60 * np.cos(yy / 11.0 - c) for c in range(3)], -1)
    return np.clip(base + rng.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)


def encode(img, quality=90, sampling="420", restart=0, optimize=False, progressive=False):
    params = [cv2.IMWRITE_JPEG_QUALITY, quality]
    if img.ndim == 3:
        params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING[sampling]]
    if restart:
        params += [cv2.IMWRITE_JPEG_RST_INTERVAL, restart]
    if optimize:
        params += [cv2.IMWRITE_JPEG_OPTIMIZE, 1]
    if progressive:
        params += [cv2.IMWRITE_JPEG_PROGRESSIVE, 1]
    ok, enc = cv2.imencode(".jpg", img, params)
    assert ok
    return bytes(enc)


def reference_rgb(data):
    """What the reference's load_image returns (data_loader.py:53-58)."""
    return cv2.cvtColor(cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)


def jpeg_golden():
    """Committed fixtures: JPEG files and what the reference's own load_image returned for them
    (tests/golden/make_golden.py jpeg, run where /root/reference and its OpenCV are importable)."""
    import json
    from pathlib import Path
    z = np.load(Path(__file__).parent / "golden" / "jpeg_golden.npz")
    cases = json.loads(str(z["cases"][0]))
    return [(tuple(c), bytes(z[f"file_{i}"]), z[f"rgb_{i}"]) for i, c in enumerate(cases)]


def test_oracle_matches_reference_goldens():
    for case, data, rgb in jpeg_golden():
        assert np.array_equal(jo.decode_rgb_oriented(data), rgb), case


CASES = [(h, w, q, s, r) for (h, w) in [(8, 8), (1, 1), (2, 3), (17, 33), (37, 53), (70, 31)] for q in (35, 90, 100)
         for s in SAMPLING for r in (0, 3)]


@pytest.mark.parametrize("sampling", list(SAMPLING))
def test_oracle_matches_cv2(sampling):
    rng = np.random.default_rng(5)
    for (h, w, q, s, r) in CASES:
        if s != sampling:
            continue
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if q == 35 else photo_like(rng, h, w)
        data = encode(img, q, s, r, optimize=(q == 100))
        assert np.array_equal(jo.decode_rgb(data), reference_rgb(data)), (h, w, q, s, r)


def test_oracle_grey():
    rng = np.random.default_rng(6)
    grey = photo_like(rng, 40, 23)[:, :, 0]
    for progressive in (False, True):
        data = encode(grey, 85, progressive=progressive)
        assert np.array_equal(jo.decode_rgb(data), reference_rgb(data))


@pytest.mark.parametrize("sampling", ["444", "422", "420", "411"])
def test_oracle_progressive_matches_cv2(sampling):
    rng = np.random.default_rng(9)
    for (h, w) in [(8, 8), (1, 1), (17, 33), (37, 53), (70, 31)]:
        for q, r in ((35, 0), (90, 3), (100, 0)):
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if q == 35 else photo_like(rng, h, w)
            data = encode(img, q, sampling, r, progressive=True)
            assert np.array_equal(jo.decode_rgb(data), reference_rgb(data)), (h, w, q, sampling, r)


@pytest.mark.parametrize("sampling", ["444", "420"])
def test_host_progressive_decoder_matches_oracle(sampling):
    rng = np.random.default_rng(10)
    for (h, w, q, r) in [(8, 8, 90, 0), (37, 53, 35, 0), (70, 31, 90, 3), (130, 97, 100, 0), (96, 64, 75, 5)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if q == 35 else photo_like(rng, h, w)
        data = encode(img, q, sampling, r, progressive=True)
        dst, bw, bh, _ = host_coefficients(data)
        exp = jo.decode_coefficients_multiscan(data)
        off = 0
        for k, cf in enumerate(exp["coefs"]):
            assert (bh[k], bw[k]) == cf.shape[:2]
            assert np.array_equal(dst[off:off + cf.size].reshape(cf.shape), cf), (h, w, q, sampling, r, k)
            off += cf.size


def emul_jpeg():
    """tests/cpu_emul/jpeg_entropy_host.cpp (host entropy decoders: the checker of the GPU Huffman decoder) built with
    the product's marker parser compiled as C++.  Never part of libwicca_b200.so."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    out = root / "tests" / "cpu_emul" / "_build"
    out.mkdir(parents=True, exist_ok=True)
    so = out / "libemul_jpeg.so"
    srcs = [root / "tests" / "cpu_emul" / "jpeg_entropy_host.cpp", root / "wicca_b200" / "csrc" / "jpeg_host.cu"]
    deps = srcs + [root / "wicca_b200" / "csrc" / "jpeg_host.h"]
    if not so.exists() or so.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include",
                        f"-I{root / 'wicca_b200' / 'csrc'}", str(srcs[0]), "-x", "c++", str(srcs[1]), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    lib.emul_jpeg_coeff_count.restype = C.c_longlong
    lib.emul_jpeg_coeff_count.argtypes = [C.c_char_p, C.c_size_t]
    lib.emul_jpeg_decode_coeffs.restype = C.c_int
    lib.emul_jpeg_decode_coeffs.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_char_p]
    return lib


def host_coefficients(data):
    lib = emul_jpeg()
    n = lib.emul_jpeg_coeff_count(data, len(data))
    assert n > 0, n
    dst = np.empty(n, np.int16)
    bw, bh = (C.c_int * 3)(), (C.c_int * 3)()
    qt = np.empty(192, np.uint16)
    why = C.create_string_buffer(256)
    rc = lib.emul_jpeg_decode_coeffs(data, len(data), dst.ctypes.data, n, bw, bh, qt.ctypes.data, why)
    assert rc == 0, why.value
    return dst, list(bw), list(bh), qt


@pytest.mark.parametrize("sampling", list(SAMPLING))
def test_host_huffman_decoder_matches_oracle(sampling):
    """The product's marker parser + the host Huffman decoder that checks the GPU one (tests/cpu_emul), against the
    oracle's, coefficient by coefficient."""
    rng = np.random.default_rng(7)
    for (h, w, q, s, r) in CASES + [(130, 97, 90, sampling, 5)]:
        if s != sampling:
            continue
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if q == 35 else photo_like(rng, h, w)
        data = encode(img, q, s, r, optimize=(q == 100))
        dst, bw, bh, qt = host_coefficients(data)
        exp = jo.decode_coefficients(data)
        off = 0
        for k, cf in enumerate(exp["coefs"]):
            assert (bh[k], bw[k]) == cf.shape[:2]
            assert np.array_equal(dst[off:off + cf.size].reshape(cf.shape), cf), (h, w, q, s, r, k)
            assert np.array_equal(qt[64 * k:64 * k + 64], exp["qt"][k])
            off += cf.size
        assert off == dst.size


def test_probe_and_unsupported_flavours():
    from wicca_b200 import UnsupportedImageError, jpeg_info
    rng = np.random.default_rng(8)
    img = photo_like(rng, 33, 47)
    info = jpeg_info(encode(img, 90, "420"))
    assert info == {"height": 33, "width": 47, "components": 3, "h_max": 2, "v_max": 2}
    assert jpeg_info(encode(img[:, :, 0], 90))["components"] == 1
    with pytest.raises(UnsupportedImageError):                          # progressive: not decoded on the GPU, no CPU fallback
        jpeg_info(encode(img, progressive=True))
    ok, png = cv2.imencode(".png", img)
    with pytest.raises(UnsupportedImageError):
        jpeg_info(bytes(png))
    with pytest.raises(ValueError):
        jpeg_info(encode(img)[:40])                     # truncated header
    # EXIF orientation 6: cv2 rotates the image, and the probe reports the rotated size
    rotated = with_exif_orientation(encode(img), 6)
    assert reference_rgb(rotated).shape == (47, 33, 3)
    info = jpeg_info(rotated)
    assert (info["height"], info["width"]) == (47, 33)


def with_exif_orientation(data, orientation):
    """Splice an Exif APP1 segment carrying tag 0x0112 = orientation right after SOI."""
    exif = (b"Exif\x00\x00MM\x00\x2a\x00\x00\x00\x08\x00\x01\x01\x12\x00\x03\x00\x00\x00\x01"
            + bytes([0, orientation]) + b"\x00\x00\x00\x00\x00\x00")
    seg = b"\xff\xe1" + (len(exif) + 2).to_bytes(2, "big") + exif
    return data[:2] + seg + data[2:]


def test_host_decoder_survives_corrupt_files():
    """Ingest reads untrusted files: mutated headers and scans must end in an error code or a decode, never a crash
    (the same driver ran 24,000 mutations under ASan/UBSan during development)."""
    lib = emul_jpeg()
    rng = np.random.default_rng(99)
    base = bytearray(encode(photo_like(rng, 61, 83), 85, "420", restart=4))
    outcomes = set()
    for it in range(600):
        d = bytearray(base)
        kind = it % 4
        for _ in range(1 + int(rng.integers(0, 6))):
            pos = int(rng.integers(0, min(len(d), 700) if kind == 0 else len(d)))
            d[pos] = 0xFF if kind == 3 else int(rng.integers(0, 256))
        if kind == 2:
            d = d[:1 + int(rng.integers(0, len(d)))]
        data = bytes(d)
        n = lib.emul_jpeg_coeff_count(data, len(data))
        assert n == _capi.load().wicca_jpeg_coeff_count(data, len(data))          # same parser on both sides
        if n <= 0:
            assert n in (_capi.EINVAL, _capi.EUNSUPPORTED), n
            outcomes.add(int(n))
            continue
        if n > 1 << 24:
            continue
        dst = np.empty(n, np.int16)
        rc = lib.emul_jpeg_decode_coeffs(data, len(data), dst.ctypes.data, n, None, None, None, None)
        assert rc in (0, _capi.EINVAL), rc
        outcomes.add(int(rc))
    assert 0 in outcomes and _capi.EINVAL in outcomes
