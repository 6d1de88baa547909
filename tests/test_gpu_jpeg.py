"""GPU parity tests of the JPEG ingest path (row N2): the library's load_image / decode_jpeg against
cv2.imdecode + BGR2RGB run live (what the reference's load_image does, data_loader.py:53-58), bit for bit."""
import ctypes as C

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from tests.test_oracle_jpeg import SAMPLING, encode, photo_like, reference_rgb, with_exif_orientation  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sampling", list(SAMPLING))
def test_decode_matches_cv2(sampling):
    from wicca_b200 import decode_jpeg
    rng = np.random.default_rng(21)
    for (h, w) in [(1, 1), (2, 3), (8, 8), (17, 33), (37, 53), (70, 31), (255, 257), (600, 401)]:
        for q, restart in ((35, 0), (90, 7), (100, 0)):
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if q == 35 else photo_like(rng, h, w)
            data = encode(img, q, sampling, restart, optimize=(q == 100))
            got = decode_jpeg(data)
            assert got.dtype == np.uint8 and np.array_equal(got, reference_rgb(data)), (h, w, q, sampling, restart)


def test_decode_grey_and_flat_images():
    from wicca_b200 import decode_jpeg
    rng = np.random.default_rng(22)
    grey = photo_like(rng, 123, 77)[:, :, 0]
    data = encode(grey, 80)
    assert np.array_equal(decode_jpeg(data), reference_rgb(data))
    for value in (0, 255, 128):
        flat = np.full((64, 48, 3), value, np.uint8)
        data = encode(flat, 90)
        assert np.array_equal(decode_jpeg(data), reference_rgb(data))


def test_large_photo_and_oracle_agree():
    """A 12 MP 4:2:0 file: library == cv2; and on a crop-sized file library == oracle restatement too."""
    from oracle import jpeg_oracle as jo
    from wicca_b200 import decode_jpeg
    rng = np.random.default_rng(23)
    big = photo_like(rng, 3000, 4000)
    data = encode(big, 92, "420")
    timing = {}
    got = decode_jpeg(data, timing=timing)
    assert np.array_equal(got, reference_rgb(data))
    assert timing["host_decode_ms"] > 0 and timing["kernel_ms"] > 0
    small = encode(photo_like(rng, 90, 150), 75, "422", 4)
    assert np.array_equal(decode_jpeg(small), jo.decode_rgb(small))


def test_load_image_and_icons_from_files(tmp_path):
    from oracle import haar_oracle as ho
    from wicca_b200 import HaarCoder, icons_from_jpeg, icons_from_jpeg_files, load_image
    rng = np.random.default_rng(24)
    paths, refs = [], []
    for i, (h, w, s) in enumerate([(700, 900, "420"), (513, 1025, "422"), (640, 480, "444"), (333, 777, "420"), (1200, 800, "420")]):
        data = encode(photo_like(rng, h, w), 88, s, restart=(i % 2) * 11)
        p = tmp_path / f"img{i}.jpg"
        p.write_bytes(data)
        paths.append(str(p))
        refs.append(reference_rgb(data))
    img = load_image(paths[0])
    assert np.array_equal(img, refs[0])
    assert np.array_equal(img, cv2.cvtColor(cv2.imread(paths[0]), cv2.COLOR_BGR2RGB))
    depths = [1, 2, 3, 5]
    one = icons_from_jpeg(open(paths[1], "rb").read(), depths)
    for d, icon in zip(depths, one):
        assert np.array_equal(icon, ho.haar_icon_blocksum(refs[1], d))
        assert np.array_equal(icon, HaarCoder().get_small_copy(refs[1], d))
    per_file = icons_from_jpeg_files(paths, depths, threads=3)
    assert len(per_file) == len(paths)
    for ref, icons in zip(refs, per_file):
        for d, icon in zip(depths, icons):
            assert np.array_equal(icon, ho.haar_icon_blocksum(ref, d))
    with pytest.raises(ValueError):
        load_image("")


def test_decode_to_device_buffer():
    torch = pytest.importorskip("torch")
    from wicca_b200 import _capi, jpeg_info
    from wicca_b200.plan import pitch_bytes
    rng = np.random.default_rng(25)
    data = encode(photo_like(rng, 301, 523), 90, "420")
    info = jpeg_info(data)
    pitch = pitch_bytes(info["width"], 3)
    buf = torch.full((info["height"] + 1, pitch), 0xAB, dtype=torch.uint8, device="cuda:0")
    rc = _capi.load().wicca_jpeg_decode_dev(data, len(data), buf.data_ptr(), pitch, 0,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _capi.check(rc, "wicca_jpeg_decode_dev")
    host = buf.cpu().numpy()
    assert np.array_equal(host[:info["height"], :info["width"] * 3].reshape(info["height"], info["width"], 3), reference_rgb(data))
    assert (host[info["height"]] == 0xAB).all()                                  # the guard row is untouched
    assert (host[:info["height"], (info["width"] * 3 + 3) // 4 * 4:] == 0xAB).all()  # and so is the row padding


@pytest.mark.parametrize("sampling", ["444", "420"])
def test_progressive_files_are_refused_not_decoded_on_the_cpu(sampling):
    """Progressive scans depend on each other and are not decoded on the GPU; the product has no host entropy decoder
    behind it (no CPU fallback), so these files are reported, never decoded approximately."""
    from wicca_b200 import UnsupportedImageError, decode_jpeg, icons_from_jpeg
    rng = np.random.default_rng(27)
    data = encode(photo_like(rng, 255, 257), 90, sampling, 0, progressive=True)
    with pytest.raises(UnsupportedImageError):
        decode_jpeg(data)
    with pytest.raises(UnsupportedImageError):
        icons_from_jpeg(data, [2])


def test_unsupported_files_fail_loudly():
    from wicca_b200 import UnsupportedImageError, decode_jpeg
    rng = np.random.default_rng(26)
    ok, png = cv2.imencode(".png", photo_like(rng, 32, 32))
    with pytest.raises(UnsupportedImageError):
        decode_jpeg(bytes(png))
    with pytest.raises(ValueError):
        decode_jpeg(encode(photo_like(rng, 64, 64))[:300])                      # scan cut short after the header


@pytest.mark.parametrize("sampling", list(SAMPLING))
def test_gpu_huffman_decoder_matches_host_decoder(sampling):
    """The self-synchronising GPU Huffman decoder against the host decoder, coefficient by coefficient."""
    from tests.test_oracle_jpeg import host_coefficients
    from wicca_b200 import _capi
    lib = _capi.load()
    rng = np.random.default_rng(31)
    for (h, w, q, restart) in [(8, 8, 90, 0), (40, 56, 35, 0), (257, 511, 75, 0), (600, 800, 92, 0), (1500, 2100, 85, 0),
                               (333, 1001, 100, 0), (40, 56, 90, 1), (257, 511, 35, 3), (600, 800, 92, 50), (1500, 2100, 85, 132),
                               (333, 1001, 100, 7), (64, 64, 90, 1000)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if q == 35 else photo_like(rng, h, w)
        data = encode(img, q, sampling, restart, optimize=(q == 100))
        exp, _, _, _ = host_coefficients(data)
        got = np.empty_like(exp)
        passes = C.c_int()
        _capi.check(lib.wicca_jpeg_decode_coeffs_gpu(data, len(data), got.ctypes.data, got.size, 0, C.byref(passes)), "coeffs_gpu")
        assert np.array_equal(got, exp), (h, w, q, sampling, restart, passes.value)


@pytest.mark.parametrize("orientation", [1, 2, 3, 4, 5, 6, 7, 8])
def test_exif_orientation_is_applied_like_cv2(orientation):
    from oracle import haar_oracle as ho
    from wicca_b200 import decode_jpeg, icons_from_jpeg
    rng = np.random.default_rng(40 + orientation)
    for (h, w, s) in [(37, 53, "420"), (301, 190, "422"), (64, 64, "444")]:
        data = with_exif_orientation(encode(photo_like(rng, h, w), 90, s), orientation)
        ref = reference_rgb(data)
        assert ref.shape == ((w, h, 3) if orientation >= 5 else (h, w, 3))
        assert np.array_equal(decode_jpeg(data), ref), (orientation, h, w, s)
    icon = icons_from_jpeg(data, [2])[0]
    assert np.array_equal(icon, ho.haar_icon_blocksum(ref, 2))


def test_classifier_batches_from_files(tmp_path):
    """File paths -> every (target, depth) batch, equal to the same call on the cv2-decoded images and to the oracle."""
    from oracle import haar_oracle as ho
    from oracle import resize_oracle as ro
    from wicca_b200 import HaarCoder
    rng = np.random.default_rng(50)
    paths, refs = [], []
    for i, (h, w, s, o) in enumerate([(900, 1300, "420", 1), (1111, 801, "422", 6), (640, 960, "444", 1), (1001, 1500, "420", 3),
                                      (777, 1234, "420", 1)]):
        data = with_exif_orientation(encode(photo_like(rng, h, w), 90, s), o) if o != 1 else encode(photo_like(rng, h, w), 90, s)
        p = tmp_path / f"c{i}.jpg"
        p.write_bytes(data)
        paths.append(str(p))
        refs.append(reference_rgb(data))
    coder = HaarCoder()
    depths = [2, 4]
    targets = [((224, 224), "tf"), ((299, 299), "caffe")]
    got = coder.classifier_batches_multi_from_files(paths, depths, targets)
    exp = coder.classifier_batches_multi(refs, depths, targets)
    for (gi, gd), (ei, ed) in zip(got, exp):
        assert np.array_equal(gi, ei)
        for d in depths:
            assert np.array_equal(gd[d], ed[d])
    icon = ho.haar_icon_blocksum(refs[1], 2)
    assert np.array_equal(got[0][1][2][1], ro.preprocess_input(ro.resize_area(icon, 224, 224)[None], "tf")[0])


def test_decode_matches_reference_goldens():
    """Committed fixtures from the reference's own load_image (no OpenCV needed at test time)."""
    from tests.test_oracle_jpeg import jpeg_golden
    from wicca_b200 import decode_jpeg
    from wicca_b200 import UnsupportedImageError
    n_ok = 0
    for case, data, rgb in jpeg_golden():
        if len(case) > 8 and case[8]:                    # progressive fixture: refused, never decoded on the CPU
            with pytest.raises(UnsupportedImageError):
                decode_jpeg(data)
            continue
        assert np.array_equal(decode_jpeg(data), rgb), case
        n_ok += 1
    assert n_ok >= 14


def test_gpu_decoder_survives_corrupt_scans():
    """Corrupt entropy-coded data must not take the device down: every call returns, and a clean file still
    decodes bit-exactly afterwards (an out-of-bounds access in a kernel would poison the CUDA context)."""
    from wicca_b200 import decode_jpeg
    rng = np.random.default_rng(98)
    for sampling in ("420", "444"):
        good = encode(photo_like(rng, 333, 517), 85, sampling)
        scan_start = good.index(b"\xff\xda") + 14
        for it in range(60):
            d = bytearray(good)
            for _ in range(1 + int(rng.integers(0, 8))):
                d[int(rng.integers(scan_start, len(d) - 2))] = int(rng.integers(0, 256))
            if it % 5 == 0:
                d = d[:int(rng.integers(scan_start + 10, len(d)))]
            try:
                out = decode_jpeg(bytes(d))
                assert out.shape == (333, 517, 3)
            except ValueError:
                pass
        assert np.array_equal(decode_jpeg(good), reference_rgb(good))


def test_real_world_files_if_present():
    """JPEG files written by other encoders (ICC / Photoshop / EXIF segments) that ship with packages in the image."""
    import glob
    import sysconfig
    from wicca_b200 import decode_jpeg
    site = sysconfig.get_paths()["purelib"]
    paths = glob.glob(site + "/sklearn/datasets/images/*.jpg") + glob.glob(site + "/vllm/distributed/kv_transfer/*.jpg")
    if not paths:
        pytest.skip("no sample JPEG files in this environment")
    for p in paths:
        data = open(p, "rb").read()
        assert np.array_equal(decode_jpeg(data), reference_rgb(data)), p
