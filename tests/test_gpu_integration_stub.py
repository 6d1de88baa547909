"""The drop-in boundary, exercised the way INTEGRATION.md describes it, against the UNMODIFIED reference staged under
oracle/_ref (the reference's own `wicca` package, importable on the GPU box; skipped where it was not staged):

  * section 2's ctypes stub, written as a subclass of the REFERENCE's `WaveletCoder` ABC, binding `wicca_haar_icon_u8`;
  * `wicca_b200.HaarCoder` handed to a loop shaped like `ClassifierProcessor._get_img_batch`
    (classifying_tools.py:312-323: resize the image, get_small_copy, resize the icon, np.stack) next to the
    reference's own `HaarCoder` - positional call as at :317, keyword call as at visualization.py:91-94.
"""
import ctypes

import numpy as np
import pytest

from oracle import ref_loader
from tests.golden.make_golden import gen_input

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not ref_loader.available():
        pytest.skip("oracle/_ref is not staged on this machine")
    ref_loader.load_haar_coder()
    import wicca.wavelet_coder as ref_wc          # the reference's own module, from oracle/_ref
    return ref_wc


def test_integration_md_stub_is_a_reference_wavelet_coder(ref):
    import cv2

    from wicca.validation import validate_image   # reference code, unchanged
    from wicca_b200 import _capi
    _lib = ctypes.CDLL(str(_capi.library_path()))
    _lib.wicca_haar_icon_u8.restype = ctypes.c_int
    _lib.wicca_haar_icon_u8.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_void_p]
    _lib.wicca_last_error.restype = ctypes.c_char_p

    class HaarCoderB200(ref.WaveletCoder):        # INTEGRATION.md section 2, verbatim
        def get_small_copy(self, image, transform_depth, border_type=cv2.BORDER_REPLICATE, border_constant=0):
            validate_image(image)
            image = np.ascontiguousarray(image)
            h, w, c = image.shape
            d = int(transform_depth)
            oh, ow = (h, w) if d <= 0 else (-(-h >> d), -(-w >> d))
            out = np.empty((oh, ow, c), np.uint8)
            rc = _lib.wicca_haar_icon_u8(image.ctypes.data, h, w, c, w * c, d, border_type, float(border_constant),
                                         out.ctypes.data, 0, None)
            if rc > 0:
                raise RuntimeError(_lib.wicca_last_error().decode())
            if rc < 0:
                raise (cv2.error if rc == -3 else ValueError)(_lib.wicca_last_error().decode())
            return out

    stub, theirs = HaarCoderB200(), ref.HaarCoder()
    assert isinstance(stub, ref.WaveletCoder)
    for (h, w, d, bt) in [(517, 771, 3, cv2.BORDER_REPLICATE), (2048, 1024, 6, cv2.BORDER_REPLICATE), (300, 201, 2, cv2.BORDER_REFLECT_101),
                          (1999, 3001, 5, cv2.BORDER_CONSTANT), (64, 64, 8, cv2.BORDER_WRAP), (33, 77, 0, cv2.BORDER_REPLICATE)]:
        img = gen_input("noise", h + w + d, h, w, 3)
        a = stub.get_small_copy(img, d, bt, 17)
        b = theirs.get_small_copy(img, d, bt, 17)
        assert a.dtype == b.dtype == np.uint8 and a.shape == b.shape and np.array_equal(a, b), (h, w, d, bt)
    with pytest.raises(ValueError):
        stub.get_small_copy(np.zeros((4, 4, 3), np.float32), 1)          # the reference's own validate_image
    with pytest.raises(ValueError):
        stub.get_small_copy(None, 1)


def test_drop_in_class_inside_the_reference_batch_loop(ref):
    """The body of ClassifierProcessor._get_img_batch with our coder and with the reference's, same images."""
    import cv2

    from wicca_b200 import HaarCoder
    ours, theirs = HaarCoder(), ref.HaarCoder()
    images = [gen_input("noise", 40 + i, 1500 + 97 * i, 2100 - 55 * i, 3) for i in range(4)]
    for depth in (2, 3, 6):
        for shape in ((224, 224), (331, 331)):
            def batch(coder):
                resized_imgs, resized_icons = [], []
                for image in images:                                    # classifying_tools.py:312-321
                    resized_imgs.append(cv2.resize(image, shape, interpolation=cv2.INTER_AREA))
                    icon = coder.get_small_copy(image, depth)           # :317, positional
                    resized_icons.append(cv2.resize(icon, shape, interpolation=cv2.INTER_AREA))
                return np.stack(resized_imgs), np.stack(resized_icons)  # :323
            bi_a, bc_a = batch(ours)
            bi_b, bc_b = batch(theirs)
            assert np.array_equal(bi_a, bi_b) and np.array_equal(bc_a, bc_b), (depth, shape)
            # and the fused call produces the float32 batches preprocess_input("tf") would make of them
            src_f32, icon_f32 = ours.classifier_batches(images, depth, shape, "tf")
            from oracle import resize_oracle as ro
            assert np.array_equal(icon_f32, ro.preprocess_input(bc_b, "tf")) and np.array_equal(src_f32, ro.preprocess_input(bi_b, "tf"))
    # keyword call, as the reference's visualisation helpers make it (visualization.py:91-94)
    a = ours.get_small_copy(image=images[0], transform_depth=4)
    assert np.array_equal(a, theirs.get_small_copy(image=images[0], transform_depth=4))
    assert a.flags.c_contiguous and a.flags.writeable and a.flags.owndata
