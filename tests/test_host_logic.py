"""Host-side mirror of the reference interface: same exception types, before any device work."""
import numpy as np
import pytest

from wicca_b200 import HaarCoder, WaveletCoder, list_to_mallat, mallat_to_list


def test_class_surface():
    c = HaarCoder()
    assert isinstance(c, WaveletCoder)
    with pytest.raises(TypeError):
        WaveletCoder()          # abstract, like the reference
    import inspect
    sig = inspect.signature(c.get_small_copy)
    assert list(sig.parameters) == ["image", "transform_depth", "border_type", "border_constant"]
    assert sig.parameters["border_type"].default == 1 and sig.parameters["border_constant"].default == 0


def test_validation_errors_match_reference_types():
    c = HaarCoder()
    with pytest.raises(ValueError, match="didn't found"):
        c.get_small_copy(None, 1)
    with pytest.raises(AttributeError):
        c.get_small_copy([[1, 2], [3, 4]], 1)
    with pytest.raises(ValueError, match="empty"):
        c.get_small_copy(np.zeros((0, 4, 3), np.uint8), 1)
    with pytest.raises(ValueError, match="uint8"):
        c.get_small_copy(np.zeros((4, 4, 3), np.float32), 1)
    with pytest.raises(ValueError, match="2D or 3D"):
        c.get_small_copy(np.zeros((4, 4, 3, 1), np.uint8), 1)
    for bad in ((2,), "2", None):
        with pytest.raises(TypeError):
            c.get_small_copy(np.zeros((4, 4, 3), np.uint8), bad)
    with pytest.raises(TypeError):
        c.get_small_copy(np.zeros((4, 4, 3), np.uint8), 2.0)
    with pytest.raises(IndexError):
        c.get_small_copy(np.zeros((4, 4), np.uint8), 1)              # grayscale: reference indexes 3 axes
    with pytest.raises(IndexError):
        c.get_small_copy(np.zeros((5, 7, 1), np.uint8), 1)           # cv2 drops the channel axis when padding
    from wicca_b200 import _capi
    with pytest.raises(_capi.border_error_type()):
        c.get_small_copy(np.zeros((5, 7, 3), np.uint8), 1, 5)        # BORDER_TRANSPARENT -> cv2.error
    with pytest.raises(_capi.border_error_type()):
        c.get_small_copy(np.zeros((5, 7, 5), np.uint8), 1)           # > 4 channels cannot be padded


def test_identity_depths_need_no_gpu():
    c = HaarCoder()
    img = np.random.default_rng(0).integers(0, 256, (9, 11, 3), dtype=np.uint8)
    for d in (0, -1, False):
        out = c.get_small_copy(img, d)
        assert out is not img and out.flags.c_contiguous and out.flags.writeable and np.array_equal(out, img)
    out = c.get_small_copy(image=img[:, ::2], transform_depth=0)     # keyword call, non-contiguous input
    assert np.array_equal(out, img[:, ::2])
    g = c.get_small_copy(img[:, :, 0], 0)                            # 2-D + depth 0: the reference returns a copy
    assert g.shape == (9, 11) and np.array_equal(g, img[:, :, 0])


def test_mallat_container_roundtrip():
    rng = np.random.default_rng(1)
    plane = rng.random((16, 24, 3), dtype=np.float32)
    lst = mallat_to_list(plane, 2)
    assert lst[0].shape == (4, 6, 3) and lst[1][0].shape == (4, 6, 3) and lst[2][2].shape == (8, 12, 3)
    back, depth = list_to_mallat(lst)
    assert depth == 2 and np.array_equal(back, plane)
