"""Drop-in ``WaveletCoder`` / ``HaarCoder`` backed by ``libwicca_b200.so``.

Mirrors ``wicca/wavelet_coder.py`` of the reference: same class names, same
``get_small_copy(image, transform_depth, border_type, border_constant)``
signature (``wavelet_coder.py:50-54``), same return type (a fresh C-contiguous
``uint8`` array of shape ``(ceil(H/2^d), ceil(W/2^d), C)``) and the same
exception types for bad input, so ``ClassifierProcessor``
(``classifying_tools.py:317``) and the visualisation helpers
(``visualization.py:91-94``, ``:138-141``) work unchanged when handed this coder.

All arithmetic happens in the CUDA library; this module only validates
arguments the way the reference does and moves pointers across ctypes.
"""
from __future__ import annotations

import ctypes as C
import operator
import os
import threading
from abc import ABC, abstractmethod
from typing import Iterable, Sequence

import numpy as np

from . import _capi

# cv2.BORDER_* codes (the reference takes them from cv2; values are part of OpenCV's ABI)
BORDER_CONSTANT = 0
BORDER_REPLICATE = 1
BORDER_REFLECT = 2
BORDER_WRAP = 3
BORDER_REFLECT_101 = 4

NORM_MODES = {"identity": 0, "tf": 1, "caffe": 2, "torch": 3}


class WaveletCoder(ABC):
    """Abstract interface of ``wicca/wavelet_coder.py:26-38``."""

    @abstractmethod
    def get_small_copy(self, image: np.ndarray, transform_depth: int,
                       border_type: int = BORDER_REPLICATE,
                       border_constant: int = 0) -> np.ndarray:
        """Resize the image using wavelet transform."""


def validate_image(image) -> None:
    """Same checks, order and messages as ``wicca/validation.py:80-101``.

    The reference's last check (``np.max(image) > 255``, ``:100``) is a full scan
    that cannot fail for uint8 data and is dropped.
    """
    if image is None:
        raise ValueError("Image didn't found. Please check your input.")
    if image.shape[0] == 0 or image.shape[1] == 0 or image.size == 0:   # AttributeError for non-arrays, like the reference
        raise ValueError("Image is empty")
    if image.dtype != np.uint8:
        raise ValueError("Image must be of type uint8")


def _as_depth(transform_depth) -> int:
    """``2 ** transform_depth`` (``wavelet_coder.py:58``) raises TypeError for tuples,
    strings and None; bool / numpy integers are fine.  Floats are rejected with a
    TypeError too (the reference fails on them inside cv2 or ``range``)."""
    _ = 2 ** transform_depth            # same TypeError, same message, as the reference
    try:
        return operator.index(transform_depth)
    except TypeError:
        raise TypeError(f"transform_depth must be an integer, got {type(transform_depth).__name__}") from None


def _needs_padding(h: int, w: int, depth: int) -> bool:
    if depth <= 0:
        return False
    r = 1 << depth
    return (h % r) != 0 or (w % r) != 0


def _check_layout(image: np.ndarray, depths: Sequence[int], border_type: int) -> None:
    """Argument errors of ``get_padded_copy`` (``data_loader.py:93-117``) and of the
    transform loop (``wavelet_coder.py:61-65``), raised before any device work."""
    if not isinstance(image, np.ndarray):
        raise ValueError("Image must be a numpy array")
    if image.ndim not in (2, 3):
        raise ValueError("Image must be 2D or 3D array")
    h, w = image.shape[0], image.shape[1]
    for d in depths:
        if d <= 0:
            continue
        pad = _needs_padding(h, w, d)
        # the reference indexes low_left[::2, :, :]; a 2-D array (grayscale, or a 1-channel image
        # whose channel axis cv2.copyMakeBorder dropped) fails there with IndexError
        if image.ndim == 2 or (pad and image.shape[2] == 1):
            raise IndexError("too many indices for array: array is 2-dimensional, but 3 were indexed")
        if pad:
            if (int(border_type) & ~16) not in (0, 1, 2, 3, 4):
                raise _capi.border_error_type()(f"Unknown/unsupported border type {border_type}")
            if image.shape[2] > 4:
                raise _capi.border_error_type()("copyMakeBorder supports at most 4 channels")


def _row_major_view(image: np.ndarray) -> tuple[np.ndarray, int]:
    """Return (array, row_stride_bytes) with contiguous pixels inside each row."""
    if image.ndim == 2:
        image = image[:, :, None]
    h, w, c = image.shape
    s0, s1, s2 = image.strides
    ok = (s2 == 1 or c == 1) and (s1 == c or w == 1) and s0 >= w * c
    if not ok:
        image = np.ascontiguousarray(image)
        s0 = w * c
    return image, int(s0)


class HaarCoder(WaveletCoder):
    """The Haar LL-subband "icon" coder of ``wicca/wavelet_coder.py:41-67`` on B200.

    ``HaarCoder()`` takes no arguments, like the reference.  The CUDA device is
    ``self.device`` (default: ``$WICCA_B200_DEVICE`` or 0).  Instances are
    thread-safe: the library leases a separate stream and scratch buffers to
    every concurrent call.
    """

    def __init__(self):
        super().__init__()
        self._ONE_STEP_RATIO = 2
        self.device = int(os.environ.get("WICCA_B200_DEVICE", "0"))
        self._tls = threading.local()

    # ------------------------------------------------------------------ reference API
    def get_small_copy(self, image: np.ndarray,
                       transform_depth: int,
                       border_type: int = BORDER_REPLICATE,
                       border_constant: int = 0
                       ) -> np.ndarray:
        """``HaarCoder.get_small_copy`` (``wavelet_coder.py:50-67``): icon of the image after
        ``transform_depth`` Haar levels (LL sub-band, truncated to uint8)."""
        return self.get_small_copies(image, (transform_depth,), border_type, border_constant)[0]

    # ------------------------------------------------------------------ additive extras
    def get_small_copies(self, image: np.ndarray, transform_depths: Iterable[int],
                         border_type: int = BORDER_REPLICATE, border_constant: int = 0) -> list[np.ndarray]:
        """Icons at several depths from one upload and one pass over the image (the reference
        recomputes per depth, ``classifying_tools.py:546-551``).  Same result per depth as
        :meth:`get_small_copy`."""
        validate_image(image)
        depths = [_as_depth(d) for d in transform_depths]
        if not depths:
            return []
        _check_layout(image, depths, border_type)
        lib = _capi.load()
        squeeze2d = image.ndim == 2
        view, stride = _row_major_view(image)
        h, w, c = view.shape
        outs = []
        for d in depths:
            oh, ow = (h, w) if d <= 0 else (-(-h // (1 << d)), -(-w // (1 << d)))
            outs.append(np.empty((oh, ow, c), dtype=np.uint8))
        n = len(depths)
        d_arr = (C.c_int * n)(*depths)
        p_arr = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        t = _capi.Timing()
        rc = lib.wicca_haar_icons_multi_u8(view.ctypes.data, h, w, c, stride, d_arr, n, int(border_type),
                                           float(border_constant), p_arr, int(self.device), C.byref(t))
        _capi.check(rc, "wicca_haar_icons_multi_u8")
        self._tls.timing = t.as_dict()
        if squeeze2d:
            outs = [o[:, :, 0] for o in outs]
        return outs

    def get_small_copies_batch(self, images: Sequence[np.ndarray], transform_depths: Iterable[int],
                               border_type: int = BORDER_REPLICATE, border_constant: int = 0,
                               devices: Sequence[int] | None = None,
                               out: Sequence[Sequence[np.ndarray]] | None = None) -> list[list[np.ndarray]]:
        """Icons of many images, sharded image-by-image over ``devices`` (default: every visible
        GPU) with double-buffered uploads; ``result[i][k]`` is image ``i`` at ``depths[k]`` - the
        per-image loop of ``classifying_tools.py:312-321`` without the Python overhead.
        ``out[i][k]`` (optional): preallocated C-contiguous uint8 arrays of the icon shapes to
        write into (e.g. views of a ``sharding.IconArena``); they are returned instead of new arrays."""
        depths = [_as_depth(d) for d in transform_depths]
        views = []
        for img in images:
            validate_image(img)
            _check_layout(img, depths, border_type)
            if img.ndim != 3:
                raise ValueError("batch images must be (H, W, C)")
            views.append(_row_major_view(img))
        if not views or not depths:
            return [[] for _ in views]
        cset = {v.shape[2] for v, _ in views}
        if len(cset) != 1:
            raise ValueError("all images of a batch must have the same channel count")
        c = cset.pop()
        lib = _capi.load()
        n, nd = len(views), len(depths)
        if devices is None:
            devices = list(range(max(1, lib.wicca_device_count())))
        outs: list[list[np.ndarray]] = []
        for i, (v, _) in enumerate(views):
            h, w, _c = v.shape
            shapes = [((h, w) if d <= 0 else (-(-h // (1 << d)), -(-w // (1 << d)))) + (c,) for d in depths]
            if out is None:
                outs.append([np.empty(sh, np.uint8) for sh in shapes])
                continue
            row = list(out[i])
            if len(row) != nd:
                raise ValueError(f"out[{i}] holds {len(row)} arrays for {nd} depths")
            for o, sh in zip(row, shapes):
                if not isinstance(o, np.ndarray) or o.dtype != np.uint8 or o.shape != sh or not o.flags.c_contiguous \
                        or not o.flags.writeable:
                    raise ValueError(f"out[{i}] must hold writable C-contiguous uint8 arrays of shapes {shapes}")
            outs.append(row)
        srcs = (C.c_void_p * n)(*[v.ctypes.data for v, _ in views])
        hs = (C.c_int * n)(*[v.shape[0] for v, _ in views])
        ws = (C.c_int * n)(*[v.shape[1] for v, _ in views])
        strides = (C.c_int64 * n)(*[s for _, s in views])
        d_arr = (C.c_int * nd)(*depths)
        dsts = (C.c_void_p * (n * nd))(*[o.ctypes.data for row in outs for o in row])
        dev = (C.c_int * len(devices))(*[int(x) for x in devices])
        t = _capi.Timing()
        rc = lib.wicca_batch_icons_u8(srcs, hs, ws, strides, n, c, d_arr, nd, int(border_type), float(border_constant),
                                      dsts, dev, len(devices), C.byref(t))
        _capi.check(rc, "wicca_batch_icons_u8")
        self._tls.timing = t.as_dict()
        return outs

    def forward(self, image: np.ndarray, transform_depth: int, border_type: int = BORDER_REPLICATE,
                border_constant: int = 0) -> list:
        """Full multi-level 2-D Haar analysis (extension; the reference keeps only LL).
        Returns ``[LL_d, (LH_d, HL_d, HH_d), ..., (LH_1, HL_1, HH_1)]`` as float32 views into
        one Mallat-ordered coefficient plane; ``LL_d`` truncated to uint8 is the icon."""
        validate_image(image)
        depth = _as_depth(transform_depth)
        if depth < 1:
            raise ValueError("forward transform needs transform_depth >= 1")
        _check_layout(image, (depth,), border_type)
        view, stride = _row_major_view(image)
        h, w, c = view.shape
        r = 1 << depth
        hp, wp = -(-h // r) * r, -(-w // r) * r
        plane = np.empty((hp, wp, c), dtype=np.float32)
        t = _capi.Timing()
        rc = _capi.load().wicca_haar_forward_f32(view.ctypes.data, h, w, c, stride, depth, int(border_type),
                                                 float(border_constant), plane.ctypes.data, int(self.device), C.byref(t))
        _capi.check(rc, "wicca_haar_forward_f32")
        self._tls.timing = t.as_dict()
        return mallat_to_list(plane, depth)

    def inverse(self, coeffs) -> np.ndarray:
        """Synthesis for :meth:`forward`: returns the float32 padded image ``(Hp, Wp, C)``
        (exactly the padded uint8 values for depth <= 8)."""
        plane, depth = list_to_mallat(coeffs)
        hp, wp, c = plane.shape
        out = np.empty_like(plane)
        t = _capi.Timing()
        rc = _capi.load().wicca_haar_inverse_f32(plane.ctypes.data, hp, wp, c, depth, out.ctypes.data, int(self.device),
                                                 C.byref(t))
        _capi.check(rc, "wicca_haar_inverse_f32")
        self._tls.timing = t.as_dict()
        return out

    def icons_to_batch(self, icons: Sequence[np.ndarray], shape: tuple[int, int], mode: str = "tf",
                       return_uint8: bool = False):
        """Classifier-ready batch from icons: ``cv2.resize(icon, shape, INTER_AREA)`` +
        ``np.stack`` (``classifying_tools.py:318,323``) + ``preprocess_input`` and the float32
        cast (``:286-287``).  ``shape`` is ``(width, height)`` like cv2's dsize.  Returns the
        float32 ``(B, h, w, 3)`` batch (and the uint8 batch when ``return_uint8``)."""
        if mode not in NORM_MODES:
            raise ValueError(f"unknown preprocess mode {mode!r}; expected one of {sorted(NORM_MODES)}")
        ow, oh = int(shape[0]), int(shape[1])
        if ow <= 0 or oh <= 0:
            raise ValueError("target shape must be positive")
        arrs = []
        for ic in icons:
            validate_image(ic)
            if ic.ndim != 3 or ic.shape[2] != 3:
                raise ValueError("icons must be (h, w, 3) uint8")
            arrs.append(np.ascontiguousarray(ic))
        n = len(arrs)
        out = np.empty((n, oh, ow, 3), dtype=np.float32)
        out_u8 = np.empty((n, oh, ow, 3), dtype=np.uint8) if return_uint8 else None
        if n == 0:
            return (out, out_u8) if return_uint8 else out
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        hs = (C.c_int * n)(*[a.shape[0] for a in arrs])
        ws = (C.c_int * n)(*[a.shape[1] for a in arrs])
        t = _capi.Timing()
        rc = _capi.load().wicca_icon_resize_norm_f32(ptrs, hs, ws, n, oh, ow, NORM_MODES[mode], out.ctypes.data,
                                                     out_u8.ctypes.data if return_uint8 else None, int(self.device),
                                                     C.byref(t))
        _capi.check(rc, "wicca_icon_resize_norm_f32")
        self._tls.timing = t.as_dict()
        return (out, out_u8) if return_uint8 else out

    def classifier_batches(self, images: Sequence[np.ndarray], transform_depth: int, shape: tuple[int, int],
                           mode: str = "tf", border_type: int = BORDER_REPLICATE, border_constant: int = 0,
                           with_source: bool = True, devices: Sequence[int] | None = None):
        """What ``ClassifierProcessor._get_img_batch`` + ``preprocess_input`` produce for one batch
        (``classifying_tools.py:312-323`` and ``:286-287``), in one call: ``(batch_images, batch_icons)``
        as float32 ``(B, h, w, 3)`` arrays, where ``batch_icons[i]`` comes from
        ``get_small_copy(images[i], transform_depth)`` and ``batch_images[i]`` from ``images[i]`` itself,
        both through ``cv2.resize(..., shape, INTER_AREA)`` and the ``mode`` normalisation.  The icon
        never leaves the GPU.  ``batch_images`` is ``None`` when ``with_source`` is false."""
        if mode not in NORM_MODES:
            raise ValueError(f"unknown preprocess mode {mode!r}; expected one of {sorted(NORM_MODES)}")
        depth = _as_depth(transform_depth)
        if depth < 1:
            raise ValueError("transform_depth must be >= 1")
        ow, oh = int(shape[0]), int(shape[1])
        if ow <= 0 or oh <= 0:
            raise ValueError("target shape must be positive")
        views = []
        for img in images:
            validate_image(img)
            _check_layout(img, (depth,), border_type)
            if img.ndim != 3 or img.shape[2] != 3:
                raise ValueError("classifier batches need (H, W, 3) images")
            views.append(_row_major_view(img))
        n = len(views)
        icons = np.empty((n, oh, ow, 3), dtype=np.float32)
        srcs_out = np.empty((n, oh, ow, 3), dtype=np.float32) if with_source else None
        if n == 0:
            return srcs_out, icons
        lib = _capi.load()
        if devices is None:
            devices = list(range(max(1, lib.wicca_device_count())))
        srcs = (C.c_void_p * n)(*[v.ctypes.data for v, _ in views])
        hs = (C.c_int * n)(*[v.shape[0] for v, _ in views])
        ws = (C.c_int * n)(*[v.shape[1] for v, _ in views])
        strides = (C.c_int64 * n)(*[s for _, s in views])
        dev = (C.c_int * len(devices))(*[int(x) for x in devices])
        t = _capi.Timing()
        rc = lib.wicca_batch_classifier_inputs_f32(srcs, hs, ws, strides, n, depth, int(border_type), float(border_constant),
                                                   oh, ow, NORM_MODES[mode], icons.ctypes.data,
                                                   srcs_out.ctypes.data if with_source else None, dev, len(devices),
                                                   C.byref(t))
        _capi.check(rc, "wicca_batch_classifier_inputs_f32")
        self._tls.timing = t.as_dict()
        return srcs_out, icons

    def classifier_batches_multi(self, images: Sequence[np.ndarray], transform_depths: Sequence[int],
                                 targets: Sequence[tuple], border_type: int = BORDER_REPLICATE,
                                 border_constant: int = 0, with_source: bool = True,
                                 devices: Sequence[int] | None = None):
        """Every batch the reference's ``for classifier: for depth:`` loops build from one set of images
        (``classifying_tools.py:546-551`` around ``:339-346``), from ONE upload per image.

        ``targets`` is a sequence of ``((w, h), mode)`` pairs, one per distinct classifier input
        (e.g. ``((224, 224), "tf")``, ``((224, 224), "caffe")``, ``((331, 331), "tf")``).  Returns a list
        with one ``(batch_images, {depth: batch_icons})`` entry per target, each array float32
        ``(B, h, w, 3)`` and equal to what ``classifier_batches(images, depth, (w, h), mode)`` returns;
        ``batch_images`` is ``None`` when ``with_source`` is false."""
        depths = [_as_depth(d) for d in transform_depths]
        if not depths or any(d < 1 for d in depths):
            raise ValueError("transform_depths must be a non-empty sequence of depths >= 1")
        if len(set(depths)) != len(depths):
            raise ValueError("transform_depths must not repeat")
        tlist = []
        for shape, mode in targets:
            if mode not in NORM_MODES:
                raise ValueError(f"unknown preprocess mode {mode!r}; expected one of {sorted(NORM_MODES)}")
            ow, oh = int(shape[0]), int(shape[1])
            if ow <= 0 or oh <= 0:
                raise ValueError("target shape must be positive")
            tlist.append((oh, ow, NORM_MODES[mode]))
        if not tlist:
            raise ValueError("need at least one target")
        views = []
        for img in images:
            validate_image(img)
            _check_layout(img, tuple(depths), border_type)
            if img.ndim != 3 or img.shape[2] != 3:
                raise ValueError("classifier batches need (H, W, 3) images")
            views.append(_row_major_view(img))
        n, nd, nt = len(views), len(depths), len(tlist)
        out = []
        for oh, ow, _ in tlist:
            src = np.empty((n, oh, ow, 3), dtype=np.float32) if with_source else None
            out.append((src, {d: np.empty((n, oh, ow, 3), dtype=np.float32) for d in depths}))
        if n == 0:
            return out
        lib = _capi.load()
        if devices is None:
            devices = list(range(max(1, lib.wicca_device_count())))
        srcs = (C.c_void_p * n)(*[v.ctypes.data for v, _ in views])
        hs = (C.c_int * n)(*[v.shape[0] for v, _ in views])
        ws = (C.c_int * n)(*[v.shape[1] for v, _ in views])
        strides = (C.c_int64 * n)(*[s for _, s in views])
        c_depths = (C.c_int * nd)(*depths)
        c_targets = (_capi.Target * nt)(*[_capi.Target(*t) for t in tlist])
        dst_icons = (C.c_void_p * (nt * nd))(*[out[t][1][d].ctypes.data for t in range(nt) for d in depths])
        dst_images = (C.c_void_p * nt)(*[out[t][0].ctypes.data for t in range(nt)]) if with_source else None
        dev = (C.c_int * len(devices))(*[int(x) for x in devices])
        t = _capi.Timing()
        rc = lib.wicca_batch_classifier_inputs_multi_f32(srcs, hs, ws, strides, n, c_depths, nd, int(border_type),
                                                         float(border_constant), c_targets, nt, dst_icons, dst_images, dev,
                                                         len(devices), C.byref(t))
        _capi.check(rc, "wicca_batch_classifier_inputs_multi_f32")
        self._tls.timing = t.as_dict()
        return out

    def classifier_batches_multi_from_files(self, file_paths: Sequence[str], transform_depths: Sequence[int],
                                            targets: Sequence[tuple], border_type: int = BORDER_REPLICATE,
                                            border_constant: int = 0, with_source: bool = True,
                                            devices: Sequence[int] | None = None):
        """``classifier_batches_multi`` starting where ``ClassifierProcessor._get_img_batch`` starts: from file
        paths (``classifying_tools.py:312-314``).  Baseline JPEG files are decoded on the GPU
        (``wicca_b200.data_loader``), so only the file bytes cross PCIe; other formats raise
        ``UnsupportedImageError``.  Same return value as ``classifier_batches_multi``."""
        depths = [_as_depth(d) for d in transform_depths]
        if not depths or any(d < 1 for d in depths):
            raise ValueError("transform_depths must be a non-empty sequence of depths >= 1")
        if len(set(depths)) != len(depths):
            raise ValueError("transform_depths must not repeat")
        tlist = []
        for shape, mode in targets:
            if mode not in NORM_MODES:
                raise ValueError(f"unknown preprocess mode {mode!r}; expected one of {sorted(NORM_MODES)}")
            ow, oh = int(shape[0]), int(shape[1])
            if ow <= 0 or oh <= 0:
                raise ValueError("target shape must be positive")
            tlist.append((oh, ow, NORM_MODES[mode]))
        if not tlist:
            raise ValueError("need at least one target")
        blobs = []
        for path in file_paths:
            if not path:
                raise ValueError("File path cannot be empty")
            with open(path, "rb") as fh:
                blobs.append(fh.read())
        n, nd, nt = len(blobs), len(depths), len(tlist)
        out = []
        for oh, ow, _ in tlist:
            src = np.empty((n, oh, ow, 3), dtype=np.float32) if with_source else None
            out.append((src, {d: np.empty((n, oh, ow, 3), dtype=np.float32) for d in depths}))
        if n == 0:
            return out
        lib = _capi.load()
        if devices is None:
            devices = list(range(max(1, lib.wicca_device_count())))
        datas = (C.c_void_p * n)(*[C.cast(C.c_char_p(b), C.c_void_p).value for b in blobs])
        lens = (C.c_size_t * n)(*[len(b) for b in blobs])
        c_targets = (_capi.Target * nt)(*[_capi.Target(*t) for t in tlist])
        dst_icons = (C.c_void_p * (nt * nd))(*[out[t][1][d].ctypes.data for t in range(nt) for d in depths])
        dst_images = (C.c_void_p * nt)(*[out[t][0].ctypes.data for t in range(nt)]) if with_source else None
        dev = (C.c_int * len(devices))(*[int(x) for x in devices])
        t = _capi.Timing()
        rc = lib.wicca_batch_classifier_inputs_multi_from_jpeg(datas, lens, n, (C.c_int * nd)(*depths), nd, int(border_type),
                                                               float(border_constant), c_targets, nt, dst_icons, dst_images,
                                                               dev, len(devices), C.byref(t))
        _capi.check(rc, "wicca_batch_classifier_inputs_multi_from_jpeg")
        self._tls.timing = t.as_dict()
        return out

    @property
    def last_timing(self) -> dict | None:
        """Device-side stage times (ms) of this thread's last call."""
        return getattr(self._tls, "timing", None)


# ---------------------------------------------------------------------- coefficient containers
def mallat_to_list(plane: np.ndarray, depth: int) -> list:
    """Split a Mallat-ordered plane into ``[LL_d, (LH_d, HL_d, HH_d), ..., (LH_1, HL_1, HH_1)]``
    (views, no copies).  HL = high-pass along x (right of LL), LH = high-pass along y (below)."""
    hp, wp, _ = plane.shape
    out = []
    for lvl in range(1, depth + 1):
        h, w = hp >> lvl, wp >> lvl
        out.append((plane[h:2 * h, 0:w], plane[0:h, w:2 * w], plane[h:2 * h, w:2 * w]))
    ll = plane[0:hp >> depth, 0:wp >> depth]
    return [ll] + out[::-1]


def list_to_mallat(coeffs) -> tuple[np.ndarray, int]:
    """Inverse of :func:`mallat_to_list` (copies into a fresh plane)."""
    depth = len(coeffs) - 1
    if depth < 1:
        raise ValueError("need at least one detail level")
    ll = np.asarray(coeffs[0], dtype=np.float32)
    h, w, c = ll.shape
    plane = np.empty((h << depth, w << depth, c), dtype=np.float32)
    plane[:h, :w] = ll
    for i, (lh, hl, hh) in enumerate(coeffs[1:]):
        hh_, ww_ = h << i, w << i
        if np.shape(lh) != (hh_, ww_, c) or np.shape(hl) != (hh_, ww_, c) or np.shape(hh) != (hh_, ww_, c):
            raise ValueError("detail sub-band shapes do not form a dyadic pyramid")
        plane[hh_:2 * hh_, 0:ww_] = lh
        plane[0:hh_, ww_:2 * ww_] = hl
        plane[hh_:2 * hh_, ww_:2 * ww_] = hh
    return plane, depth
