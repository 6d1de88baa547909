"""PyTorch tensor bridge: hand device buffers to the C ABI without a host round trip.

Only pointers cross the boundary (``tensor.data_ptr()`` and the current CUDA stream); PyTorch is
the allocator here, not the compute path.  Use this when the image is already in HBM (decoded on
the GPU, or reused across depths / classifiers).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

from . import _capi


def icons_from_cuda_tensor(image, transform_depths: Sequence[int], border_type: int = 1, border_constant: float = 0.0):
    """Icons of a ``(H, W, C)`` uint8 CUDA tensor at ``transform_depths`` (each >= 1), returned as CUDA
    uint8 tensors ``(ceil(H/2^d), ceil(W/2^d), C)`` on the same device; enqueued on the current stream.

    The one-pass kernel needs 16-byte aligned rows: a tensor whose row length ``W*C`` is not a multiple
    of 16 is first copied into a pitched staging tensor (one extra device-to-device pass).  Icon
    tensors are views of pitched storage (``icon.is_contiguous()`` may be False); call ``.contiguous()``
    if a dense tensor is required.
    """
    import torch  # noqa: PLC0415

    if not (image.is_cuda and image.dtype == torch.uint8 and image.dim() == 3):
        raise ValueError("image must be a (H, W, C) uint8 CUDA tensor")
    depths = [int(d) for d in transform_depths]
    if not depths or min(depths) < 1:
        raise ValueError("depths must be >= 1")
    h, w, c = image.shape
    dev = image.device.index or 0
    lib = _capi.load()
    row = w * c
    if image.stride(2) == 1 and image.stride(1) == c and image.stride(0) % 16 == 0 and image.data_ptr() % 16 == 0:
        src, pitch = image, image.stride(0)
    else:
        pitch = int(lib.wicca_pitch_bytes(w, c))
        src = torch.empty((h, pitch), dtype=torch.uint8, device=image.device)
        src[:, :row].copy_(image.reshape(h, row))
    outs, ptrs, pitches = [], [], []
    for d in depths:
        oh, ow = -(-h // (1 << d)), -(-w // (1 << d))
        op = ow * c if d > 8 else (ow * c + 127) // 128 * 128
        buf = torch.empty((oh, op), dtype=torch.uint8, device=image.device)
        outs.append(buf[:, : ow * c].unflatten(1, (ow, c)))
        ptrs.append(buf.data_ptr())
        pitches.append(op)
    n = len(depths)
    stream = torch.cuda.current_stream(image.device).cuda_stream
    rc = lib.wicca_haar_icons_multi_dev(src.data_ptr(), h, w, c, pitch, (C.c_int * n)(*depths), n, int(border_type),
                                        float(border_constant), (C.c_void_p * n)(*ptrs), (C.c_int64 * n)(*pitches), dev,
                                        C.c_void_p(stream))
    _capi.check(rc, "wicca_haar_icons_multi_dev")
    # `src` (if it is a staging copy) must stay alive until the kernel has run: tie it to the stream
    if src is not image:
        src.record_stream(torch.cuda.current_stream(image.device))
    return outs
