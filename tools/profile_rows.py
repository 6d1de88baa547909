"""One (6393, 8284, 4) image through haar_icon_rows_kernel at depths 3, 6 and 8 (for an ncu capture)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wicca_b200 import _capi
from wicca_b200.plan import pitch_bytes

H, W, ch = 6393, 8284, 4
lib = _capi.load()
pitch = pitch_bytes(W, ch)
img = torch.randint(0, 256, (H, pitch), dtype=torch.uint8, device="cuda:0")
stream = torch.cuda.current_stream().cuda_stream
for d in (3, 6, 8):
    oh, ow = -(-H // (1 << d)), -(-W // (1 << d))
    op = -(-ow * ch // 128) * 128
    out = torch.zeros((oh, op), dtype=torch.uint8, device="cuda:0")
    for _ in range(2):
        _capi.check(lib.wicca_haar_icons_multi_dev(img.data_ptr(), H, W, ch, pitch, (C.c_int * 1)(d), 1, 1, 0.0,
                                                   (C.c_void_p * 1)(out.data_ptr()), (C.c_int64 * 1)(op), 0, C.c_void_p(stream)), "dev")
torch.cuda.synchronize()
print("ok")
