"""Forward / inverse sub-band timings for every channel count (A/B helper: WICCA_FORWARD_TILE=1 selects the old kernel)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wicca_b200 import _capi
from wicca_b200.plan import pitch_bytes

lib = _capi.load()
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
S = int(os.environ.get("S", 8192))
for ch in (1, 2, 3, 4):
    pitch = pitch_bytes(S, ch)
    img = torch.randint(0, 256, (S, pitch), dtype=torch.uint8, device=dev)
    coeffs = torch.empty((S, S, ch), dtype=torch.float32, device=dev)
    work = torch.empty((S * S * ch * 5 // 16 + 64,), dtype=torch.float32, device=dev)
    rec = torch.empty((S, S, ch), dtype=torch.float32, device=dev)
    for depth in (1, 3, 6):
        def fwd():
            _capi.check(lib.wicca_haar_forward_dev(img.data_ptr(), S, S, ch, pitch, depth, 1, 0.0, coeffs.data_ptr(),
                                                   work.data_ptr(), 0, C.c_void_p(stream)), "forward_dev")

        def inv():
            _capi.check(lib.wicca_haar_inverse_dev(coeffs.data_ptr(), S, S, ch, depth, rec.data_ptr(), work.data_ptr(), 0,
                                                   C.c_void_p(stream)), "inverse_dev")
        out = {}
        for name, fn in (("fwd", fwd), ("inv", inv)):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(10):
                fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            byts = (5 if name == "fwd" else 8) * ch * S * S
            out[name] = {"ms": round(ms, 4), "GBps": round(byts / ms / 1e6, 1)}
        print(json.dumps({"C": ch, "depth": depth, **out}))
