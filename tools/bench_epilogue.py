"""Epilogue kernel (rows A5+A6, N1) alone: `resize_area_rows_kernel` on 30 resident icons of depth 1-3 -> 224 / 299 / 331 px
and on the 30 source images -> 224 px, CUDA events around 20 launches each, one JSON line per row.  Before the timing
every configuration is compared with `cv2.resize(..., INTER_AREA)` on three of the images (uint8 result, bit for bit).
    python tools/bench_epilogue.py > gpurun_out/epilogue.jsonl"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wicca_b200 import _capi
from wicca_b200.plan import IconPlan, pitch_bytes

PEAK = 6544.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
try:
    import cv2
except Exception:
    cv2 = None

H, W, n = 6393, 8284, 30
dev = torch.device("cuda:0")
lib = _capi.load()
pitch = pitch_bytes(W, 3)
g = torch.Generator(device=dev); g.manual_seed(0)
imgs = [torch.randint(0, 256, (H, pitch), dtype=torch.uint8, device=dev, generator=g) for _ in range(n)]
stream = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def check(name, src_of, out_u8, target):
    """uint8 result of images 0, 7, 29 against cv2 on the same source."""
    if cv2 is None:
        return None
    bad = 0
    for i in (0, 7, 29):
        want = cv2.resize(src_of(i), (target, target), interpolation=cv2.INTER_AREA)
        bad += int((out_u8[i].cpu().numpy() != want).sum())
    if bad:
        print(json.dumps({"row": name, "MISMATCH_vs_cv2": bad}), flush=True)
        sys.exit(1)
    return 0


depths = [1, 2, 3]
plan = IconPlan(0, [t.data_ptr() for t in imgs], [H] * n, [W] * n, [pitch] * n, depths)
plan.launch(stream)
torch.cuda.synchronize()
for target in (224, 299, 331):
    out = torch.empty((n, target, target, 3), dtype=torch.float32, device=dev)
    out8 = torch.empty((n, target, target, 3), dtype=torch.uint8, device=dev)
    for k, d in enumerate(depths):
        _, ih, iw, _ = plan.icon_info(0, k)
        plan.resize_norm(k, target, target, 1, out.data_ptr(), out8.data_ptr(), stream)
        torch.cuda.synchronize()
        ok = check(f"epilogue d{d} -> {target}", lambda i: plan.read_icon(i, k), out8, target)
        tf = out8[:3].cpu().numpy().astype(np.float32) / np.float32(127.5) - np.float32(1.0)     # NumPy: true division
        assert np.array_equal(tf, out[:3].cpu().numpy()), "tf normalisation differs from true division"
        ms = timed(lambda: plan.resize_norm(k, target, target, 1, out.data_ptr(), 0, stream))
        byt = n * (ih * iw * 3 + target * target * 3 * 4)
        print(json.dumps({"row": "A5+A6 epilogue", "config": f"30 icons of depth {d} ({ih}x{iw}) -> {target}x{target} tf", "ms": round(ms, 4),
                          "GBps": round(byt / ms / 1e6, 1), "frac_of_measured_peak": round(byt / ms / 1e6 / PEAK, 3),
                          "mismatches_vs_cv2": ok}), flush=True)
plan.close()

# row N1: the 53 MP source images -> 224 px
out_s = torch.empty((n, 224, 224, 3), dtype=torch.float32, device=dev)
out_s8 = torch.empty((n, 224, 224, 3), dtype=torch.uint8, device=dev)
ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in imgs])
Hs, Ws, Ps = (C.c_int * n)(*[H] * n), (C.c_int * n)(*[W] * n), (C.c_int64 * n)(*[pitch] * n)


def src_resize(u8=None):
    _capi.check(lib.wicca_resize_norm_dev(ptrs, Hs, Ws, Ps, n, 224, 224, 1, out_s.data_ptr(), u8, 0, C.c_void_p(stream)), "resize_norm_dev")


src_resize(C.c_void_p(out_s8.data_ptr()))
torch.cuda.synchronize()
ok = check("N1", lambda i: np.ascontiguousarray(imgs[i].cpu().numpy()[:, :W * 3].reshape(H, W, 3)), out_s8, 224)
ms = timed(src_resize, warm=2)
print(json.dumps({"row": "N1 source-image resize", "config": f"30 x ({H},{W},3) -> 224x224 tf", "ms": round(ms, 4),
                  "GBps": round(n * H * W * 3 / ms / 1e6, 1), "frac_of_measured_peak": round(n * H * W * 3 / ms / 1e6 / PEAK, 3),
                  "mismatches_vs_cv2": ok}), flush=True)
