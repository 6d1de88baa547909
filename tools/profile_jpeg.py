"""Decode one synthetic 53 MP baseline JPEG (q90, 4:2:0) to icons a few times (for an ncu launch list of the JPEG
ingest path: Huffman passes, scans, IDCT, colour, icon kernel)."""
import os
import sys
import time

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wicca_b200 import icons_from_jpeg

H, W = 6393, 8284
rng = np.random.default_rng(3)
yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
img = np.stack([128 + 90 * np.sin(xx / (37.0 + 9 * c) + c) + 70 * np.cos(yy / (23.0 + 5 * c) - c) for c in range(3)], -1)
img = np.clip(img + rng.normal(0, 6, img.shape).astype(np.float32), 0, 255).astype(np.uint8)
ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90])
data = bytes(enc)
for _ in range(3):
    tm = {}
    t0 = time.perf_counter()
    icons_from_jpeg(data, [1, 2, 3, 4, 5, 6], timing=tm)
    print("wall ms", round((time.perf_counter() - t0) * 1e3, 2), {k: round(v, 3) for k, v in tm.items()})
