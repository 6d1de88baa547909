#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel, meant to run under compute-sanitizer."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from oracle import haar_oracle as ho, resize_oracle as ro
from wicca_b200 import HaarCoder

c = HaarCoder()
rng = np.random.default_rng(0)
for (h, w) in [(64, 128), (65, 129), (200, 259), (333, 517), (1, 1), (700, 1100)]:
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    for bt in (1, 0, 2, 3, 4):
        out = c.get_small_copies(img, [1, 2, 3, 4, 5, 6], bt, 7)
        for d, o in zip(range(1, 7), out):
            assert np.array_equal(o, ho.haar_icon_blocksum(img, d, bt, 7)), (h, w, bt, d)
    assert np.array_equal(c.get_small_copy(img, 8), ho.haar_icon_blocksum(img, 8))
    assert np.array_equal(c.get_small_copy(img, 10), ho.haar_icon_fp32(img, 10))
img4 = rng.integers(0, 256, (97, 161, 4), dtype=np.uint8)
assert np.array_equal(c.get_small_copy(img4, 2, 2), ho.haar_icon_fp32(img4, 2, 2))
img = rng.integers(0, 256, (130, 200, 3), dtype=np.uint8)
co = c.forward(img, 3)
assert np.array_equal(c.inverse(co), ho.get_padded_copy(img, 8).astype(np.float32))
icons = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in [(100, 130), (300, 259), (448, 448)]]
f32, u8 = c.icons_to_batch(icons, (224, 224), "torch", return_uint8=True)
for i, ic in enumerate(icons):
    assert np.array_equal(u8[i], ro.resize_area(ic, 224, 224))
batch = c.get_small_copies_batch([rng.integers(0, 256, (300 + i, 400 + i, 3), dtype=np.uint8) for i in range(5)], [2, 6])
assert len(batch) == 5
print("sanitize smoke ok")
