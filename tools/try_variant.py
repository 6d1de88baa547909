#!/usr/bin/env python3
"""Developer check: run one kernel variant (WICCA_ICON_VARIANT) through the host API against the oracle.
Needs a developer build of the library (`WICCA_DEV=1 python -m wicca_b200._build --force`): the release build
compiles the kernel variants out and ignores WICCA_ICON_VARIANT."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from oracle import haar_oracle as ho
from wicca_b200 import HaarCoder

c = HaarCoder()
rng = np.random.default_rng(0)
ok = True
for (h, w) in [(64, 128), (777, 1301), (2000, 3000), (6393, 8284)]:
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    for ds in ([1, 2, 3, 4, 5, 6], [6], [1]):
        out = c.get_small_copies(img, ds)
        for d, o in zip(ds, out):
            if not np.array_equal(o, ho.haar_icon_blocksum(img, d)):
                ok = False
                print("MISMATCH", os.environ.get("WICCA_ICON_VARIANT"), (h, w), ds, d)
print("variant", os.environ.get("WICCA_ICON_VARIANT"), "OK" if ok else "FAIL")
