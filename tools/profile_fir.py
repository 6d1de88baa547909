"""One db4 icon of a 53 MP image (for an ncu capture of the tiled FIR kernel)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wicca_b200 import OrthogonalWaveletCoder

img = np.random.default_rng(5).integers(0, 256, (6393, 8284, 3), dtype=np.uint8)
coder = OrthogonalWaveletCoder(sys.argv[1] if len(sys.argv) > 1 else "db4")
for _ in range(3):
    coder.get_small_copy(img, 3)
print(coder.last_timing)
