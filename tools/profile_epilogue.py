"""Epilogue (rows A5+A6) on 30 resident depth-1 icons for 224 and 331 px targets (for an ncu capture of
`resize_area_rows_kernel`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wicca_b200.plan import IconPlan, pitch_bytes

H, W, n = 6393, 8284, 30
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
pitch = pitch_bytes(W, 3)
imgs = [torch.randint(0, 256, (H, pitch), dtype=torch.uint8, device=dev) for _ in range(n)]
stream = torch.cuda.current_stream().cuda_stream
plan = IconPlan(0, [t.data_ptr() for t in imgs], [H] * n, [W] * n, [pitch] * n, [depth])
plan.launch(stream)
for target in (224, 331):
    out = torch.empty((n, target, target, 3), dtype=torch.float32, device=dev)
    for _ in range(3):
        plan.resize_norm(0, target, target, 1, out.data_ptr(), 0, stream)
torch.cuda.synchronize()
plan.close()
print("ok")
