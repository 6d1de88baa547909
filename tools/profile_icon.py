#!/usr/bin/env python3
"""Short program for ncu: a few launches of the fused icon kernel on device-resident images."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from wicca_b200.plan import IconPlan, pitch_bytes

n = int(os.environ.get("N_IMG", "4"))
H, W = 6393, 8284
pitch = pitch_bytes(W, 3)
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
imgs = [torch.randint(0, 256, (H, pitch), dtype=torch.uint8, device="cuda:0", generator=g) for _ in range(n)]
st = torch.cuda.current_stream().cuda_stream
for ds in ([1, 2, 3, 4, 5, 6], [6], [1]):
    plan = IconPlan(0, [t.data_ptr() for t in imgs], [H] * n, [W] * n, [pitch] * n, ds)
    for _ in range(2):
        plan.launch(st)
    torch.cuda.synchronize()
    plan.close()
print("done")
