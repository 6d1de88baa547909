#!/usr/bin/env python3
"""Launch each (variant, depth set) twice; run under `ncu --metrics gpu__time_duration.sum` to get
kernel-only durations without host launch gaps.
Needs a developer build of the library (`WICCA_DEV=1 python -m wicca_b200._build --force`): the release build
compiles the kernel variants out and ignores WICCA_ICON_VARIANT."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from wicca_b200.plan import IconPlan, pitch_bytes

n = int(os.environ.get("N_IMG", "30"))
H, W = 6393, 8284
pitch = pitch_bytes(W, 3)
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
imgs = [torch.randint(0, 256, (H, pitch), dtype=torch.uint8, device="cuda:0", generator=g) for _ in range(n)]
st = torch.cuda.current_stream().cuda_stream
variants = [int(v) for v in os.environ.get("VARIANTS", "0").split(",")]
depth_sets = json.loads(os.environ.get("DEPTH_SETS", "[[1,2,3,4,5,6]]"))
for ds in depth_sets:
    for v in variants:
        os.environ["WICCA_ICON_VARIANT"] = str(v)
        plan = IconPlan(0, [t.data_ptr() for t in imgs], [H] * n, [W] * n, [pitch] * n, ds)
        for _ in range(2):
            plan.launch(st)
        torch.cuda.synchronize()
        print("LAUNCHED", json.dumps({"depths": ds, "variant": v}), flush=True)
        plan.close()
