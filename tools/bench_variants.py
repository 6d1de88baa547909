#!/usr/bin/env python3
"""Developer microbenchmark: device-resident fused icon kernel, per variant / depth set.
Not the graded bench (that is bench.py); used to pick kernel parameters on the GPU box.
Needs a developer build of the library (`WICCA_DEV=1 python -m wicca_b200._build --force`): the release build
compiles the kernel variants out and ignores WICCA_ICON_VARIANT."""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np
import torch

from wicca_b200.plan import IconPlan, pitch_bytes


def main():
    n_img = int(os.environ.get("N_IMG", "30"))
    H, W = 6393, 8284
    variants = [int(v) for v in os.environ.get("VARIANTS", "0,20,22,23").split(",")]
    depth_sets = json.loads(os.environ.get("DEPTH_SETS", "[[1,2,3,4,5,6],[1],[3],[6]]"))
    dev = torch.device("cuda:0")
    pitch = pitch_bytes(W, 3)
    g = torch.Generator(device=dev); g.manual_seed(0)
    imgs = [torch.randint(0, 256, (H, pitch), dtype=torch.uint8, device=dev, generator=g) for _ in range(n_img)]
    stream = torch.cuda.current_stream().cuda_stream
    # plain copy for scale
    a = torch.empty(1 << 30, dtype=torch.uint8, device=dev); b = torch.empty_like(a)
    for _ in range(3): b.copy_(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"copy_GBps": 2 * a.numel() / best / 1e6}))
    del a, b
    results = []
    for ds in depth_sets:
        for v in variants:
            os.environ["WICCA_ICON_VARIANT"] = str(v)
            plan = IconPlan(0, [t.data_ptr() for t in imgs], [H] * n_img, [W] * n_img, [pitch] * n_img, ds)
            info = plan.info()
            for _ in range(3): plan.launch(stream)
            torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                e0.record(); plan.launch(stream); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            ts.sort()
            byt = info["bytes_read"] + info["bytes_written"]
            r = {"depths": ds, "variant": v, "ms_med": ts[len(ts) // 2], "ms_best": ts[0],
                 "GBps_med": byt / ts[len(ts) // 2] / 1e6, "GBps_best": byt / ts[0] / 1e6,
                 "MPps_med": n_img * H * W / ts[len(ts) // 2] / 1e3, "launches": info["launches"]}
            results.append(r)
            print(json.dumps(r), flush=True)
            plan.close()
    # single image launches (latency-bound regime)
    os.environ["WICCA_ICON_VARIANT"] = "0"
    plan = IconPlan(0, [imgs[0].data_ptr()], [H], [W], [pitch], [1, 2, 3, 4, 5, 6])
    for _ in range(3): plan.launch(stream)
    torch.cuda.synchronize()
    ts = []
    for i in range(20):
        e0.record(); plan.launch(stream); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); info = plan.info()
    print(json.dumps({"single_image_ms_med": ts[10], "GBps": (info["bytes_read"] + info["bytes_written"]) / ts[10] / 1e6}))


if __name__ == "__main__":
    main()
