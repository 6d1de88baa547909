#!/usr/bin/env python3
"""Headline benchmark: MP/s of the multi-level Haar DWT + LL icon path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (NumPy port)

Workload (BASELINE.json configs[1]): a batch of 30 synthetic (6393, 8284, 3) uint8 images
(8284 x 6393 px, 52.96 MP each); one STEP = every image of the batch -> its six icons at depths
1..6.  MP are input megapixels, each image counted ONCE per step (the six icons come out of one
pass), so the reference arm - which has to call get_small_copy six times per image - is measured
in the same unit on the same work.

`value`  : device-resident (images already in HBM, pitched layout), K steps back to back, CUDA
           events on the launch stream, max over ranks.
`e2e`    : the same step through the C ABI with HOST buffers (wicca_batch_icons_u8: pinned host
           images -> H2D -> kernel -> D2H of the icons), copies inside the timed region.
`roofline`, `cpu_baseline`, `clocks`, `gpu_launches`: see DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

H, W, CH = 6393, 8284, 3
BATCH = 30
DEPTHS = [1, 2, 3, 4, 5, 6]
MP_PER_IMAGE = H * W / 1e6
METRIC = "MP/s multi-level Haar DWT+icon (depths 1-6 per image, batch of 30 x 8284x6393x3)"
UNIT = "MP/s"


def workload_config(n_gpus: int) -> dict:
    return {"workload": "configs[1]: batch of 30 synthetic 8284x6393x3 uint8 images, icons at depths 1-6 per image "
                        "(one fused pass), BORDER_REPLICATE",
            "images_per_gpu": BATCH, "image_hwc": [H, W, CH], "depths": DEPTHS,
            "global_batch": BATCH * n_gpus, "parallelism": f"images sharded, {n_gpus} GPU(s), no collective",
            "l2_policy": "inputs (4.77 GB per GPU) are 38x larger than L2; no flush needed"}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the NumPy port of the reference (oracle/haar_oracle.py)
# ----------------------------------------------------------------------------------------------
def _cpu_sample(job) -> float:
    """Six get_small_copy-equivalents (depths 1..6) on one synthetic (rows, 8284, 3) image; seconds."""
    seed, rows = job
    from oracle import haar_oracle as ho
    img = ho.synthetic_image(seed, rows, W, CH)
    t0 = time.perf_counter()
    for d in DEPTHS:
        ho.haar_icon_fp32(img, d)
    return time.perf_counter() - t0


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline_single(budget_images: int = 2) -> dict:
    """1 core, bounded sample: `budget_images` full-size images x 6 depths (about 6 s per image)."""
    secs = [_cpu_sample((i, H)) for i in range(budget_images)]
    best = min(secs)
    return {"value": MP_PER_IMAGE / best, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{budget_images} image(s) of {H}x{W}x3 x depths 1-6 through oracle.haar_icon_fp32 (NumPy "
                      f"restatement of wavelet_coder.py:50-67), best image: {best:.2f} s"}


def run_reference(args) -> int:
    """bench.py --impl reference: the reference's CPU algorithm (NumPy port) on all host cores, one
    image per worker process per step.  Each step is a bounded sample: full-width images whose
    height is sized so that the whole --steps/--warmup run stays within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = host_cores()
    workers = max(1, min(cores, 64))            # each worker holds up to ~1 GB of NumPy temporaries
    budget_s = float(os.environ.get("WICCA_REF_BUDGET_S", "150"))
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        def step(base, rows):
            t0 = time.perf_counter()
            pool.map(_cpu_sample, [(base + i, rows) for i in range(workers)])
            return time.perf_counter() - t0
        step(0, 64)                                                    # spin the pool up
        cal_rows = 512
        cal = step(100, cal_rows)                                      # aggregate rate with every core busy
        per_step = budget_s / max(1, args.steps + args.warmup)
        rows = int(cal_rows * per_step / max(cal, 1e-6))
        rows = max(64, min(H, rows // 64 * 64))
        if rows >= H - 64:
            rows = H
        for w in range(args.warmup):
            step(1000 * (w + 1), rows)
        times = [step(100_000 + 1000 * k, rows) for k in range(args.steps)]
    total = sum(times)
    mp_per_step = workers * rows * W / 1e6
    value = args.steps * mp_per_step / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                             "sample": f"each step: {workers} images of {rows}x{W}x3 (one per worker process) x depths "
                                       "1-6 through oracle.haar_icon_fp32, the NumPy restatement of the reference "
                                       "(the reference itself is Python; /root/reference does not exist on the GPU box)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0: float, t1: float) -> dict:
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if not (t0 <= ts <= t1 + 0.05):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args) -> int:
    import torch
    import torch.distributed as dist

    from wicca_b200 import _capi
    from wicca_b200.plan import IconPlan, pitch_bytes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    distributed = world > 1
    if distributed:
        if not os.environ.get("WICCA_KEEP_NCCL_DEBUG"):
            os.environ.pop("NCCL_DEBUG", None)      # NCCL prints its banner on stdout; keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    lib = _capi.load()
    pitch = pitch_bytes(W, CH)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    # resident inputs: 30 pitched images per GPU, uniform uint8 noise
    imgs = [torch.randint(0, 256, (H, pitch), dtype=torch.uint8, device=dev, generator=gen) for _ in range(BATCH)]
    plan = IconPlan(local, [t.data_ptr() for t in imgs], [H] * BATCH, [W] * BATCH, [pitch] * BATCH, DEPTHS)
    info = plan.info()
    alg_bytes = info["bytes_read"] + info["bytes_written"]
    stream = torch.cuda.current_stream().cuda_stream

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- device-resident timing -----------------------------------------------------------
    for _ in range(args.warmup):
        plan.launch(stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_clk0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        plan.launch(stream)
    e1.record()
    barrier()
    t_clk1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    ms_t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    ms_per_step = ms_max / args.steps
    value = world * BATCH * MP_PER_IMAGE / (ms_per_step / 1e3)

    # ---- end to end through the C ABI with host buffers -------------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    n_host = min(BATCH, args.e2e_host_images)           # distinct pinned host images, cycled
    host_ptrs, host_arrays = [], []
    rng = np.random.default_rng(99 + rank)
    for i in range(n_host):
        p = C.c_void_p()
        _capi.check(lib.wicca_host_alloc_near(C.byref(p), H * W * CH, local), "wicca_host_alloc_near")
        arr = np.ctypeslib.as_array((C.c_uint8 * (H * W * CH)).from_address(p.value)).reshape(H, W, CH)
        arr[:] = rng.integers(0, 256, (H, W, CH), dtype=np.uint8)
        host_ptrs.append(p)
        host_arrays.append(arr)
    icon_shapes = [(-(-H // (1 << d)), -(-W // (1 << d)), CH) for d in DEPTHS]
    out_ptrs = []

    def pinned_array(shape):
        n = int(np.prod(shape))
        q = C.c_void_p()
        _capi.check(lib.wicca_host_alloc_near(C.byref(q), n, local), "wicca_host_alloc_near")
        out_ptrs.append(q)
        return np.ctypeslib.as_array((C.c_uint8 * n).from_address(q.value)).reshape(shape)

    outs = [[pinned_array(s) for s in icon_shapes] for _ in range(BATCH)]
    nd = len(DEPTHS)
    srcs = (C.c_void_p * BATCH)(*[host_ptrs[i % n_host].value for i in range(BATCH)])
    hs = (C.c_int * BATCH)(*[H] * BATCH)
    ws = (C.c_int * BATCH)(*[W] * BATCH)
    strides = (C.c_int64 * BATCH)(*[0] * BATCH)
    d_arr = (C.c_int * nd)(*DEPTHS)
    dsts = (C.c_void_p * (BATCH * nd))(*[o.ctypes.data for row in outs for o in row])
    devs = (C.c_int * 1)(local)
    tim = _capi.Timing()

    def e2e_step():
        _capi.check(lib.wicca_batch_icons_u8(srcs, hs, ws, strides, BATCH, CH, d_arr, nd, 1, 0.0, dsts, devs, 1,
                                             C.byref(tim)), "wicca_batch_icons_u8")

    e2e_step()                                             # warm-up (allocations, page faults)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())
    e2e_value = world * BATCH * e2e_steps * MP_PER_IMAGE / e2e_s
    h2d_bytes = BATCH * H * W * CH
    d2h_bytes = BATCH * sum(int(np.prod(s)) for s in icon_shapes)
    stage_ms = tim.as_dict()                               # sums over the 30 images of the last step

    # parity spot check of the e2e outputs against the device-resident plan (same kernel, other data path)
    if rank == 0:
        from oracle import haar_oracle as ho  # noqa: PLC0415  (checker only, outside every timed region)
        chk = ho.haar_icon_blocksum(host_arrays[0], 6)
        assert np.array_equal(outs[0][5], chk), "e2e icon mismatch vs oracle"

    sampler.stop()
    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
        achieved = alg_bytes / (ms_per_step / 1e3) / 1e9
        traffic = None
        try:
            traffic = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
        cpu = cpu_baseline_single(args.cpu_images) if world == 1 and args.cpu_images > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 in / u32 block sums / u8 out", "data": "synthetic", "config": workload_config(world),
            "gpu_launches": args.steps * info["launches"],
            "kernel": "wicca::haar_icon_tma2_kernel<6> (one launch per step: 30 images x 6 depths)",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "achieved = (30 x H*W*3 read + sum of the six icons written) / mean launch time, "
                                 "K launches back to back between two CUDA events on the launch stream"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "api": "wicca_batch_icons_u8 (pinned host images and icons, two upload slots per GPU so H2D(i+1) overlaps kernel/D2H(i))",
                    "stage_ms_sum_over_images": stage_ms,
                    "h2d_GBps": h2d_bytes / max(stage_ms["h2d_ms"], 1e-9) / 1e6},
            "clocks": sampler.summary(t_clk0, t_clk1 + e2e_s + 5.0),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
    for p in host_ptrs + out_ptrs:
        lib.wicca_host_free(p)
    plan.close()
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3, help="end-to-end (host buffer) steps, each ~4.8 GB of H2D")
    ap.add_argument("--e2e-host-images", type=int, default=10, help="distinct pinned host images cycled through the batch")
    ap.add_argument("--cpu-images", type=int, default=2, help="images of the bounded cpu_baseline sample (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    raise SystemExit(main())
