#!/usr/bin/env python3
"""Headline benchmark: MP/s of the multi-level Haar DWT + LL icon path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path (oracle/_ref)

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
# reference arm / cpu baseline: the UNMODIFIED reference HaarCoder staged under oracle/_ref
# (wicca/wavelet_coder.py:50-67, padding by cv2.copyMakeBorder, wicca/data_loader.py:116-117);
# the NumPy port of oracle/haar_oracle.py only when that import fails.  Always full 6393 x 8284 images.
# ----------------------------------------------------------------------------------------------
_CPU_IMAGES: list = []          # filled before the worker pool forks (copy-on-write, read only)
_CPU_CODER = None


def reference_coder():
    """(callable get_small_copy(image, depth), kind, description)."""
    try:
        from oracle import ref_loader
        cls, _ = ref_loader.load_haar_coder()
        return cls().get_small_copy, "reference", ("oracle/_ref: the unmodified wicca/wavelet_coder.py HaarCoder.get_small_copy "
                                                   "(validate_image, cv2.copyMakeBorder padding, float32 NumPy levels)")
    except Exception as exc:  # noqa: BLE001
        from oracle import haar_oracle as ho
        return ho.haar_icon_fp32, "port", f"oracle.haar_icon_fp32, NumPy restatement of wavelet_coder.py:50-67 (oracle/_ref unavailable: {exc})"


def _cpu_task(job) -> float:
    """One get_small_copy(image, depth) of the reference on one full-size image; seconds."""
    idx, depth = job
    t0 = time.perf_counter()
    _CPU_CODER(_CPU_IMAGES[idx], depth)
    return time.perf_counter() - t0


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline_single(budget_images: int = 2) -> dict:
    """1 core (the reference is single-threaded), bounded sample: `budget_images` full-size images x depths 1-6."""
    from oracle import haar_oracle as ho
    fn, kind, what = reference_coder()
    secs = []
    for i in range(budget_images):
        img = ho.synthetic_image(i, H, W, CH)
        t0 = time.perf_counter()
        for d in DEPTHS:
            fn(img, d)
        secs.append(time.perf_counter() - t0)
    best = min(secs)
    return {"value": MP_PER_IMAGE / best, "unit": UNIT, "cores": 1, "kind": kind, "source": what,
            "sample": f"{budget_images} image(s) of {H}x{W}x3, get_small_copy at each of depths 1-6 (six calls per image, "
                      f"as classifying_tools.py:546-551 does), best image: {best:.2f} s"}


def run_reference(args) -> int:
    """bench.py --impl reference: the reference's own CPU implementation on every host core.  One step = M full-size
    (6393, 8284, 3) images x depths 1-6 = 6M get_small_copy calls spread over one worker process per core; M is the
    bounded sample (a step of the GPU arm is 30 such images), sized so that the run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    from oracle import haar_oracle as ho
    global _CPU_CODER
    _CPU_CODER, kind, what = reference_coder()
    cores = host_cores()
    workers = max(1, min(cores, 64))
    budget_s = float(os.environ.get("WICCA_REF_BUDGET_S", "200"))
    m_images = int(os.environ.get("WICCA_REF_IMAGES", "0")) or max(1, workers // 2)      # 3 calls per worker and step
    for i in range(m_images):
        _CPU_IMAGES.append(ho.synthetic_image(i, H, W, CH))
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        def step(m):
            jobs = [(i, d) for d in DEPTHS for i in range(m)]
            t0 = time.perf_counter()
            pool.map(_cpu_task, jobs, chunksize=1)
            return time.perf_counter() - t0
        cal = step(m_images)                                           # spins the pool up, and calibrates
        n_steps = max(1, args.steps + args.warmup)
        if cal * n_steps > budget_s and not os.environ.get("WICCA_REF_IMAGES"):
            m_images = max(1, min(m_images, int(m_images * budget_s / (cal * n_steps))))
        for _ in range(args.warmup):
            step(m_images)
        times = [step(m_images) for _ in range(args.steps)]
    total = sum(times)
    value = args.steps * m_images * MP_PER_IMAGE / total
    used = min(workers, 6 * m_images)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "source": what,
                             "sample": f"each step: {m_images} full-size images of {H}x{W}x3 x depths 1-6 = {6 * m_images} "
                                       f"get_small_copy calls over {workers} worker processes (one per host core); "
                                       "never cropped - a step of the GPU arm is 30 such images"},
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
def _pinned(lib, _capi, nbytes: int, device: int):
    """(ctypes pointer, flat uint8 view) of page-locked memory placed next to `device`."""
    p = C.c_void_p()
    _capi.check(lib.wicca_host_alloc_near(C.byref(p), max(1, nbytes), device), "wicca_host_alloc_near")
    return p, np.ctypeslib.as_array((C.c_uint8 * max(1, nbytes)).from_address(p.value))


def _peak():
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:  # noqa: BLE001
        return 6650.0, "6650 GB/s (of fallback)"


def host_link_ceiling(torch, dist, dev, world, host_flats, out_flats, h2d_bytes, d2h_bytes, barrier):
    """What the host side of the PCIe links gives this job, measured with plain large copies on every rank at once:
    (i) H2D only, (ii) one e2e step's traffic - h2d_bytes up on one stream, d2h_bytes down on another - with no
    kernels, no 2-D pitch, no per-image synchronisation.  An e2e step cannot finish faster than (ii).  The copies walk
    the SAME page-locked buffers as the e2e leg (its distinct source images in turn, one output block per image), so the
    host's caches see the same footprint: one 159 MB source copied thirty times would partly be served from them."""
    host_flat, out_flat = host_flats[0], out_flats[0]
    n_up = max(1, round(h2d_bytes / host_flat.numel()))
    n_down = max(1, round(d2h_bytes / out_flat.numel()))
    d_up = torch.empty(host_flat.numel(), dtype=torch.uint8, device=dev)
    d_down = torch.zeros(out_flat.numel(), dtype=torch.uint8, device=dev)
    s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(up: bool, down: bool) -> float:
        barrier()
        t0 = time.perf_counter()
        # the downward copies are paced by the upward ones (copy i down waits for copy i-1 up), as in the real
        # pipeline where an image's icons can only flow back after the image has arrived; unpaced, the 1:3 traffic
        # mix degenerates into a burst of writes followed by reads and the host memory system does worse
        n = max(n_up if up else 0, n_down if down else 0)
        for i in range(n):
            ev = None
            if up and i < n_up:
                with torch.cuda.stream(s_up):
                    d_up.copy_(host_flats[i % len(host_flats)], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(s_up)
            if down and i < n_down:
                with torch.cuda.stream(s_down):
                    if up and prev[0] is not None:
                        s_down.wait_event(prev[0])
                    out_flats[i % len(out_flats)].copy_(d_down, non_blocking=True)
            prev[0] = ev
        s_up.synchronize(); s_down.synchronize()
        prev[0] = None
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item())

    prev = [None]
    run(True, True)                                   # warm-up
    t_up = min(run(True, False) for _ in range(2))
    t_both = min(run(True, True) for _ in range(2))
    up_b, down_b = n_up * host_flat.numel(), n_down * out_flat.numel()
    del d_up, d_down
    return {"h2d_only_GBps": world * up_b / t_up / 1e9, "step_traffic_s": t_both,
            "h2d_GBps_with_d2h": world * up_b / t_both / 1e9, "d2h_GBps_with_h2d": world * down_b / t_both / 1e9,
            "note": "a probe of the same traffic with plain copies, best of two; run-to-run spread of either side is a few per "
                    "cent, so frac_of_ceiling reads 0.94-1.04",
            "method": f"every rank at once, the e2e leg's page-locked buffers ({len(host_flats)} sources in turn, {len(out_flats)} output "
                      f"blocks): {n_up} x {host_flat.numel() / 1e6:.0f} MB cudaMemcpyAsync up on one stream + {n_down} x "
                      f"{out_flat.numel() / 1e6:.0f} MB down on another (the byte counts of one e2e step), max over ranks"}


def extra_subbands(torch, lib, _capi, dev, local, stream, all_max, peak):
    """configs[2]: full forward + inverse sub-band transform of a 16384 x 16384 x 3 image, depths 1 / 3 / 6."""
    from wicca_b200.plan import pitch_bytes
    S = 16384
    pitch = pitch_bytes(S, 3)
    g = torch.Generator(device=dev); g.manual_seed(7)
    img = torch.randint(0, 256, (S, pitch), dtype=torch.uint8, device=dev, generator=g)
    coeffs = torch.empty((S, S, 3), dtype=torch.float32, device=dev)
    work = torch.empty((S * S * 3 * 5 // 16 + 64,), dtype=torch.float32, device=dev)
    rec = torch.empty((S, S, 3), dtype=torch.float32, device=dev)
    rows, px = [], S * S

    def timed(fn, reps=3, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return all_max(e0.elapsed_time(e1) / reps)

    for depth in (1, 3, 6):
        def fwd():
            _capi.check(lib.wicca_haar_forward_dev(img.data_ptr(), S, S, 3, pitch, depth, 1, 0.0, coeffs.data_ptr(),
                                                   work.data_ptr(), local, C.c_void_p(stream)), "forward_dev")

        def inv():
            _capi.check(lib.wicca_haar_inverse_dev(coeffs.data_ptr(), S, S, 3, depth, rec.data_ptr(), work.data_ptr(), local,
                                                   C.c_void_p(stream)), "inverse_dev")
        ms_f, ms_i = timed(fwd), timed(inv)
        coeffs.zero_(); rec.zero_()
        fwd(); inv(); torch.cuda.synchronize()
        err = 0.0
        for r0 in range(0, S, 2048):                   # max |rec - x| over the whole image, in bands (no 3 GB temporary)
            band = rec[r0:r0 + 2048] - img[r0:r0 + 2048, : S * 3].reshape(-1, S, 3).float()
            err = max(err, float(band.abs().max().item()))
        rows.append({"depth": depth, "forward_ms": ms_f, "inverse_ms": ms_i, "max_abs_rec_minus_x": err,
                     "forward_frac": 15 * px / ms_f / 1e6 / peak, "inverse_frac": 24 * px / ms_i / 1e6 / peak,
                     "forward_MP_per_s": px / ms_f / 1e3, "inverse_MP_per_s": px / ms_i / 1e3})
    del img, coeffs, work, rec
    torch.cuda.empty_cache()
    return {"workload": "configs[2]: 16384x16384x3 u8 -> float32 Mallat plane (all sub-bands kept) -> float32 image, per GPU",
            "algorithmic_bytes_per_px": {"forward": 15, "inverse": 24}, "rows": rows}


def extra_epilogue(torch, plan_cls, imgs, pitch, local, stream, dev, all_max, peak):
    """configs[3]: icon + cv2.resize(INTER_AREA) to 224 / 331 + preprocess_input('tf'), batches of 30 from the
    resident images; the icon never leaves the GPU (wicca_plan_resize_norm after wicca_plan_launch)."""
    n = len(imgs)
    rows = []

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return all_max(e0.elapsed_time(e1) / reps)

    for depth in (1, 3):
        plan = plan_cls(local, [t.data_ptr() for t in imgs], [H] * n, [W] * n, [pitch] * n, [depth])
        info = plan.info()
        _, ih, iw, _ = plan.icon_info(0, 0)
        for target in (224, 331):
            out = torch.empty((n, target, target, 3), dtype=torch.float32, device=dev)
            ms_all = timed(lambda: (plan.launch(stream), plan.resize_norm(0, target, target, 1, out.data_ptr(), 0, stream)))
            ms_epi = timed(lambda: plan.resize_norm(0, target, target, 1, out.data_ptr(), 0, stream))
            byt_epi = n * (ih * iw * 3 + target * target * 3 * 4)
            byt_all = info["bytes_read"] + n * target * target * 3 * 4      # the icon stays on the GPU: not counted
            rows.append({"depth": depth, "target": target, "icon_hw": [ih, iw], "icon_plus_epilogue_ms": ms_all,
                         "epilogue_ms": ms_epi, "epilogue_frac": byt_epi / ms_epi / 1e6 / peak,
                         "icon_plus_epilogue_frac": byt_all / ms_all / 1e6 / peak,
                         "MP_per_s": n * MP_PER_IMAGE / ms_all * 1e3, "batches_of_30_per_s": 1e3 / ms_all})
            del out
        plan.close()
    return {"workload": "configs[3]: 30 resident 8284x6393x3 images -> depth-d icon -> INTER_AREA 224 / 331 -> preprocess_input "
                        "'tf' -> (30, t, t, 3) float32, per GPU", "rows": rows}


def extra_sharded(torch, dist, lib, _capi, local, rank, world, barrier):
    """configs[4]: 130 ragged ~52 MP host images, depths 2-6, STRONG-scaled over the ranks (image i -> rank i % world,
    classifying_tools.py:312-321), icons gathered on the host through the shared-memory arena; wall clock."""
    from wicca_b200 import HaarCoder
    from wicca_b200.sharding import IconArena, shard_indices, sharded_small_copies
    n, distinct, depths = 130, 13, [2, 3, 4, 5, 6]
    rng = np.random.default_rng(0)
    shapes = [(H + int(rng.integers(-256, 257)), W + int(rng.integers(-256, 257)), 3) for _ in range(distinct)]
    mine = shard_indices(n, rank, world)
    need = sorted({i % distinct for i in mine} | (set(range(distinct)) if rank == 0 else set()))
    ptrs, arrs = {}, {}
    for k in need:                                   # image i is distinct image i % 13 (seeded by k: every rank agrees)
        h, w, _ = shapes[k]
        p, flat = _pinned(lib, _capi, h * w * 3, local)
        a = flat[: h * w * 3].reshape(h, w, 3)
        a[:] = np.random.default_rng(1000 + k).integers(0, 256, (h, w, 3), dtype=np.uint8)
        ptrs[k], arrs[k] = p, a
    coder = HaarCoder()
    coder.device = local
    transform = lambda images, ds, out=None: coder.get_small_copies_batch(images, ds, devices=[local], out=out)  # noqa: E731
    arena = IconArena([shapes[i % distinct] for i in range(n)], depths)
    get = lambda i: arrs[i % distinct]  # noqa: E731
    sharded_small_copies(get, min(n, 2 * world), depths, transform, arena=arena)        # warm-up: contexts, buffers
    times = []
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        icons = sharded_small_copies(get, n, depths, transform, arena=arena)            # ends with the gather barrier
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        times.append(float(dt.item()))
    res = None
    if rank == 0:
        # order / bit equality: everything the ranks gathered against this rank's own single-GPU run, and a sample
        # against the C oracle (checker only, outside the timed region)
        from oracle import c_oracle
        single = coder.get_small_copies_batch([get(i) for i in range(n)], depths, devices=[local])
        equal = all(np.array_equal(a, b) for ra, rb in zip(icons, single) for a, b in zip(ra, rb))
        for i in (0, 57, 129):
            exp = c_oracle.haar_icons_multi(get(i), depths)
            equal = equal and all(np.array_equal(a, b) for a, b in zip(icons[i], exp))
        mp_total = sum(shapes[i % distinct][0] * shapes[i % distinct][1] for i in range(n)) / 1e6
        best = min(times)
        res = {"workload": "configs[4]: 130 ragged ~52 MP page-locked host images (13 distinct shapes, H/W = 6393/8284 +- 256), "
                           "depths 2-6, image i on rank i % N, double-buffered H2D, icons gathered into a shared-memory arena",
               "scaling": "strong", "n_gpus": world, "seconds": best, "MP_per_s": mp_total / best,
               "h2d_GBps": mp_total * 3e6 / best / 1e9, "gathered_equals_single_gpu_and_oracle": bool(equal)}
        del single
    del icons
    arena.close()
    for p in ptrs.values():
        lib.wicca_host_free(p)
    return res


def extra_jpeg(torch, dist, lib, _capi, local, rank, world, barrier):
    """Row N2 as the e2e branch that is not bound by the host side of PCIe: 30 baseline JPEG files of 53 MP per GPU ->
    icons at depths 1-6 on the host (load_image + get_small_copy, data_loader.py:53-58 + wavelet_coder.py:50-67); only
    the 12 MB files cross the link."""
    try:
        import cv2
    except Exception as exc:  # noqa: BLE001
        return {"skipped": f"cv2 (the encoder of the synthetic files) not importable: {exc}"} if rank == 0 else None
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.empty((H, W, 3), np.float32)
    for c in range(3):
        img[:, :, c] = 128 + 90 * np.sin(xx / (37.0 + 9 * c) + c) + 70 * np.cos(yy / (23.0 + 5 * c) - c)
    del yy, xx
    img += np.random.default_rng(3).normal(0, 6, (H, W, 1)).astype(np.float32)
    img = np.clip(img, 0, 255).astype(np.uint8)
    ok, enc = cv2.imencode(".jpg", img[:, :, ::-1], [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                                    cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420])
    data = bytes(enc)
    n = BATCH
    nd = len(DEPTHS)
    shapes = [(lib.wicca_icon_dim(H, d), lib.wicca_icon_dim(W, d), 3) for d in DEPTHS]
    sizes = [int(np.prod(sh)) for sh in shapes]
    out_ptrs, outs = [], []
    for _ in range(n):                                  # page-locked, like the outputs of the raw-RGB e2e leg
        q, flat = _pinned(lib, _capi, sum(sizes), local)
        out_ptrs.append(q)
        row, off = [], 0
        for sh, nb in zip(shapes, sizes):
            row.append(flat[off:off + nb].reshape(sh)); off += nb
        outs.append(row)
    datas = (C.c_void_p * n)(*[C.cast(C.c_char_p(data), C.c_void_p).value] * n)
    lens = (C.c_size_t * n)(*[len(data)] * n)
    dsts = (C.c_void_p * (n * nd))(*[o.ctypes.data for per in outs for o in per])
    threads = max(2, min(16, host_cores() // world))
    hm = C.c_float()

    def step():
        _capi.check(lib.wicca_batch_icons_from_jpeg(datas, lens, n, (C.c_int * nd)(*DEPTHS), nd, 1, 0.0, dsts,
                                                    (C.c_int * 1)(local), 1, threads, C.byref(hm)), "wicca_batch_icons_from_jpeg")
    def timed_steps(fn):
        fn()
        ts = []
        for _ in range(2):
            barrier()
            t0 = time.perf_counter()
            fn()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            ts.append(float(dt.item()))
        return ts

    times = timed_steps(step)
    # the same files at depths 2-6 only (configs[4]'s depth set): 13 MB of icons per image flow back instead of 53 MB, so
    # neither direction of the host link is the limit any more
    deep = DEPTHS[1:]
    dsts_deep = (C.c_void_p * (n * len(deep)))(*[o.ctypes.data for per in outs for o in per[1:]])

    def step_deep():
        _capi.check(lib.wicca_batch_icons_from_jpeg(datas, lens, n, (C.c_int * len(deep))(*deep), len(deep), 1, 0.0, dsts_deep,
                                                    (C.c_int * 1)(local), 1, threads, C.byref(hm)), "wicca_batch_icons_from_jpeg")
    times_deep = timed_steps(step_deep)
    step()                                            # leave the depth 1-6 icons in place for the check below
    if rank != 0:
        for q in out_ptrs:
            lib.wicca_host_free(q)
        return None
    from oracle import c_oracle
    ref = cv2.cvtColor(cv2.imdecode(enc, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)      # what load_image returns
    exp = c_oracle.haar_icons_multi(ref, DEPTHS)
    equal = all(np.array_equal(a, b) for k in (0, n - 1) for a, b in zip(outs[k], exp))
    best = min(times)
    d2h = n * sum(sizes)
    del outs
    for q in out_ptrs:
        lib.wicca_host_free(q)
    return {"value": world * n * MP_PER_IMAGE / best, "unit": UNIT, "seconds_per_step": best,
            "h2d_bytes_per_step": n * len(data), "d2h_bytes_per_step": d2h,
            "workload": f"30 JPEG files per GPU ({H}x{W}, q90 4:2:0, {len(data) / 1e6:.1f} MB each, photo-like synthetic content) -> "
                        f"icons depths 1-6 on the host; Huffman decoding, IDCT, colour and icons on the GPU; {threads} host "
                        "threads per rank strip the byte stuffing",
            "equals_cv2_imdecode_then_oracle": bool(equal), "api": "wicca_batch_icons_from_jpeg",
            "depths_2_6": {"value": world * n * MP_PER_IMAGE / min(times_deep), "unit": UNIT, "seconds_per_step": min(times_deep),
                           "d2h_bytes_per_step": n * sum(sizes[1:])}}


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
    # stdout carries exactly ONE line, the JSON record.  NCCL prints its banner / NCCL_DEBUG lines on stdout (whatever
    # NCCL_DEBUG is set to is left alone: the driver reads those lines), so file descriptor 1 is pointed at stderr for
    # the duration of the run and the record is written to the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if distributed:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(x: float) -> float:
        if not distributed:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
    peak, peak_src = _peak()

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
    ms_per_step = all_max(e0.elapsed_time(e1)) / args.steps
    value = world * BATCH * MP_PER_IMAGE / (ms_per_step / 1e3)

    # ---- end to end through the C ABI with host buffers -------------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    n_host = min(BATCH, args.e2e_host_images)           # distinct page-locked host images, cycled through the batch
    host_ptrs, host_arrays, host_flats = [], [], []
    rng = np.random.default_rng(99 + rank)
    for i in range(n_host):
        p, flat = _pinned(lib, _capi, H * W * CH, local)
        arr = flat.reshape(H, W, CH)
        arr[:] = rng.integers(0, 256, (H, W, CH), dtype=np.uint8)
        host_ptrs.append(p); host_arrays.append(arr); host_flats.append(flat)
    icon_shapes = [(-(-H // (1 << d)), -(-W // (1 << d)), CH) for d in DEPTHS]
    icon_sizes = [int(np.prod(s)) for s in icon_shapes]
    out_ptrs, outs = [], []
    for _ in range(BATCH):                              # one page-locked block per image, the six icons back to back
        q, flat = _pinned(lib, _capi, sum(icon_sizes), local)
        out_ptrs.append(q)
        row, off = [], 0
        for s, nb in zip(icon_shapes, icon_sizes):
            row.append(flat[off:off + nb].reshape(s)); off += nb
        outs.append(row)
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
    e2e_s = all_max(time.perf_counter() - t0)
    e2e_value = world * BATCH * e2e_steps * MP_PER_IMAGE / e2e_s
    h2d_bytes = BATCH * H * W * CH
    d2h_bytes = BATCH * sum(icon_sizes)
    stage_ms = tim.as_dict()                               # sums over the 30 images of the last step

    # parity of EVERY icon of the last e2e step (30 images x 6 depths) against the C oracle, outside the timed region
    e2e_checked = 0
    if rank == 0:
        from concurrent.futures import ThreadPoolExecutor
        from oracle import c_oracle  # noqa: PLC0415  (checker only)
        with ThreadPoolExecutor(min(8, n_host)) as ex:
            expected = list(ex.map(lambda a: c_oracle.haar_icons_multi(a, DEPTHS), host_arrays))
        for i in range(BATCH):
            for got, exp in zip(outs[i], expected[i % n_host]):
                assert got.shape == exp.shape and np.array_equal(got, exp), f"e2e icon mismatch vs oracle (image {i})"
                e2e_checked += 1
        del expected
    # the host's measured limit for the same traffic, all ranks at once
    # (after the parity check: the probe's downward copies overwrite the icons)
    out_flat_t = [torch.from_numpy(np.ctypeslib.as_array((C.c_uint8 * sum(icon_sizes)).from_address(q.value))) for q in out_ptrs]
    ceiling = host_link_ceiling(torch, dist, dev, world, [torch.from_numpy(f) for f in host_flats], out_flat_t, h2d_bytes, d2h_bytes,
                                barrier)
    ceiling_value = world * BATCH * MP_PER_IMAGE / ceiling["step_traffic_s"]

    for p in host_ptrs + out_ptrs:
        lib.wicca_host_free(p)
    del host_arrays, host_flats, outs, out_flat_t
    t_clk2 = time.perf_counter()

    # ---- the other configs of BASELINE.json, bounded (about 20 s) ---------------------------
    extra = {}

    def guarded(name, fn):
        """The headline record must not be lost to a failure in a side measurement: report it instead."""
        try:
            return fn()
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] extra '{name}' failed on rank {rank}: {exc!r}", file=sys.stderr, flush=True)
            return {"error": repr(exc)[:300]}

    if not args.no_extra:
        extra["configs3_icon_resize_norm"] = guarded("configs3", lambda: extra_epilogue(torch, IconPlan, imgs, pitch, local, stream, dev, all_max, peak))
    plan.close()
    del imgs
    torch.cuda.empty_cache()
    e2e_jpeg = None
    if not args.no_extra:
        extra["configs2_subband_round_trip"] = guarded("configs2", lambda: extra_subbands(torch, lib, _capi, dev, local, stream, all_max, peak))
        extra["configs4_sharded_ragged_batch"] = guarded("configs4", lambda: extra_sharded(torch, dist, lib, _capi, local, rank, world, barrier))
        e2e_jpeg = guarded("e2e_jpeg", lambda: extra_jpeg(torch, dist, lib, _capi, local, rank, world, barrier))

    sampler.stop()
    line = None
    if rank == 0:
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
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps, "distinct_host_images": n_host,
                    "api": "wicca_batch_icons_u8 (page-locked host images and icons; upload slots per GPU so H2D(i+1) overlaps kernel/D2H(i))",
                    "stage_ms_sum_over_images": stage_ms,
                    "h2d_GBps": world * h2d_bytes * e2e_steps / e2e_s / 1e9,
                    "host_ceiling_GBps": ceiling["h2d_GBps_with_d2h"], "host_ceiling_MP_per_s": ceiling_value,
                    "frac_of_ceiling": e2e_value / ceiling_value, "host_ceiling": ceiling,
                    "icons_checked_against_oracle": e2e_checked},
            "clocks": sampler.summary(t_clk0, t_clk2),
        }
        if e2e_jpeg is not None:
            line["e2e_jpeg"] = e2e_jpeg
        if extra:
            line["extra"] = extra
        if cpu is not None:
            line["cpu_baseline"] = cpu
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
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
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[2..4] / e2e_jpeg block")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    raise SystemExit(main())
