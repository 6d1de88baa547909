#!/usr/bin/env python3
"""Per-row measurements for SURVEY.md section 8 (everything except the headline line, which is bench.py):
single-depth icons, full sub-band forward/inverse (configs[2]), fused resize+normalise epilogue
(configs[3]) and the sharded host batch with pinned ingest (configs[4]).  Device times are CUDA
events around K back-to-back launches; one JSON line per measurement."""
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np
import torch

from wicca_b200 import HaarCoder, _capi
from wicca_b200.plan import IconPlan, pitch_bytes

PEAK = 6544.7
try:
    PEAK = float(json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass
H, W = 6393, 8284
dev = torch.device("cuda:0")
lib = _capi.load()
stream = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def emit(**kw):
    print(json.dumps(kw), flush=True)


def row_icons():
    n = 30
    pitch = pitch_bytes(W, 3)
    g = torch.Generator(device=dev); g.manual_seed(0)
    imgs = [torch.randint(0, 256, (H, pitch), dtype=torch.uint8, device=dev, generator=g) for _ in range(n)]
    for ds in ([1], [2], [3], [4], [5], [6], [2, 3, 4, 5, 6], [1, 2, 3, 4, 5, 6]):
        plan = IconPlan(0, [t.data_ptr() for t in imgs], [H] * n, [W] * n, [pitch] * n, ds)
        info = plan.info()
        ms = timed(lambda: plan.launch(stream), reps=20)
        byt = info["bytes_read"] + info["bytes_written"]
        emit(row="A3 icon", config=f"30 x ({H},{W},3), depths {ds}, device-resident, one launch", ms=ms,
             MP_per_s=n * H * W / ms / 1e3, GBps=byt / ms / 1e6, frac_of_measured_peak=byt / ms / 1e6 / PEAK,
             algorithmic_bytes=byt)
        if ds == [1, 2, 3, 4, 5, 6]:
            # row N1: the source-image branch, cv2.resize(image, (224, 224), INTER_AREA), on the resident images
            out_s = torch.empty((n, 224, 224, 3), dtype=torch.float32, device=dev)
            ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in imgs])
            def src_resize():
                _capi.check(lib.wicca_resize_norm_dev(ptrs, (C.c_int * n)(*[H] * n), (C.c_int * n)(*[W] * n), (C.c_int64 * n)(*[pitch] * n),
                                                      n, 224, 224, 1, out_s.data_ptr(), None, 0, C.c_void_p(stream)), "resize_norm_dev")
            ms_s = timed(src_resize, reps=20, warm=2)
            emit(row="N1 source-image resize", config=f"30 x ({H},{W},3) -> 224x224 tf, device-resident", ms=ms_s,
                 MP_per_s=n * H * W / ms_s / 1e3, GBps=n * H * W * 3 / ms_s / 1e6, frac_of_measured_peak=n * H * W * 3 / ms_s / 1e6 / PEAK)
            # fused epilogue on the resident icons (configs[3])
            for target in (224, 331):
                out = torch.empty((n, target, target, 3), dtype=torch.float32, device=dev)
                for k, d in enumerate(ds):
                    _, ih, iw, _ = plan.icon_info(0, k)
                    ms_e = timed(lambda: plan.resize_norm(k, target, target, 1, out.data_ptr(), 0, stream), reps=5, warm=2)
                    byt_e = n * (ih * iw * 3 + target * target * 3 * 4)
                    emit(row="A5+A6 epilogue", config=f"30 icons of depth {d} ({ih}x{iw}) -> {target}x{target} tf, device-resident",
                         ms=ms_e, GBps=byt_e / ms_e / 1e6, frac_of_measured_peak=byt_e / ms_e / 1e6 / PEAK,
                         algorithmic_bytes=byt_e, batches_of_30_per_s=1e3 / ms_e)
                # icon + epilogue for one depth, as a classifier would consume it (depth 3, both sizes)
            plan3 = IconPlan(0, [t.data_ptr() for t in imgs], [H] * n, [W] * n, [pitch] * n, [3])
            out = torch.empty((n, 224, 224, 3), dtype=torch.float32, device=dev)
            ms_f = timed(lambda: (plan3.launch(stream), plan3.resize_norm(0, 224, 224, 1, out.data_ptr(), 0, stream)), reps=10)
            emit(row="A3+A5+A6 fused", config="30 images -> depth-3 icon -> 224x224 tf batch, device-resident", ms=ms_f,
                 MP_per_s=n * H * W / ms_f / 1e3)
            plan3.close()
        plan.close()
    del imgs
    torch.cuda.empty_cache()


def row_deep():
    """Depths 7-8 and other channel counts (verdict r1 item 6): one 53 MP image, device-resident, through the
    device-pointer ABI (wicca_haar_icons_multi_dev)."""
    g = torch.Generator(device=dev); g.manual_seed(2)
    for ch, depth_sets in ((3, ([8], [7], [7, 8], [1, 2, 3, 4, 5, 6, 7, 8], [6])), (4, ([1], [3], [6], [8])), (1, ([3], [8]))):
        pitch = pitch_bytes(W, ch)
        img = torch.randint(0, 256, (H, pitch), dtype=torch.uint8, device=dev, generator=g)
        host = img[:, : W * ch].reshape(H, W, ch).cpu().numpy()
        for ds in depth_sets:
            outs, ptrs, pitches = [], [], []
            for d in ds:
                oh, ow = -(-H // (1 << d)), -(-W // (1 << d))
                op = -(-ow * ch // 128) * 128
                t = torch.zeros((oh, op), dtype=torch.uint8, device=dev)
                outs.append((t, oh, ow)); ptrs.append(t.data_ptr()); pitches.append(op)
            n = len(ds)

            def run():
                _capi.check(lib.wicca_haar_icons_multi_dev(img.data_ptr(), H, W, ch, pitch, (C.c_int * n)(*ds), n, 1, 0.0,
                                                           (C.c_void_p * n)(*ptrs), (C.c_int64 * n)(*pitches), 0,
                                                           C.c_void_p(stream)), "icons_multi_dev")
            ms = timed(run, reps=10, warm=2)
            from oracle import c_oracle
            exp = c_oracle.haar_icons_multi(host, ds)
            ok = all(np.array_equal(t[:, : ow * ch].reshape(oh, ow, ch).cpu().numpy(), e) for (t, oh, ow), e in zip(outs, exp))
            byt = H * W * ch + sum(oh * ow * ch for _, oh, ow in outs)
            emit(row="A3 beyond the one-pass domain", config=f"one ({H},{W},{ch}) image, depths {ds}, device-resident", ms=ms,
                 MP_per_s=H * W / ms / 1e3, GBps_if_one_pass=byt / ms / 1e6, frac_of_measured_peak_if_one_pass=byt / ms / 1e6 / PEAK,
                 equals_oracle=bool(ok), lib=os.environ.get("WICCA_B200_LIB", "in-tree"))
        del img
    torch.cuda.empty_cache()


def row_subbands():
    S = 16384
    pitch = pitch_bytes(S, 3)
    g = torch.Generator(device=dev); g.manual_seed(1)
    img = torch.randint(0, 256, (S, pitch), dtype=torch.uint8, device=dev, generator=g)
    coeffs = torch.empty((S, S, 3), dtype=torch.float32, device=dev)
    work = torch.empty((S * S * 3 * 5 // 16 + 64,), dtype=torch.float32, device=dev)
    rec = torch.empty((S, S, 3), dtype=torch.float32, device=dev)
    for depth in [int(x) for x in os.environ.get("WICCA_ROWS_SUBBAND_DEPTHS", "1,3,6").split(",")]:
        def fwd():
            _capi.check(lib.wicca_haar_forward_dev(img.data_ptr(), S, S, 3, pitch, depth, 1, 0.0, coeffs.data_ptr(),
                                                   work.data_ptr(), 0, C.c_void_p(stream)), "forward_dev")

        def inv():
            _capi.check(lib.wicca_haar_inverse_dev(coeffs.data_ptr(), S, S, 3, depth, rec.data_ptr(), work.data_ptr(), 0,
                                                   C.c_void_p(stream)), "inverse_dev")
        ms_f = timed(fwd, reps=5, warm=2)
        ms_i = timed(inv, reps=5, warm=2)
        fwd(); inv(); torch.cuda.synchronize()
        err = float((rec[:, :, :] - img[:, : S * 3].reshape(S, S, 3).float()).abs().max().item())
        px = S * S
        emit(row="A4 forward", config=f"{S}x{S}x3 u8 -> fp32 Mallat plane, depth {depth}", ms=ms_f, MP_per_s=px / ms_f / 1e3,
             GBps=15 * px / ms_f / 1e6, frac_of_measured_peak=15 * px / ms_f / 1e6 / PEAK, algorithmic_bytes=15 * px)
        emit(row="A4 inverse", config=f"{S}x{S}x3 fp32 plane -> fp32 image, depth {depth}", ms=ms_i, MP_per_s=px / ms_i / 1e3,
             GBps=24 * px / ms_i / 1e6, frac_of_measured_peak=24 * px / ms_i / 1e6 / PEAK, algorithmic_bytes=24 * px,
             round_trip_max_abs_error=err)
    del img, coeffs, work, rec
    torch.cuda.empty_cache()


def row_batch():
    """configs[4]: 130 ragged ~52 MP images, depths 2-6, pinned host ingest, all visible GPUs."""
    ndev = lib.wicca_device_count()
    n, distinct = 130, 13
    rng = np.random.default_rng(0)
    shapes = [(H + int(rng.integers(-256, 257)), W + int(rng.integers(-256, 257))) for _ in range(distinct)]
    ptrs, arrs = [], []
    for (h, w) in shapes:
        p = C.c_void_p()
        _capi.check(lib.wicca_host_alloc(C.byref(p), h * w * 3), "host_alloc")
        a = np.ctypeslib.as_array((C.c_uint8 * (h * w * 3)).from_address(p.value)).reshape(h, w, 3)
        a[:] = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ptrs.append(p); arrs.append(a)
    coder = HaarCoder()
    images = [arrs[i % distinct] for i in range(n)]
    depths = [2, 3, 4, 5, 6]
    for devices in ([0], list(range(ndev))) if ndev > 1 else ([0],):
        coder.get_small_copies_batch(images[: 6 * len(devices)], depths, devices=devices)   # warm-up: every device, every upload slot
        t0 = time.perf_counter()
        out = coder.get_small_copies_batch(images, depths, devices=devices)
        dt = time.perf_counter() - t0
        mp = sum(a.shape[0] * a.shape[1] for a in images) / 1e6
        byt = sum(a.nbytes for a in images)
        emit(row="(e) sharded host batch", config=f"130 ragged ~52 MP images, depths 2-6, pinned host -> icons on host, {len(devices)} GPU(s)",
             s=dt, MP_per_s=mp / dt, h2d_GBps=byt / dt / 1e9, stage_ms_sum=coder.last_timing)
    from oracle import haar_oracle as ho
    assert np.array_equal(out[5][1], ho.haar_icon_blocksum(images[5], 3))
    # configs[3] end to end: host images -> (batch_images, batch_icons) float32 for a classifier, batch of 30
    for depth, target in ((3, 224), (3, 331), (2, 224)):
        coder.classifier_batches(images[:4], depth, (target, target), "tf")
        t0 = time.perf_counter()
        bi, bc = coder.classifier_batches(images[:30], depth, (target, target), "tf")
        dt = time.perf_counter() - t0
        mp = sum(a.shape[0] * a.shape[1] for a in images[:30]) / 1e6
        emit(row="A3+A5+A6+N1 one call", config=f"30 pinned host images -> depth-{depth} icons + source -> two ({target},{target},3) tf batches",
             s=dt, MP_per_s=mp / dt, batches_of_30_per_s=1 / dt, stage_ms_sum=coder.last_timing)
    # row N3: the reference's classifier x depth loops (9 distinct classifier inputs of the demo x depths 2-6 = 45
    # _get_img_batch calls per batch) from ONE upload per image
    targets = [((224, 224), "tf"), ((224, 224), "caffe"), ((224, 224), "torch"), ((224, 224), "identity"),
               ((240, 240), "identity"), ((260, 260), "identity"), ((299, 299), "tf"), ((331, 331), "tf"), ((300, 300), "identity")]
    coder.classifier_batches_multi(images[:4], depths, targets)
    t0 = time.perf_counter()
    res = coder.classifier_batches_multi(images[:30], depths, targets)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    one = coder.classifier_batches(images[:30], 3, (299, 299), "tf")
    dt_one = time.perf_counter() - t0
    assert np.array_equal(one[1], res[6][1][3]) and np.array_equal(one[0], res[6][0])
    mp = sum(a.shape[0] * a.shape[1] for a in images[:30]) / 1e6
    emit(row="N3 one upload, all classifier inputs", config=f"30 pinned host images -> {len(targets)} targets x depths 2-6 = "
         f"{len(targets) * len(depths)} icon batches + {len(targets)} source batches (fp32)", s=dt, MP_per_s=mp / dt,
         one_target_one_depth_call_s=dt_one, same_work_by_single_calls_s=dt_one * len(targets) * len(depths),
         stage_ms_sum=coder.last_timing)
    for p in ptrs:
        lib.wicca_host_free(p)


def row_jpeg():
    """Row N2: load_image (cv2.imread + BGR2RGB, data_loader.py:53-58) from a 53 MP baseline JPEG."""
    import cv2
    from concurrent.futures import ThreadPoolExecutor
    from wicca_b200 import decode_jpeg, icons_from_jpeg
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.stack([128 + 90 * np.sin(xx / (37.0 + 9 * c) + c) + 70 * np.cos(yy / (23.0 + 5 * c) - c) for c in range(3)], -1)
    img = np.clip(img + rng.normal(0, 6, img.shape).astype(np.float32), 0, 255).astype(np.uint8)
    del yy, xx
    for quality, sampling, tag in ((90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "4:2:0"), (95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "4:4:4")):
        ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sampling])
        data = bytes(enc)
        ref_t = []
        for _ in range(3):
            t0 = time.perf_counter(); ref = cv2.cvtColor(cv2.imdecode(enc, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB); ref_t.append(time.perf_counter() - t0)
        decode_jpeg(data)
        ours_t, tm = [], {}
        for _ in range(3):
            t0 = time.perf_counter(); got = decode_jpeg(data, timing=tm); ours_t.append(time.perf_counter() - t0)
        assert np.array_equal(got, ref)
        icon_t = []
        for _ in range(3):
            t0 = time.perf_counter(); icons = icons_from_jpeg(data, [1, 2, 3, 4, 5, 6], timing=tm); icon_t.append(time.perf_counter() - t0)
        from oracle import haar_oracle as ho
        assert np.array_equal(icons[2], ho.haar_icon_blocksum(ref, 3))
        emit(row="N2 JPEG ingest, one file", config=f"({H},{W},3) baseline JPEG q{quality} {tag}, {len(data) / 1e6:.1f} MB",
             cv2_imdecode_bgr2rgb_s=min(ref_t), decode_to_host_rgb_s=min(ours_t), jpeg_to_icons_depths_1_6_s=min(icon_t),
             host_stage_ms=tm["host_decode_ms"], scan_upload_ms=tm["h2d_ms"], huffman_idct_colour_icon_kernels_ms=tm["kernel_ms"], icons_d2h_ms=tm["d2h_ms"], MP_per_s_to_icons=H * W / 1e6 / min(icon_t),
             MP_per_s_cv2=H * W / 1e6 / min(ref_t))
        # many files, all host cores: the reference's loader in a thread pool (cv2 releases the GIL) against the batch entry
        n, cores = 32, len(os.sched_getaffinity(0))
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(lambda _: cv2.cvtColor(cv2.imdecode(enc, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB).shape, range(n)))
        dt_ref = time.perf_counter() - t0
        depths = [2, 3, 4, 5, 6]
        outs = [[np.empty((lib.wicca_icon_dim(H, d), lib.wicca_icon_dim(W, d), 3), np.uint8) for d in depths] for _ in range(n)]
        datas = (C.c_void_p * n)(*[C.cast(C.c_char_p(data), C.c_void_p).value] * n)
        lens = (C.c_size_t * n)(*[len(data)] * n)
        dsts = (C.c_void_p * (n * len(depths)))(*[o.ctypes.data for per in outs for o in per])
        ndev = lib.wicca_device_count()
        hm = C.c_float()
        dt = 1e9
        for rep in range(4):                                 # the first pass allocates the per-thread contexts
            t0 = time.perf_counter()
            _capi.check(lib.wicca_batch_icons_from_jpeg(datas, lens, n, (C.c_int * len(depths))(*depths), len(depths), 1, 0.0, dsts,
                                                        (C.c_int * ndev)(*range(ndev)), ndev, 0, C.byref(hm)), "batch_icons_from_jpeg")
            if rep:
                dt = min(dt, time.perf_counter() - t0)
        assert np.array_equal(outs[7][1], ho.haar_icon_blocksum(ref, 3))
        if quality == 90:
            # the whole of _get_img_batch for every classifier input and depth, from file bytes (N2 + N3)
            import tempfile
            tdir = tempfile.mkdtemp()
            paths = []
            for k in range(30):
                pth = os.path.join(tdir, f"f{k}.jpg")
                with open(pth, "wb") as fh:
                    fh.write(data)
                paths.append(pth)
            targets = [((224, 224), "tf"), ((224, 224), "caffe"), ((224, 224), "torch"), ((224, 224), "identity"),
                       ((240, 240), "identity"), ((260, 260), "identity"), ((299, 299), "tf"), ((331, 331), "tf"), ((300, 300), "identity")]
            coder = HaarCoder()
            coder.classifier_batches_multi_from_files(paths[:4], depths, targets)
            t0 = time.perf_counter()
            res = coder.classifier_batches_multi_from_files(paths, depths, targets)
            dt_all = time.perf_counter() - t0
            emit(row="N2+N3 files to all classifier inputs", config=f"30 JPEG files (53 MP each) -> {len(targets)} targets x depths 2-6 = "
                 f"{len(targets) * len(depths)} icon batches + {len(targets)} source batches (fp32)", s=dt_all,
                 MP_per_s=30 * H * W / 1e6 / dt_all, stage_ms_sum=coder.last_timing)
            for pth in paths:
                os.remove(pth)
            os.rmdir(tdir)
        emit(row="N2 JPEG ingest, 32 files", config=f"32 x ({H},{W},3) JPEG q{quality} {tag} -> icons depths 2-6, {cores} host threads, {ndev} GPU(s)",
             s=dt, MP_per_s=n * H * W / 1e6 / dt, files_per_s=n / dt, host_stage_ms_sum=hm.value,
             cv2_thread_pool_decode_only_s=dt_ref, cv2_MP_per_s=n * H * W / 1e6 / dt_ref)


def row_wavelets():
    """Row N4: longer orthogonal filters behind the WaveletCoder interface (host image in, icon out)."""
    from oracle import fir_oracle as fo
    from wicca_b200 import OrthogonalWaveletCoder
    img = np.random.default_rng(5).integers(0, 256, (H, W, 3), dtype=np.uint8)
    small = img[:1024, :1024]
    for name in ("db2", "db4", "coif1"):
        coder = OrthogonalWaveletCoder(name)
        coder.get_small_copy(img, 3)
        t0 = time.perf_counter(); icon = coder.get_small_copy(img, 3); dt = time.perf_counter() - t0
        stages = coder.last_timing
        t0 = time.perf_counter(); exp = fo.wavelet_icon(small, 3, name); dt_cpu = time.perf_counter() - t0
        assert np.array_equal(coder.get_small_copy(small, 3), exp)
        emit(row="N4 orthogonal wavelet icon", config=f"get_small_copy(({H},{W},3) pageable ndarray, depth 3), {name} ({len(coder.taps)} taps)",
             ms=dt * 1e3, MP_per_s=H * W / 1e6 / dt, stage_ms=stages,
             cpu_oracle_MP_per_s_1_core=small.shape[0] * small.shape[1] / 1e6 / dt_cpu)


def row_oneshot():
    """The call the reference's callers make: one get_small_copy on a host array (pageable vs pinned)."""
    coder = HaarCoder()
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    p = C.c_void_p()
    _capi.check(lib.wicca_host_alloc(C.byref(p), H * W * 3), "host_alloc")
    pin = np.ctypeslib.as_array((C.c_uint8 * (H * W * 3)).from_address(p.value)).reshape(H, W, 3)
    pin[:] = img
    for name, arr in (("pageable", img), ("pinned", pin)):
        for d in (3, [1, 2, 3, 4, 5, 6]):
            f = (lambda: coder.get_small_copy(arr, d)) if isinstance(d, int) else (lambda: coder.get_small_copies(arr, d))
            f(); f()
            ts = []
            for _ in range(5):
                t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
            emit(row="A3 one-shot host call", config=f"get_small_copy/ies(({H},{W},3) {name} ndarray, depth {d})", ms=1e3 * min(ts),
                 MP_per_s=H * W / 1e6 / min(ts), stage_ms=coder.last_timing)
    del pin
    lib.wicca_host_free(p)


def cpu_side():
    from oracle import haar_oracle as ho, resize_oracle as ro
    img = ho.synthetic_image(0, 4096, 4096, 3)
    t0 = time.perf_counter(); ho.haar_icon_fp32(img, 3); t1 = time.perf_counter()
    emit(row="CPU oracle", config="configs[0]: 4096x4096x3 depth 3, NumPy port, 1 core", s=t1 - t0, MP_per_s=16.777 / (t1 - t0))
    t0 = time.perf_counter(); co = ho.haar_forward(img, 3); t1 = time.perf_counter(); ho.haar_inverse(co); t2 = time.perf_counter()
    emit(row="CPU oracle", config="A4 forward / inverse 4096x4096x3 depth 3, NumPy, 1 core", forward_s=t1 - t0, inverse_s=t2 - t1,
         forward_MP_per_s=16.777 / (t1 - t0), inverse_MP_per_s=16.777 / (t2 - t1))
    icon = ho.synthetic_image(1, 800, 1036, 3)
    t0 = time.perf_counter(); ro.preprocess_input(ro.resize_area(icon, 224, 224)[None], "tf"); t1 = time.perf_counter()
    emit(row="CPU oracle", config="A5+A6 one 800x1036 icon -> 224 tf (NumPy restatement; cv2 itself takes ~4 ms)", s=t1 - t0)


if __name__ == "__main__":
    which = sys.argv[1:] or ["icons", "deep", "subbands", "batch", "jpeg", "wavelets", "oneshot", "cpu"]
    if "icons" in which:
        row_icons()
    if "deep" in which:
        row_deep()
    if "subbands" in which:
        row_subbands()
    if "batch" in which:
        row_batch()
    if "jpeg" in which:
        row_jpeg()
    if "wavelets" in which:
        row_wavelets()
    if "oneshot" in which:
        row_oneshot()
    if "cpu" in which:
        cpu_side()
