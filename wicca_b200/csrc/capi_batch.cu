// capi_batch.cu - sharded host batch: images are independent units (the per-image loop of
// wicca/classifying_tools.py:312-321), so image i goes to devices[i % n_devices] and nothing
// crosses between GPUs.  One worker thread per device; three upload slots (stream + buffers) per
// worker so the H2D copies of images j+1, j+2 overlap the kernel and icon D2H of image j.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "host_common.h"
#include "kernels.h"
#include "resize_tables.h"

using namespace wicca;

namespace {

struct BatchArgs {
    const uint8_t* const* srcs; const int* Hs; const int* Ws; const int64_t* strides;
    int n_images, C; const int* depths; int n_depths; int border_type, bconst;
    uint8_t* const* dsts;
};

struct WorkerResult {
    int rc = 0;
    std::string msg;
    double h2d = 0, kernel = 0, d2h = 0, total = 0;
};

void add_times(WorkerResult& r, Ctx& c) {
    float a = 0, b = 0, d = 0, e = 0;
    cudaEventElapsedTime(&a, c.ev[0], c.ev[1]);
    cudaEventElapsedTime(&b, c.ev[1], c.ev[2]);
    cudaEventElapsedTime(&d, c.ev[2], c.ev[3]);
    cudaEventElapsedTime(&e, c.ev[0], c.ev[3]);
    r.h2d += a; r.kernel += b; r.d2h += d; r.total += e;
}

// Upload slots per worker: slot s owns a stream, a device image buffer and icon buffers.  Three keep the H2D engine
// busy across the host's synchronise-and-enqueue gap (two leave a bubble whenever the kernel + icon read-back of one
// slot ends while the other slot's upload is already draining).  WICCA_UPLOAD_SLOTS=2..4 overrides.
int upload_slots() {
    int n = 3;
    if (const char* e = getenv("WICCA_UPLOAD_SLOTS")) n = atoi(e);
    return n < 2 ? 2 : (n > 4 ? 4 : n);
}

int worker_body(const BatchArgs& a, int device, int first, int step, WorkerResult& res) {
    ScopedAffinity bind(device);         // this worker (and the buffers it allocates) stays on the GPU's NUMA node
    ScopedLinkShared shared;             // uploads and read-backs of different images overlap: flat copies over the link
    const int n_slots = upload_slots();
    CtxLease slot[4];
    bool busy[4] = {false, false, false, false};
    for (int s = 0; s < n_slots; ++s) {
        int rc = acquire_ctx(device, &slot[s].c);
        if (rc) return rc;
    }
    // one allocation per slot, sized for the largest image of this worker (ragged batches would
    // otherwise grow the buffer - and synchronise the device - several times)
    size_t max_src = 0, max_tight = 0;
    for (int i = first; i < a.n_images; i += step) {
        max_src = std::max(max_src, (size_t)wicca_pitch_bytes(a.Ws[i], a.C) * a.Hs[i] + 256);
        if (is_pinned_host(a.srcs[i])) max_tight = std::max(max_tight, (size_t)a.Ws[i] * a.C * a.Hs[i]);   // lands flat in d_stage first
    }
    for (int s = 0; s < n_slots; ++s) {
        WICCA_CUDA(slot[s].c->d_src.reserve(max_src));
        if (max_tight) WICCA_CUDA(slot[s].c->d_stage.reserve(max_tight));
    }
    int j = 0;
    for (int i = first; i < a.n_images; i += step) {
        const int H = a.Hs[i], W = a.Ws[i];
        const int64_t rowb = (int64_t)W * a.C;
        const int64_t stride = (a.strides && a.strides[i]) ? a.strides[i] : rowb;
        uint8_t* const* dsts = a.dsts + (size_t)i * a.n_depths;
        bool device_work = false;
        for (int k = 0; k < a.n_depths; ++k) {
            if (a.depths[k] <= 0) {
                for (int y = 0; y < H; ++y) memcpy(dsts[k] + (size_t)y * rowb, a.srcs[i] + (size_t)y * stride, (size_t)rowb);
            } else {
                device_work = true;
            }
        }
        if (!device_work) continue;
        const int s = j % n_slots;
        ++j;
        Ctx& c = *slot[s].c;
        if (busy[s]) {
            WICCA_CUDA(cudaStreamSynchronize(c.stream));
            c.flush_pending();
            add_times(res, c);
        }
        const int64_t pitch = wicca_pitch_bytes(W, a.C);
        WICCA_CUDA(c.d_src.reserve((size_t)pitch * H + 256));
        WICCA_CUDA(cudaEventRecord(c.ev[0], c.stream));
        int rc = upload_image_async(c, a.srcs[i], H, rowb, stride, pitch);
        if (rc) { cudaStreamSynchronize(c.stream); c.pending.clear(); return rc; }
        WICCA_CUDA(cudaEventRecord(c.ev[1], c.stream));
        rc = icons_from_resident(c, H, W, a.C, pitch, a.depths, a.n_depths, a.border_type, a.bconst, dsts);
        if (rc) { cudaStreamSynchronize(c.stream); c.pending.clear(); return rc; }
        busy[s] = true;
    }
    // drain in issue order
    for (int k = 0; k < n_slots; ++k) {
        const int s = (j + k) % n_slots;
        if (busy[s]) {
            WICCA_CUDA(cudaStreamSynchronize(slot[s].c->stream));
            slot[s].c->flush_pending();
            add_times(res, *slot[s].c);
        }
    }
    return 0;
}

}  // namespace

extern "C" int wicca_batch_icons_u8(const uint8_t* const* srcs, const int* Hs, const int* Ws, const int64_t* strides,
                                    int n_images, int C, const int* depths, int n_depths, int border_type,
                                    double border_const, uint8_t* const* dsts, const int* devices, int n_devices,
                                    wicca_timing* t) {
    if (t) memset(t, 0, sizeof(*t));
    if (n_images < 0) return fail(WICCA_EINVAL, "negative image count");
    if (n_images == 0) return 0;
    if (!srcs || !Hs || !Ws || !dsts) return fail(WICCA_EINVAL, "null array");
    if (n_devices <= 0) return fail(WICCA_EDEVICE, "need at least one device");
    for (int i = 0; i < n_images; ++i) {
        int rc = validate_icon_args(srcs[i], Hs[i], Ws[i], C, depths, n_depths, border_type);
        if (rc) return rc;
        const int64_t rowb = (int64_t)Ws[i] * C;
        if (strides && strides[i] && strides[i] < rowb) return fail(WICCA_EINVAL, "strides[%d] < W*C", i);
        for (int k = 0; k < n_depths; ++k)
            if (!dsts[(size_t)i * n_depths + k]) return fail(WICCA_EINVAL, "dsts[%d][%d] is NULL", i, k);
    }
    std::vector<int> devs(n_devices);
    for (int k = 0; k < n_devices; ++k) {
        devs[k] = devices ? devices[k] : k;
        int rc = check_device(devs[k]);
        if (rc) return rc;
    }
    BatchArgs a{srcs, Hs, Ws, strides, n_images, C, depths, n_depths, border_type, saturate_u8(border_const), dsts};
    const int nw = n_devices < n_images ? n_devices : n_images;
    std::vector<WorkerResult> results(nw);
    std::vector<std::thread> threads;
    for (int k = 0; k < nw; ++k)
        threads.emplace_back([&, k] {
            results[k].rc = worker_body(a, devs[k], k, nw, results[k]);
            if (results[k].rc) results[k].msg = last_error_ref();
        });
    for (auto& th : threads) th.join();
    wicca_timing sum = {0, 0, 0, 0};
    for (auto& r : results) {
        if (r.rc) { last_error_ref() = r.msg; return r.rc; }
        sum.h2d_ms += (float)r.h2d; sum.kernel_ms += (float)r.kernel; sum.d2h_ms += (float)r.d2h; sum.total_ms += (float)r.total;
    }
    if (t) *t = sum;
    return 0;
}

// ------------------------------------------------------------------------------------------
// Classifier-ready batches straight from host images: the body of ClassifierProcessor._get_img_batch
// (classifying_tools.py:312-323: resize(image), get_small_copy(image, depth), resize(icon), np.stack)
// followed by preprocess_input + float32 cast (:286-287), with the icon never leaving the device.
// ------------------------------------------------------------------------------------------
namespace {

struct ClsArgs {
    const uint8_t* const* srcs; const int* Hs; const int* Ws; const int64_t* strides; int n_images;
    const size_t* jpeg_lens;                  // not NULL: srcs[i] are JPEG files of jpeg_lens[i] bytes, decoded on the GPU
    const int* depths; int n_depths; int border_type, bconst;
    const wicca_target* targets; int n_targets;
    float* const* dst_icons; float* const* dst_images;
};

int cls_worker(const ClsArgs& a, int device, int first, int step, WorkerResult& res) {
    ScopedAffinity bind(device);
    ScopedLinkShared shared;
    CtxLease slot[2];
    bool busy[2] = {false, false};
    for (int s = 0; s < 2; ++s) {
        int rc = acquire_ctx(device, &slot[s].c);
        if (rc) return rc;
    }
    size_t max_src = 0, max_tight = 0;
    for (int i = first; i < a.n_images; i += step) {
        max_src = std::max(max_src, (size_t)wicca_pitch_bytes(a.Ws[i], 3) * a.Hs[i] + 256);
        if (!a.jpeg_lens && is_pinned_host(a.srcs[i])) max_tight = std::max(max_tight, (size_t)a.Ws[i] * 3 * a.Hs[i]);
    }
    // per image and target: n_depths icon slots followed by the source-image slot
    const int n_out = a.n_depths + (a.dst_images ? 1 : 0);
    std::vector<size_t> slot_elems(a.n_targets), out_off(a.n_targets + 1, 0);
    for (int t = 0; t < a.n_targets; ++t) {
        slot_elems[t] = (size_t)a.targets[t].out_h * a.targets[t].out_w * 3;
        out_off[t + 1] = out_off[t] + slot_elems[t] * n_out;
    }
    const size_t out_bytes = out_off[a.n_targets] * sizeof(float);
    for (int s = 0; s < 2; ++s) {
        Ctx& c = *slot[s].c;
        WICCA_CUDA(c.d_src.reserve(max_src));
        if (max_tight) WICCA_CUDA(c.d_stage.reserve(max_tight));
        WICCA_CUDA(c.d_f32a.reserve(out_bytes));
        WICCA_CUDA(c.h_out.reserve(out_bytes));
    }
    int j = 0;
    for (int i = first; i < a.n_images; i += step, ++j) {
        Ctx& c = *slot[j & 1].c;
        if (busy[j & 1]) {
            WICCA_CUDA(cudaStreamSynchronize(c.stream));
            c.flush_pending();
            add_times(res, c);
        }
        int H = a.Hs[i], W = a.Ws[i];
        int64_t pitch = wicca_pitch_bytes(W, 3);
        int rc;
        if (a.jpeg_lens) {
            float host_ms = 0;
            rc = jpeg_file_to_resident(c, a.srcs[i], a.jpeg_lens[i], &H, &W, &pitch, &host_ms);
        } else {
            const int64_t rowb = (int64_t)W * 3;
            const int64_t stride = (a.strides && a.strides[i]) ? a.strides[i] : rowb;
            WICCA_CUDA(cudaEventRecord(c.ev[0], c.stream));
            rc = upload_image_async(c, a.srcs[i], H, rowb, stride, pitch);
        }
        if (rc) { cudaStreamSynchronize(c.stream); c.pending.clear(); return rc; }
        WICCA_CUDA(cudaEventRecord(c.ev[1], c.stream));
        std::vector<IconOut> outs;
        rc = enqueue_icons_resident(c, H, W, 3, pitch, a.depths, a.n_depths, a.border_type, a.bconst, outs);
        if (rc) { cudaStreamSynchronize(c.stream); c.pending.clear(); return rc; }
        // tap tables per target for [icons..., source image]; one resize launch per target fills its slots
        std::vector<ResizeSrc> rs;
        for (int k = 0; k < a.n_depths; ++k) rs.push_back({outs[k].d_ptr, outs[k].h, outs[k].w, outs[k].pitch});
        if (a.dst_images) rs.push_back({(const uint8_t*)c.d_src.p, H, W, pitch});
        std::vector<ResizeTableBlob> blobs;
        std::vector<size_t> blob_off(a.n_targets + 1, 0);
        for (int t = 0; t < a.n_targets; ++t) {
            blobs.push_back(build_resize_tables(rs, a.targets[t].out_h, a.targets[t].out_w));
            blob_off[t + 1] = (size_t)align_up((int64_t)(blob_off[t] + blobs[t].bytes.size()), 256);
        }
        WICCA_CUDA(c.h_bounce.reserve(blob_off[a.n_targets]));
        WICCA_CUDA(c.d_misc.reserve(blob_off[a.n_targets]));
        for (int t = 0; t < a.n_targets; ++t) memcpy((uint8_t*)c.h_bounce.p + blob_off[t], blobs[t].bytes.data(), blobs[t].bytes.size());
        WICCA_CUDA(cudaMemcpyAsync(c.d_misc.p, c.h_bounce.p, blob_off[a.n_targets], cudaMemcpyHostToDevice, c.stream));
        for (int t = 0; t < a.n_targets; ++t) {
            cudaError_t e = launch_resize_norm(blobs[t].view((uint8_t*)c.d_misc.p + blob_off[t]), n_out, a.targets[t].out_h,
                                               a.targets[t].out_w, a.targets[t].norm_mode, (float*)c.d_f32a.p + out_off[t],
                                               nullptr, blobs[t].max_src_w, blobs[t].n_area, blobs[t].n_other, c.stream);
            if (e != cudaSuccess) { cudaStreamSynchronize(c.stream); return cuda_fail(e, "resize/normalise kernel"); }
        }
        WICCA_CUDA(cudaEventRecord(c.ev[2], c.stream));
        // D2H through the slot's pinned buffer (the destination batches are usually pageable NumPy memory)
        WICCA_CUDA(cudaMemcpyAsync(c.h_out.p, c.d_f32a.p, out_bytes, cudaMemcpyDeviceToHost, c.stream));
        for (int t = 0; t < a.n_targets; ++t) {
            const float* h = (const float*)c.h_out.p + out_off[t];
            const size_t bytes = slot_elems[t] * sizeof(float);
            for (int k = 0; k < a.n_depths; ++k)
                c.pending.push_back({a.dst_icons[(size_t)t * a.n_depths + k] + (size_t)i * slot_elems[t], h + (size_t)k * slot_elems[t], bytes});
            if (a.dst_images)
                c.pending.push_back({a.dst_images[t] + (size_t)i * slot_elems[t], h + (size_t)a.n_depths * slot_elems[t], bytes});
        }
        WICCA_CUDA(cudaEventRecord(c.ev[3], c.stream));
        busy[j & 1] = true;
    }
    for (int s = 0; s < 2; ++s)
        if (busy[s]) {
            WICCA_CUDA(cudaStreamSynchronize(slot[s].c->stream));
            slot[s].c->flush_pending();
            add_times(res, *slot[s].c);
        }
    return 0;
}

}  // namespace

namespace {

int run_classifier_batches(const uint8_t* const* srcs, const size_t* jpeg_lens, const int* Hs, const int* Ws, const int64_t* strides,
                           int n_images, const int* depths, int n_depths, int border_type, double border_const,
                           const wicca_target* targets, int n_targets, float* const* dst_icons, float* const* dst_images,
                           const int* devices, int n_devices, int workers_per_device, wicca_timing* t) {
    if (t) memset(t, 0, sizeof(*t));
    if (n_images < 0) return fail(WICCA_EINVAL, "negative image count");
    if (n_images == 0) return 0;
    if (!srcs || !depths || !targets || !dst_icons) return fail(WICCA_EINVAL, "null array");
    if (n_depths <= 0 || n_targets <= 0) return fail(WICCA_EINVAL, "need at least one depth and one target");
    for (int k = 0; k < n_depths; ++k)
        if (depths[k] < 1) return fail(WICCA_EDEPTH, "transform depth must be >= 1");
    for (int q = 0; q < n_targets; ++q) {
        if (targets[q].out_h <= 0 || targets[q].out_w <= 0 || targets[q].norm_mode < 0 || targets[q].norm_mode > 3)
            return fail(WICCA_EINVAL, "bad target size / mode (target %d)", q);
        for (int k = 0; k < n_depths; ++k)
            if (!dst_icons[(size_t)q * n_depths + k]) return fail(WICCA_EINVAL, "dst_icons[%d][%d] is NULL", q, k);
        if (dst_images && !dst_images[q]) return fail(WICCA_EINVAL, "dst_images[%d] is NULL", q);
    }
    if (n_devices <= 0) return fail(WICCA_EDEVICE, "need at least one device");
    for (int i = 0; i < n_images; ++i) {
        int rc = validate_icon_args(srcs[i], Hs[i], Ws[i], 3, depths, n_depths, border_type);
        if (rc) return rc;
        if (!jpeg_lens && strides && strides[i] && strides[i] < (int64_t)Ws[i] * 3) return fail(WICCA_EINVAL, "strides[%d] < W*3", i);
    }
    std::vector<int> devs;
    for (int k = 0; k < n_devices; ++k) {
        const int d = devices ? devices[k] : k;
        int rc = check_device(d);
        if (rc) return rc;
        for (int w = 0; w < workers_per_device; ++w) devs.push_back(d);
    }
    ClsArgs a{srcs, Hs, Ws, strides, n_images, jpeg_lens, depths, n_depths, border_type, saturate_u8(border_const), targets,
              n_targets, dst_icons, dst_images};
    const int nw = (int)devs.size() < n_images ? (int)devs.size() : n_images;
    std::vector<WorkerResult> results(nw);
    std::vector<std::thread> threads;
    for (int k = 0; k < nw; ++k)
        threads.emplace_back([&, k] {
            results[k].rc = cls_worker(a, devs[k], k, nw, results[k]);
            if (results[k].rc) results[k].msg = last_error_ref();
        });
    for (auto& th : threads) th.join();
    wicca_timing sum = {0, 0, 0, 0};
    for (auto& r : results) {
        if (r.rc) { last_error_ref() = r.msg; return r.rc; }
        sum.h2d_ms += (float)r.h2d; sum.kernel_ms += (float)r.kernel; sum.d2h_ms += (float)r.d2h; sum.total_ms += (float)r.total;
    }
    if (t) *t = sum;
    return 0;
}

}  // namespace

extern "C" int wicca_batch_classifier_inputs_multi_f32(const uint8_t* const* srcs, const int* Hs, const int* Ws,
                                                       const int64_t* strides, int n_images, const int* depths,
                                                       int n_depths, int border_type, double border_const,
                                                       const wicca_target* targets, int n_targets,
                                                       float* const* dst_icons, float* const* dst_images,
                                                       const int* devices, int n_devices, wicca_timing* t) {
    if (n_images > 0 && (!Hs || !Ws)) return fail(WICCA_EINVAL, "null array");
    return run_classifier_batches(srcs, nullptr, Hs, Ws, strides, n_images, depths, n_depths, border_type, border_const, targets,
                                  n_targets, dst_icons, dst_images, devices, n_devices, 1, t);
}

extern "C" int wicca_batch_classifier_inputs_multi_from_jpeg(const uint8_t* const* datas, const size_t* lens, int n_images,
                                                             const int* depths, int n_depths, int border_type,
                                                             double border_const, const wicca_target* targets, int n_targets,
                                                             float* const* dst_icons, float* const* dst_images,
                                                             const int* devices, int n_devices, wicca_timing* t) {
    if (n_images > 0 && (!datas || !lens)) return fail(WICCA_EINVAL, "null array");
    std::vector<int> Hs(n_images > 0 ? n_images : 0), Ws(Hs.size());
    for (int i = 0; i < n_images; ++i) {
        if (!datas[i]) return fail(WICCA_EINVAL, "datas[%d] is NULL", i);
        int rc = jpeg_output_dims(datas[i], lens[i], &Hs[i], &Ws[i]);
        if (rc) return rc;
    }
    // the host stage per file (marker parsing, byte unstuffing) is a few ms: four workers per GPU keep it fed
    return run_classifier_batches(datas, lens, Hs.data(), Ws.data(), nullptr, n_images, depths, n_depths, border_type, border_const,
                                  targets, n_targets, dst_icons, dst_images, devices, n_devices, 4, t);
}

extern "C" int wicca_batch_classifier_inputs_f32(const uint8_t* const* srcs, const int* Hs, const int* Ws,
                                                 const int64_t* strides, int n_images, int depth, int border_type,
                                                 double border_const, int out_h, int out_w, int norm_mode,
                                                 float* dst_icons, float* dst_images, const int* devices, int n_devices,
                                                 wicca_timing* t) {
    if (n_images > 0 && !dst_icons) return fail(WICCA_EINVAL, "null array");
    const wicca_target target = {out_h, out_w, norm_mode};
    float* icons[1] = {dst_icons};
    float* images[1] = {dst_images};
    return wicca_batch_classifier_inputs_multi_f32(srcs, Hs, Ws, strides, n_images, &depth, 1, border_type, border_const, &target,
                                                   1, icons, dst_images ? images : nullptr, devices, n_devices, t);
}
