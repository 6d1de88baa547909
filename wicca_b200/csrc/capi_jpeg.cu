// capi_jpeg.cu - C entry points of the JPEG ingest path (row N2).  Replaces `cv2.imread(path)` +
// `cv2.cvtColor(image, cv2.COLOR_BGR2RGB)` (wicca/data_loader.py:53-58) for baseline JPEG files: the host parses
// the markers and strips the byte stuffing; Huffman decoding and everything after it run on the GPU, and the RGB
// image lands directly in the pitched device buffer the icon kernel reads.
#include <string.h>

#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "host_common.h"
#include "jpeg_gpu.h"
#include "jpeg_host.h"
#include "kernels.h"

using namespace wicca;

namespace {

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int parse_or_fail(const uint8_t* data, size_t len, JpegFrame& f) {
    std::string why;
    const int rc = jpeg_parse(data, len, f, why);
    if (rc) return fail(rc, "%s", why.c_str());
    return 0;
}

int upsample_mode(const JpegFrame& f, const JpegComponent& q) {
    const int hf = f.hmax / q.h, vf = f.vmax / q.v;
    if (hf == 1 && vf == 1) return 0;
    if (hf == 2 && vf == 1) return q.dw > 2 ? 1 : 4;       // jdsample.c: fancy only when downsampled_width > 2
    if (hf == 2 && vf == 2) return q.dw > 2 ? 2 : 4;
    if (hf == 1 && vf == 2) return 3;
    return 4;
}

// Huffman stage on the GPU: the host only strips the byte stuffing; coefficients end up in c.d_f32a.
// Returns 1 when the scan cannot be decoded here (damaged restart structure, no fixed point within the pass budget):
// the caller reports the file as WICCA_EUNSUPPORTED - there is no host decoder behind this one.
int jpeg_gpu_huffman_stage(Ctx& c, const uint8_t* data, size_t len, const JpegFrame& f, cudaStream_t stream, float* host_ms,
                           int* passes_out) {
    const size_t cap = len - f.scan_offset + 32;
    WICCA_CUDA(c.h_in.reserve(cap));
    const double t0 = now_ms();
    std::vector<uint32_t> starts;
    const size_t n_bytes = jpeg_unstuff_scan(data, len, f, (uint8_t*)c.h_in.p, &starts);
    if (host_ms) *host_ms += (float)(now_ms() - t0);
    if (n_bytes == 0 || n_bytes >= ((size_t)1 << 28)) return 1;
    if (f.restart_interval) {
        // exactly one marker between consecutive intervals, or the file is damaged
        const int64_t mcus = (int64_t)f.mcux * f.mcuy;
        if ((int64_t)starts.size() != (mcus + f.restart_interval - 1) / f.restart_interval - 1) return 1;
    } else if (!starts.empty()) {
        return 1;
    }
    WICCA_CUDA(cudaEventRecord(c.ev[0], stream));                 // device work starts here
    JpegGpuScan sc;
    memset(&sc, 0, sizeof sc);
    sc.total_bits = (uint32_t)(n_bytes * 8);
    sc.n_sub = (sc.total_bits + kSubBits - 1) / kSubBits;
    sc.mcux = f.mcux; sc.ncomp = f.ncomp; sc.total_coefs = f.total_coefs;
    int slot = 0;
    for (int k = 0; k < f.ncomp; ++k) {
        const JpegComponent& q = f.comp[k];
        sc.comp[k] = {q.coef_offset, (int64_t)q.blocks_w * q.blocks_h, q.h, q.v, q.blocks_w};
        for (int by = 0; by < q.v; ++by)
            for (int bx = 0; bx < q.h; ++bx, ++slot) sc.slot[slot] = {q.coef_offset, q.h, q.v, bx, by, q.blocks_w, q.td, q.ta};
    }
    sc.blocks_per_mcu = slot;
    sc.total_blocks = (int64_t)f.mcux * f.mcuy * slot;
    // host copy of the tables + the flag the fixed-point loop polls, in page-locked memory
    // (a buffer of its own: nothing else may write to it while the upload below is in flight)
    WICCA_CUDA(c.h_jpeg.reserve(sizeof(JpegGpuTables) + 64 + (starts.size() + 1) * sizeof(uint32_t)));
    JpegGpuTables* ht = (JpegGpuTables*)c.h_jpeg.p;
    memset(ht, 0, sizeof *ht);
    for (int id = 0; id < 4; ++id) {
        const JpegHuff* src[2] = {&f.dc[id], &f.ac[id]};
        for (int kind = 0; kind < 2; ++kind) {
            if (!src[kind]->present) continue;
            const int t = 4 * kind + id;
            memcpy(ht->look[t], src[kind]->look, sizeof ht->look[t]);
            memcpy(ht->valoffset[t], src[kind]->valoffset, sizeof ht->valoffset[t]);
            memcpy(ht->symbols[t], src[kind]->symbols, sizeof ht->symbols[t]);
            // canonical codes are assigned in increasing order: every code of <= l bits, left-aligned to 16 bits,
            // is below (largest code of the longest length <= l, + 1) << (16 - that length)
            int32_t lim = 0;
            for (int l = 1; l <= 16; ++l) {
                if (src[kind]->maxcode[l] >= 0) lim = (src[kind]->maxcode[l] + 1) << (16 - l);
                if (l >= 10) ht->limit[t][l - 10] = lim;
            }
        }
    }
    int* h_changed = (int*)((uint8_t*)c.h_jpeg.p + sizeof(JpegGpuTables));
    // device scratch: [scan words][tables][exit a][exit b][start used][count][base][changed][chunk sums]
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (size_t)align_up((int64_t)(off + bytes), 256); return o; };
    const size_t o_words = take(n_bytes + 16), o_tab = take(sizeof(JpegGpuTables));
    const size_t o_ea = take((size_t)sc.n_sub * 8), o_eb = take((size_t)sc.n_sub * 8), o_su = take((size_t)sc.n_sub * 8);
    const size_t o_cnt = take((size_t)sc.n_sub * 4), o_base = take((size_t)sc.n_sub * 4), o_chg = take(4);
    const size_t o_sums = take(jpeg_gpu_chunk_sum_capacity(sc));
    int64_t max_blocks = 0;
    for (int k = 0; k < f.ncomp; ++k) max_blocks = std::max(max_blocks, sc.comp[k].n_blocks);
    const size_t o_bounds = take((starts.size() + 1) * sizeof(uint32_t));
    const size_t o_prefix = take(starts.empty() ? 4 : (size_t)max_blocks * sizeof(int32_t));
    WICCA_CUDA(c.d_misc.reserve(off));
    WICCA_CUDA(c.d_f32a.reserve((size_t)f.total_coefs * sizeof(int16_t)));
    uint8_t* m = (uint8_t*)c.d_misc.p;
    sc.words = (const uint32_t*)(m + o_words);
    sc.tables = (const JpegGpuTables*)(m + o_tab);
    sc.start_used = (uint64_t*)(m + o_su);
    sc.count = (uint32_t*)(m + o_cnt);
    sc.base = (uint32_t*)(m + o_base);
    sc.changed = (int*)(m + o_chg);
    sc.coefs = (int16_t*)c.d_f32a.p;
    sc.bounds = (const uint32_t*)(m + o_bounds);
    sc.n_bounds = (uint32_t)starts.size();
    sc.blocks_per_interval = (int64_t)f.restart_interval * sc.blocks_per_mcu;
    sc.dc_prefix = (int32_t*)(m + o_prefix);
    {
        // interval boundaries as bit offsets + sentinel, staged behind the tables in the page-locked buffer
        uint32_t* hb = (uint32_t*)((uint8_t*)c.h_jpeg.p + sizeof(JpegGpuTables) + 64);
        for (size_t k = 0; k < starts.size(); ++k) hb[k] = starts[k] * 8u;
        hb[starts.size()] = 0xFFFFFFFFu;
        WICCA_CUDA(cudaMemcpyAsync(m + o_bounds, hb, (starts.size() + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    }
    WICCA_CUDA(cudaMemcpyAsync(m + o_words, c.h_in.p, n_bytes + 16, cudaMemcpyHostToDevice, stream));
    WICCA_CUDA(cudaMemcpyAsync(m + o_tab, ht, sizeof(JpegGpuTables), cudaMemcpyHostToDevice, stream));
    WICCA_CUDA(cudaEventRecord(c.ev[4], stream));                 // scan resident
    cudaError_t e = launch_jpeg_huffman(sc, (uint64_t*)(m + o_ea), (uint64_t*)(m + o_eb), (int64_t*)(m + o_sums), h_changed, 1024,
                                        passes_out, stream);
    if (e == cudaErrorNotReady) return 1;
    if (e != cudaSuccess) return cuda_fail(e, "JPEG Huffman kernels");
    return 0;
}

// Device stage: coefficients (in c.d_f32a, left there by the Huffman stage) -> d_dst (RGB, rows d_pitch bytes apart)
// on `stream`, through c's scratch.
int jpeg_device_stage(Ctx& c, const JpegFrame& f, uint8_t* d_dst, int64_t d_pitch, cudaStream_t stream) {
    const size_t coef_bytes = (size_t)f.total_coefs * sizeof(int16_t);
    size_t plane_bytes = 0;
    JpegImageDesc d;
    memset(&d, 0, sizeof d);
    d.ncomp = f.ncomp; d.width = f.width; d.height = f.height; d.dst = d_dst; d.dst_pitch = d_pitch;
    if (f.orientation != 1) {                                    // decode upright into scratch, then flip / transpose
        const int64_t tmp_pitch = wicca_pitch_bytes(f.width, 3);
        WICCA_CUDA(c.d_tmp.reserve((size_t)tmp_pitch * f.height + 256));
        d.dst = (uint8_t*)c.d_tmp.p; d.dst_pitch = tmp_pitch;
    }
    size_t plane_off[3];
    for (int k = 0; k < f.ncomp; ++k) {
        const JpegComponent& q = f.comp[k];
        plane_off[k] = plane_bytes;
        plane_bytes += (size_t)q.blocks_w * 8 * q.blocks_h * 8;
    }
    WICCA_CUDA(c.d_f32a.reserve(coef_bytes));
    WICCA_CUDA(c.d_f32b.reserve(plane_bytes + 256));
    for (int k = 0; k < f.ncomp; ++k) {
        const JpegComponent& q = f.comp[k];
        JpegPlaneDesc& p = d.comp[k];
        p.coefs = (const int16_t*)c.d_f32a.p + q.coef_offset;
        p.plane = (uint8_t*)c.d_f32b.p + plane_off[k];
        p.blocks_w = q.blocks_w; p.blocks_h = q.blocks_h; p.plane_pitch = q.blocks_w * 8;
        p.dw = q.dw; p.dh = q.dh; p.hf = f.hmax / q.h; p.vf = f.vmax / q.v;
        p.mode = upsample_mode(f, q);
        memcpy(p.qt, f.qt[q.tq], sizeof p.qt);
    }
    cudaError_t e = launch_jpeg_decode(d, stream);
    if (e == cudaSuccess && f.orientation != 1)
        e = launch_jpeg_orient(d.dst, d.dst_pitch, f.height, f.width, f.orientation, d_dst, d_pitch, stream);
    if (e != cudaSuccess) return cuda_fail(e, "JPEG decode kernels");
    return 0;
}

// Both stages: JPEG bytes -> RGB in d_dst.  ev[0] is recorded when the device work starts, ev[4] when its input
// (scan bytes or coefficients) is resident.
int jpeg_to_device(Ctx& c, const uint8_t* data, size_t len, const JpegFrame& f, uint8_t* d_dst, int64_t d_pitch,
                   cudaStream_t stream, float* host_ms) {
    if (f.multiscan)
        return fail(WICCA_EUNSUPPORTED, "progressive / multi-scan JPEG: its scans depend on each other and are not decoded on "
                                        "the GPU (read this file with cv2.imread)");
    int rc = jpeg_gpu_huffman_stage(c, data, len, f, stream, host_ms, nullptr);
    if (rc == 1) {
        cudaStreamSynchronize(stream);
        return fail(WICCA_EUNSUPPORTED, "JPEG scan with a damaged restart structure / no decoder fixed point (read this file with cv2.imread)");
    }
    if (rc) return rc;
    return jpeg_device_stage(c, f, d_dst, d_pitch, stream);
}

}  // namespace

namespace wicca {

int jpeg_output_dims(const uint8_t* data, size_t len, int* H, int* W) {
    JpegFrame f;
    int rc = parse_or_fail(data, len, f);
    if (rc) return rc;
    jpeg_output_size(f, H, W);
    return 0;
}

int jpeg_file_to_resident(Ctx& c, const uint8_t* data, size_t len, int* H, int* W, int64_t* pitch, float* host_ms) {
    JpegFrame f;
    int rc = parse_or_fail(data, len, f);
    if (rc) return rc;
    jpeg_output_size(f, H, W);
    *pitch = wicca_pitch_bytes(*W, 3);
    WICCA_CUDA(c.d_src.reserve((size_t)*pitch * *H + 256));
    return jpeg_to_device(c, data, len, f, (uint8_t*)c.d_src.p, *pitch, c.stream, host_ms);
}

}  // namespace wicca

extern "C" {

int wicca_jpeg_probe(const uint8_t* data, size_t len, int* H, int* W, int* n_components, int* h_max, int* v_max) {
    JpegFrame f;
    int rc = parse_or_fail(data, len, f);
    if (rc) return rc;
    if (f.multiscan) return fail(WICCA_EUNSUPPORTED, "progressive / multi-scan JPEG (not decoded on the GPU; read it with cv2.imread)");
    int oh, ow;
    jpeg_output_size(f, &oh, &ow);
    if (H) *H = oh;
    if (W) *W = ow;
    if (n_components) *n_components = f.ncomp;
    if (h_max) *h_max = f.hmax;
    if (v_max) *v_max = f.vmax;
    return 0;
}

int64_t wicca_jpeg_coeff_count(const uint8_t* data, size_t len) {
    JpegFrame f;
    int rc = parse_or_fail(data, len, f);
    return rc ? (int64_t)rc : f.total_coefs;
}

int wicca_jpeg_decode_coeffs_gpu(const uint8_t* data, size_t len, int16_t* dst, int64_t dst_count, int device, int* passes) {
    JpegFrame f;
    int rc = parse_or_fail(data, len, f);
    if (rc) return rc;
    if (!dst || dst_count < f.total_coefs) return fail(WICCA_EINVAL, "coefficient buffer too small (%lld needed)", (long long)f.total_coefs);
    if (f.multiscan) return fail(WICCA_EUNSUPPORTED, "progressive / multi-scan files are not decoded on the GPU");
    rc = check_device(device);
    if (rc) return rc;
    CtxLease lease;
    rc = acquire_ctx(device, &lease.c);
    if (rc) return rc;
    Ctx& c = *lease.c;
    rc = jpeg_gpu_huffman_stage(c, data, len, f, c.stream, nullptr, passes);
    if (rc == 1) { cudaStreamSynchronize(c.stream); return fail(WICCA_ESTATE, "GPU Huffman decoder found no fixed point within its pass budget"); }
    if (rc) { cudaStreamSynchronize(c.stream); return rc; }
    WICCA_CUDA(cudaMemcpyAsync(dst, c.d_f32a.p, (size_t)f.total_coefs * sizeof(int16_t), cudaMemcpyDeviceToHost, c.stream));
    WICCA_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

int wicca_jpeg_decode_u8(const uint8_t* data, size_t len, uint8_t* dst, int64_t dst_stride, int device, wicca_timing* t,
                         float* host_decode_ms) {
    if (t) memset(t, 0, sizeof(*t));
    if (host_decode_ms) *host_decode_ms = 0;
    JpegFrame f;
    int rc = parse_or_fail(data, len, f);
    if (rc) return rc;
    int oh, ow;
    jpeg_output_size(f, &oh, &ow);
    const int64_t rowb = (int64_t)ow * 3;
    if (!dst) return fail(WICCA_EINVAL, "dst is NULL");
    if (dst_stride == 0) dst_stride = rowb;
    if (dst_stride < rowb) return fail(WICCA_EINVAL, "dst_stride < W*3");
    rc = check_device(device);
    if (rc) return rc;
    CtxLease lease;
    rc = acquire_ctx(device, &lease.c);
    if (rc) return rc;
    Ctx& c = *lease.c;
    const int64_t pitch = wicca_pitch_bytes(ow, 3);
    WICCA_CUDA(c.d_src.reserve((size_t)pitch * oh + 256));
    float host_ms = 0;
    rc = jpeg_to_device(c, data, len, f, (uint8_t*)c.d_src.p, pitch, c.stream, &host_ms);
    if (rc) { cudaStreamSynchronize(c.stream); return rc; }
    WICCA_CUDA(cudaEventRecord(c.ev[2], c.stream));
    uint8_t* target = dst;
    int64_t target_stride = dst_stride;
    const bool direct = is_pinned_host(dst);
    if (!direct) {
        WICCA_CUDA(c.h_bounce.reserve((size_t)rowb * oh));
        target = (uint8_t*)c.h_bounce.p;
        target_stride = rowb;
    }
    WICCA_CUDA(cudaMemcpy2DAsync(target, (size_t)target_stride, c.d_src.p, (size_t)pitch, (size_t)rowb, (size_t)oh,
                                 cudaMemcpyDeviceToHost, c.stream));
    WICCA_CUDA(cudaEventRecord(c.ev[3], c.stream));
    WICCA_CUDA(cudaStreamSynchronize(c.stream));
    if (!direct) {
        if (dst_stride == rowb) {
            c.pending.push_back({dst, target, (size_t)rowb * oh});
            c.flush_pending();
        } else {
            for (int y = 0; y < oh; ++y) memcpy(dst + (size_t)y * dst_stride, target + (size_t)y * rowb, (size_t)rowb);
        }
    }
    if (t) {
        cudaEventElapsedTime(&t->h2d_ms, c.ev[0], c.ev[4]);         // coefficient upload
        cudaEventElapsedTime(&t->kernel_ms, c.ev[4], c.ev[2]);      // IDCT + upsampling/colour kernels
        cudaEventElapsedTime(&t->d2h_ms, c.ev[2], c.ev[3]);
        cudaEventElapsedTime(&t->total_ms, c.ev[0], c.ev[3]);
    }
    if (host_decode_ms) *host_decode_ms = host_ms;
    return 0;
}

int wicca_jpeg_decode_dev(const uint8_t* data, size_t len, uint8_t* d_dst, int64_t d_pitch, int device, void* stream_v) {
    JpegFrame f;
    int rc = parse_or_fail(data, len, f);
    if (rc) return rc;
    if (!d_dst) return fail(WICCA_EINVAL, "d_dst is NULL");
    int oh, ow;
    jpeg_output_size(f, &oh, &ow);
    if (d_pitch < (int64_t)ow * 3) return fail(WICCA_EINVAL, "d_pitch < W*3");
    rc = check_device(device);
    if (rc) return rc;
    CtxLease lease;
    rc = acquire_ctx(device, &lease.c);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_v;
    rc = jpeg_to_device(*lease.c, data, len, f, d_dst, d_pitch, stream, nullptr);
    // the staging buffers belong to the leased context: they may be reused as soon as this returns
    cudaError_t e = cudaStreamSynchronize(stream);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
    return 0;
}

int wicca_jpeg_icons_multi_u8(const uint8_t* data, size_t len, const int* depths, int n_depths, int border_type,
                              double border_const, uint8_t* const* dsts, int device, wicca_timing* t, float* host_decode_ms) {
    if (t) memset(t, 0, sizeof(*t));
    if (host_decode_ms) *host_decode_ms = 0;
    JpegFrame f;
    int rc = parse_or_fail(data, len, f);
    if (rc) return rc;
    if (!dsts) return fail(WICCA_EINVAL, "dsts is NULL");
    uint8_t probe = 0;                                     // validate_icon_args only checks the pointer for NULL
    int oh, ow;
    jpeg_output_size(f, &oh, &ow);
    rc = validate_icon_args(&probe, oh, ow, 3, depths, n_depths, border_type);
    if (rc) return rc;
    for (int k = 0; k < n_depths; ++k) {
        if (!dsts[k]) return fail(WICCA_EINVAL, "dsts[%d] is NULL", k);
        if (depths[k] <= 0) return fail(WICCA_EDEPTH, "depths must be >= 1 here (depth 0 is the decoded image: use wicca_jpeg_decode_u8)");
    }
    rc = check_device(device);
    if (rc) return rc;
    CtxLease lease;
    rc = acquire_ctx(device, &lease.c);
    if (rc) return rc;
    Ctx& c = *lease.c;
    const int64_t pitch = wicca_pitch_bytes(ow, 3);
    WICCA_CUDA(c.d_src.reserve((size_t)pitch * oh + 256));
    float host_ms = 0;
    rc = jpeg_to_device(c, data, len, f, (uint8_t*)c.d_src.p, pitch, c.stream, &host_ms);
    if (rc) { cudaStreamSynchronize(c.stream); return rc; }
    WICCA_CUDA(cudaEventRecord(c.ev[1], c.stream));
    rc = icons_from_resident(c, oh, ow, 3, pitch, depths, n_depths, border_type, saturate_u8(border_const), dsts);
    if (rc) { cudaStreamSynchronize(c.stream); c.pending.clear(); return rc; }
    WICCA_CUDA(cudaStreamSynchronize(c.stream));
    c.flush_pending();
    if (t) {
        cudaEventElapsedTime(&t->h2d_ms, c.ev[0], c.ev[4]);        // coefficient upload
        cudaEventElapsedTime(&t->kernel_ms, c.ev[4], c.ev[2]);     // decode kernels + icon kernel
        cudaEventElapsedTime(&t->d2h_ms, c.ev[2], c.ev[3]);
        cudaEventElapsedTime(&t->total_ms, c.ev[0], c.ev[3]);
    }
    if (host_decode_ms) *host_decode_ms = host_ms;
    return 0;
}

int wicca_batch_icons_from_jpeg(const uint8_t* const* datas, const size_t* lens, int n_images, const int* depths, int n_depths,
                                int border_type, double border_const, uint8_t* const* dsts, const int* devices, int n_devices,
                                int n_threads, float* host_decode_ms) {
    if (host_decode_ms) *host_decode_ms = 0;
    if (n_images < 0) return fail(WICCA_EINVAL, "negative image count");
    if (n_images == 0) return 0;
    if (!datas || !lens || !dsts || !depths) return fail(WICCA_EINVAL, "null array");
    if (n_devices <= 0) return fail(WICCA_EDEVICE, "need at least one device");
    std::vector<int> devs(n_devices);
    for (int k = 0; k < n_devices; ++k) {
        devs[k] = devices ? devices[k] : k;
        int rc = check_device(devs[k]);
        if (rc) return rc;
    }
    if (n_threads <= 0) {
        n_threads = (int)std::thread::hardware_concurrency();
        if (n_threads > 32) n_threads = 32;
        if (n_threads < 1) n_threads = 1;
    }
    if (n_threads > n_images) n_threads = n_images;
    // Huffman decoding is serial per file and ~50x slower than everything the GPU does with the result, so the
    // unit of parallelism is the file: each worker thread takes the next file, decodes it on its own context
    // (own stream and buffers) on device i % n_devices, and the device work of different files overlaps.
    std::atomic<int> next(0);
    std::vector<int> rcs(n_threads, 0);
    std::vector<std::string> msgs(n_threads);
    std::vector<float> host_ms(n_threads, 0.f);
    auto work = [&](int tix) {
        ScopedLinkShared shared;                  // the icons of many files flow back at once: flat copies over the link
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n_images) return;
            float hm = 0;
            int rc = wicca_jpeg_icons_multi_u8(datas[i], lens[i], depths, n_depths, border_type, border_const,
                                               dsts + (size_t)i * n_depths, devs[i % n_devices], nullptr, &hm);
            host_ms[tix] += hm;
            if (rc) { rcs[tix] = rc; msgs[tix] = "image " + std::to_string(i) + ": " + last_error_ref(); next.store(n_images); return; }
        }
    };
    std::vector<std::thread> threads;
    for (int k = 1; k < n_threads; ++k) threads.emplace_back(work, k);
    work(0);
    for (auto& th : threads) th.join();
    float sum = 0;
    for (int k = 0; k < n_threads; ++k) {
        if (rcs[k]) { last_error_ref() = msgs[k]; return rcs[k]; }
        sum += host_ms[k];
    }
    if (host_decode_ms) *host_decode_ms = sum;
    return 0;
}

}  // extern "C"
