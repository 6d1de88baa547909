// capi_resize.cu - C ABI of the resize + normalise epilogue (rows A5/A6), host-buffer variant.
#include <string.h>

#include <list>
#include <memory>
#include <mutex>
#include <vector>

#include "host_common.h"
#include "kernels.h"
#include "resize_tables.h"

using namespace wicca;

namespace {
// Device-resident tap tables of wicca_resize_norm_dev, most recently used first.  An evicted entry frees its tables with
// cudaFree, which waits for any kernel still reading them.
struct TableCacheEntry {
    std::vector<int64_t> key;
    int device = 0;
    ResizeTableBlob blob;          // offsets and launch hints (the bytes themselves are dropped after the upload)
    DevBuf d_tables;
    ~TableCacheEntry() {
        if (d_tables.p) { int cur = 0; cudaGetDevice(&cur); cudaSetDevice(device); d_tables.release(); cudaSetDevice(cur); }
    }
};
constexpr size_t kTableCacheEntries = 16;
std::mutex g_table_mu;
std::list<std::shared_ptr<TableCacheEntry>> g_table_cache;
}  // namespace

extern "C" int wicca_icon_resize_norm_f32(const uint8_t* const* icons, const int* hs, const int* ws, int n, int out_h,
                                          int out_w, int norm_mode, float* dst, uint8_t* dst_u8, int device,
                                          wicca_timing* t) {
    if (n < 0 || out_h <= 0 || out_w <= 0) return fail(WICCA_EINVAL, "bad batch/target size");
    if (norm_mode < 0 || norm_mode > 3) return fail(WICCA_EINVAL, "unknown normalisation mode %d", norm_mode);
    if (t) memset(t, 0, sizeof(*t));
    if (n == 0) return 0;
    if (!icons || !hs || !ws || !dst) return fail(WICCA_EINVAL, "null pointer");
    size_t src_bytes = 0;
    for (int i = 0; i < n; ++i) {
        if (!icons[i] || hs[i] <= 0 || ws[i] <= 0) return fail(WICCA_EINVAL, "icon %d is empty", i);
        src_bytes += (size_t)(align_up((int64_t)ws[i] * 3, 128) * hs[i]);
    }
    CtxLease L;
    int rc = acquire_ctx(device, &L.c);
    if (rc) return rc;
    Ctx& c = *L.c;
    WICCA_CUDA(c.d_src.reserve(src_bytes + 256));
    std::vector<ResizeSrc> srcs(n);
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t pitch = align_up((int64_t)ws[i] * 3, 128);      // 16-byte aligned rows for the row-streaming kernel
        srcs[i] = {(const uint8_t*)c.d_src.p + off, hs[i], ws[i], pitch};
        off += (size_t)(pitch * hs[i]);
    }
    const ResizeTableBlob blob = build_resize_tables(srcs, out_h, out_w);
    WICCA_CUDA(c.h_desc.reserve(blob.bytes.size()));
    WICCA_CUDA(c.d_desc.reserve(blob.bytes.size()));
    memcpy(c.h_desc.p, blob.bytes.data(), blob.bytes.size());

    const size_t out_elems = (size_t)n * out_h * out_w * 3;
    WICCA_CUDA(c.d_f32a.reserve(out_elems * sizeof(float)));
    if (dst_u8) WICCA_CUDA(c.d_misc.reserve(out_elems));

    WICCA_CUDA(cudaEventRecord(c.ev[0], c.stream));
    WICCA_CUDA(cudaMemcpyAsync(c.d_desc.p, c.h_desc.p, blob.bytes.size(), cudaMemcpyHostToDevice, c.stream));
    for (int i = 0; i < n; ++i)
        WICCA_CUDA(cudaMemcpy2DAsync((void*)srcs[i].d_ptr, (size_t)srcs[i].pitch, icons[i], (size_t)ws[i] * 3, (size_t)ws[i] * 3,
                                     (size_t)hs[i], cudaMemcpyHostToDevice, c.stream));
    WICCA_CUDA(cudaEventRecord(c.ev[1], c.stream));
    cudaError_t e = launch_resize_norm(blob.view(c.d_desc.p), n, out_h, out_w, norm_mode, (float*)c.d_f32a.p,
                                       dst_u8 ? (uint8_t*)c.d_misc.p : nullptr, blob.max_src_w, blob.n_area, blob.n_other,
                                       c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "resize/normalise kernel");
    WICCA_CUDA(cudaEventRecord(c.ev[2], c.stream));
    WICCA_CUDA(cudaMemcpyAsync(dst, c.d_f32a.p, out_elems * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
    if (dst_u8) WICCA_CUDA(cudaMemcpyAsync(dst_u8, c.d_misc.p, out_elems, cudaMemcpyDeviceToHost, c.stream));
    WICCA_CUDA(cudaEventRecord(c.ev[3], c.stream));
    WICCA_CUDA(cudaStreamSynchronize(c.stream));
    if (t) {
        cudaEventElapsedTime(&t->h2d_ms, c.ev[0], c.ev[1]);
        cudaEventElapsedTime(&t->kernel_ms, c.ev[1], c.ev[2]);
        cudaEventElapsedTime(&t->d2h_ms, c.ev[2], c.ev[3]);
        cudaEventElapsedTime(&t->total_ms, c.ev[0], c.ev[3]);
    }
    return 0;
}

// Device-resident variant: the sources already live in HBM (icons, or full source images for the
// reference's `cv2.resize(image, shape, interpolation)` branch, classifying_tools.py:315).
extern "C" int wicca_resize_norm_dev(const uint8_t* const* d_srcs, const int* hs, const int* ws, const int64_t* pitches,
                                     int n, int out_h, int out_w, int norm_mode, float* d_dst, uint8_t* d_dst_u8,
                                     int device, void* stream_v) {
    if (n < 0 || out_h <= 0 || out_w <= 0) return fail(WICCA_EINVAL, "bad batch/target size");
    if (norm_mode < 0 || norm_mode > 3) return fail(WICCA_EINVAL, "unknown normalisation mode %d", norm_mode);
    if (n == 0) return 0;
    if (!d_srcs || !hs || !ws || !pitches || !d_dst) return fail(WICCA_EINVAL, "null pointer");
    std::vector<ResizeSrc> srcs(n);
    for (int i = 0; i < n; ++i) {
        if (!d_srcs[i] || hs[i] <= 0 || ws[i] <= 0 || pitches[i] < (int64_t)ws[i] * 3) return fail(WICCA_EINVAL, "image %d is malformed", i);
        srcs[i] = {d_srcs[i], hs[i], ws[i], pitches[i]};
    }
    int rc = check_device(device);
    if (rc) return rc;
    WICCA_CUDA(cudaSetDevice(device));
    cudaStream_t stream = (cudaStream_t)stream_v;
    // The tap tables depend only on (sources, target): a batch that stays resident and is resized again - the reference
    // does it once per classifier and depth, classifying_tools.py:546-551 - finds them on the device and the call is the
    // kernel launch alone.  A miss builds them on the host and uploads them synchronously (they are read by kernels on
    // whatever stream later calls pass, so the upload must not be ordered on this call's stream only).
    std::vector<int64_t> key;
    key.reserve(4 + 4 * (size_t)n);
    key.push_back(device); key.push_back(out_h); key.push_back(out_w); key.push_back(n);
    for (int i = 0; i < n; ++i) { key.push_back((int64_t)(uintptr_t)d_srcs[i]); key.push_back(hs[i]); key.push_back(ws[i]); key.push_back(pitches[i]); }
    std::shared_ptr<TableCacheEntry> entry;
    {
        std::lock_guard<std::mutex> lk(g_table_mu);
        for (auto it = g_table_cache.begin(); it != g_table_cache.end(); ++it)
            if ((*it)->key == key) { entry = *it; g_table_cache.erase(it); g_table_cache.push_front(entry); break; }
    }
    if (!entry) {
        entry = std::make_shared<TableCacheEntry>();
        entry->key = key;
        entry->device = device;
        entry->blob = build_resize_tables(srcs, out_h, out_w);
        cudaError_t e = entry->d_tables.reserve(entry->blob.bytes.size());
        if (e == cudaSuccess) e = cudaMemcpy(entry->d_tables.p, entry->blob.bytes.data(), entry->blob.bytes.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { entry->d_tables.release(); return cuda_fail(e, "resize table upload"); }
        entry->blob.bytes.clear();
        entry->blob.bytes.shrink_to_fit();
        std::lock_guard<std::mutex> lk(g_table_mu);
        g_table_cache.push_front(entry);
        while (g_table_cache.size() > kTableCacheEntries) g_table_cache.pop_back();     // ~TableCacheEntry frees the device tables
    }
    cudaError_t e = launch_resize_norm(entry->blob.view(entry->d_tables.p), n, out_h, out_w, norm_mode, d_dst, d_dst_u8,
                                       entry->blob.max_src_w, entry->blob.n_area, entry->blob.n_other, stream);
    if (e != cudaSuccess) return cuda_fail(e, "resize/normalise kernel");
    return 0;
}

namespace wicca {
void resize_table_cache_clear() {
    std::lock_guard<std::mutex> lk(g_table_mu);
    g_table_cache.clear();
}
}  // namespace wicca
