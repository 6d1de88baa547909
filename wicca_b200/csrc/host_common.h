// host_common.h - host-side plumbing shared by the C-ABI translation units (internal).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/wicca_b200.h"

namespace wicca {

// ---- thread-local error string --------------------------------------------------------
std::string& last_error_ref();
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define WICCA_CUDA(call)                                          \
    do {                                                          \
        cudaError_t _e = (call);                                  \
        if (_e != cudaSuccess) return ::wicca::cuda_fail(_e, #call); \
    } while (0)

// ---- small helpers ---------------------------------------------------------------------
inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int icon_dim(int n, int depth) { return depth <= 0 ? n : (int)(((int64_t)n + ((int64_t)1 << depth) - 1) >> depth); }
int saturate_u8(double v);                 // cv::saturate_cast<uchar>(double)
bool border_valid(int border_type);        // after masking BORDER_ISOLATED
inline int border_base(int border_type) { return border_type & ~16; }

// ---- grow-only buffers ------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes);     // keeps contents only if no growth is needed
    void release();
};
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes);
    void release();
};

// ---- per-call context: one stream + scratch, leased from a per-device pool --------------
struct Ctx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {};                // start, h2d done, kernels done, d2h done, spare x2
    DevBuf d_src, d_icons, d_desc, d_strip, d_f32a, d_f32b, d_misc, d_tmp, d_sum6;
    DevBuf d_stage, d_pack;                // tight-row staging of an upload / of results on their way to the host (see below)
    bool link_shared = false;              // set by the batch workers: uploads and read-backs of other images share the link
    PinBuf h_desc, h_bounce, h_in, h_out, h_jpeg;
    // Results bound for pageable host memory are DMA'd into h_bounce (so the copy is truly
    // asynchronous) and moved to their destination after the stream has been synchronised.
    struct Pending { void* dst; const void* src; size_t bytes; };
    std::vector<Pending> pending;
    void flush_pending();                  // call after cudaStreamSynchronize(stream)
    cudaError_t init(int dev);
    void destroy();
};

struct DeviceInfo {
    bool ok = false;
    int sm_count = 0;
    size_t smem_optin = 0;
};

int device_count_cached();
int check_device(int device);                       // 0 or WICCA_EDEVICE / cuda error
const DeviceInfo& device_info(int device);
int acquire_ctx(int device, Ctx** out);             // sets the calling thread's current device
// Batch workers declare that the host link is shared with other images' traffic for as long as the guard lives: every
// context the thread acquires meanwhile has link_shared set (upload_image_async / download_rows_async read it).
struct ScopedLinkShared {
    ScopedLinkShared();
    ~ScopedLinkShared();
    bool prev;
};
void release_ctx(Ctx* c);
void destroy_all_ctx();
void resize_table_cache_clear();                        // capi_resize.cu: cached tap tables of wicca_resize_norm_dev

struct CtxLease {
    Ctx* c = nullptr;
    ~CtxLease() { if (c) release_ctx(c); }
};

// ---- TMA descriptor ---------------------------------------------------------------------
// 2-D uint32 view (pitch/4 x H) of a pitched byte image, box (kStageRowBytes/4) x kItemH.
int encode_image_tmap(CUtensorMap* tm, const void* d_src, int H, int64_t pitch);

int encode_icon_tmap(CUtensorMap* tm, const void* d_icon, int h, int64_t w_bytes, int64_t pitch, int box_w, int box_h);

bool is_pinned_host(const void* p);
int host_copy_threads();          // threads for large host-side copies (WICCA_UPLOAD_THREADS overrides)

// NUMA locality: CPUs attached to the same socket / PCIe root as `device` (from sysfs; empty when unknown).
// ScopedAffinity binds the calling thread to them for its lifetime, so that page-locked memory
// allocated meanwhile lands on the GPU's own NUMA node and staging copies stay local.
struct ScopedAffinity {
    explicit ScopedAffinity(int device);
    ~ScopedAffinity();
    bool active = false;
    unsigned long saved[16] = {};
};

// Host image (rows `stride` bytes apart) -> c.d_src (rows `pitch` bytes apart), asynchronously on c.stream.
// Page-locked sources are DMA'd directly; pageable ones are staged band by band through the pinned
// buffer c.h_in by a few helper threads, so the CPU copy of band k+1 overlaps the DMA of band k.
// c.d_src must already be reserved.
// In a batch worker (c.link_shared) everything large crosses the link as ONE FLAT copy: a pitched 2-D copy costs nothing
// on an idle link (55.0 vs 55.4 GB/s) but 6 % of the end-to-end rate once uploads and read-backs share it (measured:
// 16.4-16.5 k -> 17.5 k MP/s, 0.92 -> 0.98 of the plain-copy ceiling).  So a tight page-locked image lands in c.d_stage and
// copy_rows_kernel (0.1 ms for 159 MB, off the link) spreads it to the pitch; results take the mirror path through
// c.d_pack.  One-shot calls have the link to themselves and keep the direct 2-D copies (0.1 ms less latency).
int upload_image_async(Ctx& c, const uint8_t* src, int H, int64_t row_bytes, int64_t stride, int64_t pitch);
// Pitched device rows -> tight host rows on c.stream (page-locked or bounce target): flat over the link when the block
// is large, through c.d_pack at offset *pack_off (advanced); the caller reserves c.d_pack for everything it sends.
int download_rows_async(Ctx& c, void* h_dst, const void* d_src, int64_t d_pitch, int64_t row_bytes, int rows, size_t* pack_off);
constexpr size_t kFlatCopyMin = (size_t)256 << 10;      // below this a 2-D copy straight over the link is cheaper

int icon_variant_from_env();

// JPEG bytes -> c.d_src (RGB, oriented, rows *pitch bytes apart) on c.stream; ev[0] marks the start of the device
// work.  capi_jpeg.cu.
int jpeg_file_to_resident(Ctx& c, const uint8_t* data, size_t len, int* H, int* W, int64_t* pitch, float* host_ms);
int jpeg_output_dims(const uint8_t* data, size_t len, int* H, int* W);

struct IconOut {          // one requested depth of one image
    int depth = 0;
    int h = 0, w = 0;
    uint8_t* d_ptr = nullptr;   // device icon (rows `pitch` bytes apart)
    int64_t pitch = 0;
};

// capi_icon.cu: image already resident in c.d_src (pitched).  enqueue_icons_resident launches the
// kernels for every depth > 0 on c.stream (icons stay in c.d_icons, described by `outs`) and records
// c.ev[2]; icons_from_resident additionally enqueues their D2H copies and records c.ev[3].
int enqueue_icons_resident(Ctx& c, int H, int W, int C, int64_t pitch, const int* depths, int n_depths, int border_type,
                           int bconst, std::vector<IconOut>& outs);
int icons_from_resident(Ctx& c, int H, int W, int C, int64_t pitch, const int* depths, int n_depths, int border_type,
                        int bconst, uint8_t* const* dsts);
int validate_icon_args(const void* src, int H, int W, int C, const int* depths, int n_depths, int border_type);

}  // namespace wicca
