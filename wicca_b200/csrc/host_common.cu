// host_common.cu - error strings, device/context pool, TMA descriptor encoding.
#include "host_common.h"
#include "kernels.h"

#include <math.h>
#include <ctype.h>
#include <stdlib.h>
#include <string.h>

#include "icon_types.h"

#include <sched.h>
#include <thread>

namespace wicca {

std::string& last_error_ref() {
    static thread_local std::string s;
    return s;
}

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s): %s [%s]", (int)e, cudaGetErrorName(e), cudaGetErrorString(e), what);
    last_error_ref() = buf;
    cudaGetLastError();   // clear the sticky-less error state
    return (int)e;
}

int saturate_u8(double v) {
    if (!(v == v)) return 0;
    double r = nearbyint(v);   // default rounding mode: half to even, like cvRound
    if (r < 0) return 0;
    if (r > 255) return 255;
    return (int)r;
}

bool border_valid(int border_type) {
    const int b = border_base(border_type);
    return b >= 0 && b <= 4;
}

cudaError_t DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { p = nullptr; return e; }
    cap = want;
    return cudaSuccess;
}
void DevBuf::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }

cudaError_t PinBuf::reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) { p = nullptr; return e; }
    cap = want;
    return cudaSuccess;
}
void PinBuf::release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }

cudaError_t Ctx::init(int dev) {
    device = dev;
    cudaError_t e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return e;
    for (auto& x : ev) {
        e = cudaEventCreate(&x);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// Threads for large host-side copies (staged uploads, result hand-over): half the cores, at most 8.
int host_copy_threads() {
    unsigned hw = std::thread::hardware_concurrency();
    int n = (int)(hw / 2);
    if (n > 8) n = 8;
    if (n < 1) n = 1;
    if (const char* e = getenv("WICCA_UPLOAD_THREADS")) n = atoi(e) > 0 ? atoi(e) : n;
    return n;
}

void Ctx::flush_pending() {
    size_t total = 0;
    for (const Pending& q : pending) total += q.bytes;
    int n_threads = host_copy_threads();
    if (total < ((size_t)8 << 20) || n_threads < 2) {
        for (const Pending& q : pending) memcpy(q.dst, q.src, q.bytes);
    } else {
        // large results (classifier batches: tens of MB per image, into freshly allocated pageable arrays):
        // 2 MB pieces handed round-robin to a few threads - copy bandwidth and first-touch faults both scale
        struct Piece { uint8_t* dst; const uint8_t* src; size_t bytes; };
        std::vector<Piece> pieces;
        const size_t piece = (size_t)2 << 20;
        for (const Pending& q : pending)
            for (size_t o = 0; o < q.bytes; o += piece)
                pieces.push_back({(uint8_t*)q.dst + o, (const uint8_t*)q.src + o, std::min(piece, q.bytes - o)});
        if (n_threads > (int)pieces.size()) n_threads = (int)pieces.size();
        auto work = [&](int t) {
            for (size_t k = (size_t)t; k < pieces.size(); k += (size_t)n_threads) memcpy(pieces[k].dst, pieces[k].src, pieces[k].bytes);
        };
        std::vector<std::thread> threads;
        for (int t = 1; t < n_threads; ++t) threads.emplace_back(work, t);
        work(0);
        for (auto& th : threads) th.join();
    }
    pending.clear();
}

bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

void Ctx::destroy() {
    if (device < 0) return;
    cudaSetDevice(device);
    d_src.release(); d_icons.release(); d_desc.release(); d_strip.release();
    d_f32a.release(); d_f32b.release(); d_misc.release(); d_tmp.release(); d_sum6.release();
    d_stage.release(); d_pack.release();
    h_desc.release(); h_bounce.release(); h_in.release(); h_out.release(); h_jpeg.release();
    for (auto& x : ev) if (x) { cudaEventDestroy(x); x = nullptr; }
    if (stream) { cudaStreamDestroy(stream); stream = nullptr; }
    device = -1;
}

namespace {
struct Pool {
    std::mutex mu;
    std::vector<Ctx*> idle;
    std::vector<Ctx*> all;
};
std::mutex g_mu;
int g_ndev = -2;                    // -2: not probed
std::vector<DeviceInfo> g_info;
std::vector<Pool*> g_pools;

void probe_locked() {
    if (g_ndev != -2) return;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    g_ndev = n;
    g_info.assign(n, DeviceInfo());
    g_pools.assign(n, nullptr);
    for (int i = 0; i < n; ++i) g_pools[i] = new Pool();
}
}  // namespace

int device_count_cached() {
    std::lock_guard<std::mutex> lk(g_mu);
    probe_locked();
    return g_ndev;
}

int check_device(int device) {
    const int n = device_count_cached();
    if (n <= 0) {
        int m = 0;
        cudaError_t e = cudaGetDeviceCount(&m);
        if (e == cudaSuccess) e = cudaErrorNoDevice;
        return cuda_fail(e, "no usable CUDA device (libwicca_b200 has no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail(WICCA_EDEVICE, "device ordinal %d out of range (0..%d)", device, n - 1);
    return 0;
}

const DeviceInfo& device_info(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceInfo& di = g_info[device];
    if (!di.ok) {
        int v = 0;
        // the device-pointer entry points take their scratch from the stream-ordered pool: keep a little cached
        // there instead of handing it back to the driver at every synchronisation (the default threshold is 0)
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            uint64_t cur = 0, keep = (uint64_t)64 << 20;
            if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur) == cudaSuccess && cur < keep)
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
        di.sm_count = v > 0 ? v : 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
        di.smem_optin = (size_t)v;
        di.ok = true;
    }
    return di;
}

static thread_local bool tl_link_shared = false;
ScopedLinkShared::ScopedLinkShared() : prev(tl_link_shared) { tl_link_shared = true; }
ScopedLinkShared::~ScopedLinkShared() { tl_link_shared = prev; }

int acquire_ctx(int device, Ctx** out) {
    int rc = check_device(device);
    if (rc) return rc;
    WICCA_CUDA(cudaSetDevice(device));
    Pool* pool = g_pools[device];
    {
        std::lock_guard<std::mutex> lk(pool->mu);
        if (!pool->idle.empty()) {
            *out = pool->idle.back();
            pool->idle.pop_back();
            (*out)->link_shared = tl_link_shared;
            return 0;
        }
    }
    Ctx* c = new Ctx();
    cudaError_t e = c->init(device);
    if (e != cudaSuccess) { c->destroy(); delete c; return cuda_fail(e, "context init"); }
    {
        std::lock_guard<std::mutex> lk(pool->mu);
        pool->all.push_back(c);
    }
    c->link_shared = tl_link_shared;
    *out = c;
    return 0;
}

void release_ctx(Ctx* c) {
    if (!c || c->device < 0) return;
    c->pending.clear();          // copies still owed belong to a call that failed: never replay them later
    Pool* pool = g_pools[c->device];
    std::lock_guard<std::mutex> lk(pool->mu);
    pool->idle.push_back(c);
}

void destroy_all_ctx() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ndev <= 0) return;
    for (Pool* pool : g_pools) {
        if (!pool) continue;
        std::lock_guard<std::mutex> lk2(pool->mu);
        for (Ctx* c : pool->all) { c->destroy(); delete c; }
        pool->all.clear();
        pool->idle.clear();
    }
}

// ---- NUMA locality ------------------------------------------------------------------------
namespace {
// Parses a sysfs cpulist ("0-15,64-79") into a cpu_set_t; returns the number of CPUs found.
int parse_cpulist(const char* text, cpu_set_t* set) {
    CPU_ZERO(set);
    int n = 0;
    const char* p = text;
    while (*p) {
        while (*p == ',' || *p == ' ' || *p == '\n') ++p;
        if (!*p) break;
        char* end = nullptr;
        long a = strtol(p, &end, 10);
        if (end == p) break;
        long b = a;
        p = end;
        if (*p == '-') { b = strtol(p + 1, &end, 10); p = end; }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, set); ++n; }
    }
    return n;
}

bool device_cpuset(int device, cpu_set_t* set) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return false; }
    for (char* q = bus; *q; ++q) *q = (char)tolower(*q);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    char text[1024] = {0};
    const size_t got = fread(text, 1, sizeof text - 1, f);
    fclose(f);
    if (got == 0) return false;
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return false;
    cpu_set_t local;
    if (parse_cpulist(text, &local) == 0) return false;
    CPU_AND(set, &local, &allowed);                 // stay inside the cgroup / taskset of the process
    return CPU_COUNT(set) > 0;
}
}  // namespace

ScopedAffinity::ScopedAffinity(int device) {
    static_assert(sizeof(saved) >= sizeof(cpu_set_t), "affinity mask does not fit");
    if (device < 0 || getenv("WICCA_NO_NUMA_BIND")) return;
    cpu_set_t want;
    if (!device_cpuset(device, &want)) return;
    cpu_set_t old;
    if (sched_getaffinity(0, sizeof old, &old) != 0) return;
    if (sched_setaffinity(0, sizeof want, &want) != 0) return;
    memcpy(saved, &old, sizeof old);
    active = true;
}

ScopedAffinity::~ScopedAffinity() {
    if (!active) return;
    cpu_set_t old;
    memcpy(&old, saved, sizeof old);
    sched_setaffinity(0, sizeof old, &old);
}

int upload_image_async(Ctx& c, const uint8_t* src, int H, int64_t row_bytes, int64_t stride, int64_t pitch) {
    const size_t total = (size_t)row_bytes * H;
    if (total < ((size_t)4 << 20) || is_pinned_host(src)) {
        if (c.link_shared && stride == row_bytes && total >= kFlatCopyMin) {
            if (pitch == row_bytes) {          // already the device layout
                WICCA_CUDA(cudaMemcpyAsync(c.d_src.p, src, total, cudaMemcpyHostToDevice, c.stream));
                return 0;
            }
            if (is_pinned_host(src)) {         // flat over the link, spread to the pitch on the device
                WICCA_CUDA(c.d_stage.reserve(total));
                WICCA_CUDA(cudaMemcpyAsync(c.d_stage.p, src, total, cudaMemcpyHostToDevice, c.stream));
                WICCA_CUDA(launch_copy_rows(c.d_src.p, pitch, c.d_stage.p, row_bytes, row_bytes, H, c.stream));
                return 0;
            }
        }
        WICCA_CUDA(cudaMemcpy2DAsync(c.d_src.p, (size_t)pitch, src, (size_t)stride, (size_t)row_bytes, (size_t)H,
                                     cudaMemcpyHostToDevice, c.stream));
        return 0;
    }
    WICCA_CUDA(c.h_in.reserve(total));
    int rows_per_band = (int)(((size_t)4 << 20) / (size_t)row_bytes);
    if (rows_per_band < 1) rows_per_band = 1;
    const int n_bands = (H + rows_per_band - 1) / rows_per_band;
    int n_threads = host_copy_threads();               // staging is a plain memcpy: bandwidth scales with cores
    if (n_threads > n_bands) n_threads = n_bands;
    std::vector<cudaError_t> errs(n_threads, cudaSuccess);
    auto work = [&](int t) {
        if (cudaSetDevice(c.device) != cudaSuccess) { errs[t] = cudaGetLastError(); return; }
        for (int b = t; b < n_bands; b += n_threads) {
            const int r0 = b * rows_per_band;
            const int nr = (r0 + rows_per_band <= H) ? rows_per_band : H - r0;
            uint8_t* stage = (uint8_t*)c.h_in.p + (size_t)r0 * row_bytes;
            if (stride == row_bytes) memcpy(stage, src + (size_t)r0 * stride, (size_t)nr * row_bytes);
            else for (int y = 0; y < nr; ++y) memcpy(stage + (size_t)y * row_bytes, src + (size_t)(r0 + y) * stride, (size_t)row_bytes);
            cudaError_t e = cudaMemcpy2DAsync((uint8_t*)c.d_src.p + (size_t)r0 * pitch, (size_t)pitch, stage, (size_t)row_bytes,
                                              (size_t)row_bytes, (size_t)nr, cudaMemcpyHostToDevice, c.stream);
            if (e != cudaSuccess) { errs[t] = e; return; }
        }
    };
    std::vector<std::thread> threads;
    for (int t = 1; t < n_threads; ++t) threads.emplace_back(work, t);
    work(0);
    for (auto& th : threads) th.join();
    for (cudaError_t e : errs)
        if (e != cudaSuccess) return cuda_fail(e, "staged image upload");
    return 0;
}

int download_rows_async(Ctx& c, void* h_dst, const void* d_src, int64_t d_pitch, int64_t row_bytes, int rows, size_t* pack_off) {
    const size_t total = (size_t)row_bytes * rows;
    if (d_pitch == row_bytes) {
        WICCA_CUDA(cudaMemcpyAsync(h_dst, d_src, total, cudaMemcpyDeviceToHost, c.stream));
    } else if (c.link_shared && total >= kFlatCopyMin && pack_off && *pack_off + total <= c.d_pack.cap) {
        uint8_t* tight = (uint8_t*)c.d_pack.p + *pack_off;
        *pack_off += (size_t)align_up((int64_t)total, 256);
        WICCA_CUDA(launch_copy_rows(tight, row_bytes, d_src, d_pitch, row_bytes, rows, c.stream));
        WICCA_CUDA(cudaMemcpyAsync(h_dst, tight, total, cudaMemcpyDeviceToHost, c.stream));
    } else {
        WICCA_CUDA(cudaMemcpy2DAsync(h_dst, (size_t)row_bytes, d_src, (size_t)d_pitch, (size_t)row_bytes, (size_t)rows,
                                     cudaMemcpyDeviceToHost, c.stream));
    }
    return 0;
}

// ---- TMA descriptor ---------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
        else cudaGetLastError();
    });
    return fn;
}

int encode_image_tmap(CUtensorMap* tm, const void* d_src, int H, int64_t pitch) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail((int)cudaErrorNotSupported, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t gdim[2] = {(cuuint64_t)(pitch / 4), (cuuint64_t)H};
    cuuint64_t gstride[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)(kStageRowBytes / 4), (cuuint32_t)kItemH};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(d_src), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail((int)cudaErrorInvalidValue, "cuTensorMapEncodeTiled failed with CUresult %d (H=%d pitch=%lld)", (int)r, H,
                    (long long)pitch);
    return 0;
}

// uint8 view (w_bytes x h) of a pitched icon, box box_w x box_h bytes/rows: target of TMA store.
int encode_icon_tmap(CUtensorMap* tm, const void* d_icon, int h, int64_t w_bytes, int64_t pitch, int box_w, int box_h) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail((int)cudaErrorNotSupported, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t gdim[2] = {(cuuint64_t)w_bytes, (cuuint64_t)h};
    cuuint64_t gstride[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(d_icon), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail((int)cudaErrorInvalidValue, "cuTensorMapEncodeTiled (icon) failed with CUresult %d (h=%d w=%lld pitch=%lld)",
                    (int)r, h, (long long)w_bytes, (long long)pitch);
    return 0;
}

int icon_variant_from_env() {
#ifdef WICCA_DEV
    const char* e = getenv("WICCA_ICON_VARIANT");   // developer knob (tools/), see haar_icon.cu launch_icon_tma
    const int v = e ? atoi(e) : 0;
    return v < 0 ? 0 : v;
#else
    return 0;                                       // the release library has no kernel knobs
#endif
}

}  // namespace wicca
