// capi_resize.cu - C ABI of the resize + normalise epilogue (rows A5/A6).  The host builds the
// per-axis tap tables exactly as OpenCV does (double arithmetic for the geometry, float32 taps).
#include <math.h>
#include <string.h>

#include <vector>

#include "host_common.h"
#include "kernels.h"

using namespace wicca;

namespace {

// computeResizeAreaTab: CSR list of (src index, fp32 weight) per destination index.
void area_tab(int ssize, int dsize, double scale, std::vector<int>& rowptr, std::vector<AreaTap>& taps) {
    for (int dx = 0; dx < dsize; ++dx) {
        rowptr.push_back((int)taps.size());
        const double fsx1 = dx * scale;
        const double fsx2 = fsx1 + scale;
        const double cell = fmin(scale, ssize - fsx1);
        int sx1 = (int)ceil(fsx1);
        int sx2 = (int)floor(fsx2);
        if (sx2 > ssize - 1) sx2 = ssize - 1;
        if (sx1 > sx2) sx1 = sx2;
        if (sx1 - fsx1 > 1e-3) taps.push_back({sx1 - 1, (float)((sx1 - fsx1) / cell)});
        for (int s = sx1; s < sx2; ++s) taps.push_back({s, (float)(1.0 / cell)});
        if (fsx2 - sx2 > 1e-3) taps.push_back({sx2, (float)(fmin(fmin(fsx2 - sx2, 1.0), cell) / cell)});
    }
    rowptr.push_back((int)taps.size());
}

// INTER_LINEAR tap computation in "area mode" (used by INTER_AREA when an axis is upscaled).
void linear_tab(int ssize, int dsize, double inv, double scale, std::vector<LinTap>& lin) {
    for (int d = 0; d < dsize; ++d) {
        int s = (int)floor(d * scale);
        float f = (float)((d + 1) - (s + 1) * inv);
        f = f <= 0 ? 0.f : f - floorf(f);
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        LinTap t;
        t.i0 = s;
        t.i1 = (s + 1 < ssize) ? s + 1 : ssize - 1;
        t.c0 = (int)nearbyintf((1.f - f) * 2048.f);
        t.c1 = (int)nearbyintf(f * 2048.f);
        lin.push_back(t);
    }
}

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
        src_bytes += (size_t)align_up((int64_t)hs[i] * ws[i] * 3, 256);
    }
    CtxLease L;
    int rc = acquire_ctx(device, &L.c);
    if (rc) return rc;
    Ctx& c = *L.c;

    std::vector<ResizeJob> jobs(n);
    std::vector<int> rowptr;
    std::vector<AreaTap> taps;
    std::vector<LinTap> lin;
    WICCA_CUDA(c.d_src.reserve(src_bytes + 256));
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
        ResizeJob& j = jobs[i];
        memset(&j, 0, sizeof j);
        j.src = (const uint8_t*)c.d_src.p + off;
        off += (size_t)align_up((int64_t)hs[i] * ws[i] * 3, 256);
        j.sh = hs[i]; j.sw = ws[i];
        const double inv_x = (double)out_w / ws[i], inv_y = (double)out_h / hs[i];
        const double sx = 1.0 / inv_x, sy = 1.0 / inv_y;      // OpenCV: scale = 1/inv, not ssize/dsize
        if (ws[i] == out_w && hs[i] == out_h) {
            j.regime = 0;
        } else if (sx >= 1 && sy >= 1) {
            const int isx = (int)nearbyint(sx), isy = (int)nearbyint(sy);
            if (fabs(sx - isx) < 2.220446049250313e-16 && fabs(sy - isy) < 2.220446049250313e-16) {
                j.regime = 1; j.isx = isx; j.isy = isy;
            } else {
                j.regime = 2;
                const int base = (int)rowptr.size();
                std::vector<int> rp; std::vector<AreaTap> tp;
                area_tab(ws[i], out_w, sx, rp, tp);
                j.xoff = base;
                for (int v : rp) rowptr.push_back(v + (int)taps.size());
                taps.insert(taps.end(), tp.begin(), tp.end());
                rp.clear(); tp.clear();
                area_tab(hs[i], out_h, sy, rp, tp);
                j.yoff = (int)rowptr.size();
                for (int v : rp) rowptr.push_back(v + (int)taps.size());
                taps.insert(taps.end(), tp.begin(), tp.end());
            }
        } else {
            j.regime = 3;
            j.xoff = (int)lin.size();
            linear_tab(ws[i], out_w, inv_x, sx, lin);
            j.yoff = (int)lin.size();
            linear_tab(hs[i], out_h, inv_y, sy, lin);
        }
    }
    // one table upload: [jobs][rowptr][taps][lin]
    const size_t o_jobs = 0;
    const size_t o_rp = (size_t)align_up((int64_t)(o_jobs + jobs.size() * sizeof(ResizeJob)), 256);
    const size_t o_tp = (size_t)align_up((int64_t)(o_rp + rowptr.size() * sizeof(int)), 256);
    const size_t o_ln = (size_t)align_up((int64_t)(o_tp + taps.size() * sizeof(AreaTap)), 256);
    const size_t tbl_bytes = o_ln + lin.size() * sizeof(LinTap) + 256;
    WICCA_CUDA(c.h_desc.reserve(tbl_bytes));
    WICCA_CUDA(c.d_desc.reserve(tbl_bytes));
    uint8_t* hb = (uint8_t*)c.h_desc.p;
    memcpy(hb + o_jobs, jobs.data(), jobs.size() * sizeof(ResizeJob));
    if (!rowptr.empty()) memcpy(hb + o_rp, rowptr.data(), rowptr.size() * sizeof(int));
    if (!taps.empty()) memcpy(hb + o_tp, taps.data(), taps.size() * sizeof(AreaTap));
    if (!lin.empty()) memcpy(hb + o_ln, lin.data(), lin.size() * sizeof(LinTap));

    const size_t out_elems = (size_t)n * out_h * out_w * 3;
    WICCA_CUDA(c.d_f32a.reserve(out_elems * sizeof(float)));
    if (dst_u8) WICCA_CUDA(c.d_misc.reserve(out_elems));

    WICCA_CUDA(cudaEventRecord(c.ev[0], c.stream));
    WICCA_CUDA(cudaMemcpyAsync(c.d_desc.p, c.h_desc.p, tbl_bytes, cudaMemcpyHostToDevice, c.stream));
    for (int i = 0; i < n; ++i)
        WICCA_CUDA(cudaMemcpyAsync((void*)jobs[i].src, icons[i], (size_t)hs[i] * ws[i] * 3, cudaMemcpyHostToDevice, c.stream));
    WICCA_CUDA(cudaEventRecord(c.ev[1], c.stream));
    ResizeTables tb;
    uint8_t* db = (uint8_t*)c.d_desc.p;
    tb.jobs = (const ResizeJob*)(db + o_jobs);
    tb.rowptr = (const int*)(db + o_rp);
    tb.taps = (const AreaTap*)(db + o_tp);
    tb.lin = (const LinTap*)(db + o_ln);
    cudaError_t e = launch_resize_norm(tb, n, out_h, out_w, norm_mode, (float*)c.d_f32a.p,
                                       dst_u8 ? (uint8_t*)c.d_misc.p : nullptr, c.stream);
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
