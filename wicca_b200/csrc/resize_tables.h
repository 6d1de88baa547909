// resize_tables.h - host-side construction of the INTER_AREA tap tables exactly as OpenCV builds
// them (double arithmetic for the geometry, float32 taps), shared by the host-buffer entry point
// and the device-resident plan epilogue.  See oracle/resize_oracle.py for the restated algorithm.
#pragma once
#include <math.h>
#include <string.h>

#include <map>
#include <utility>
#include <vector>

#include "host_common.h"
#include "kernels.h"

namespace wicca {

struct ResizeSrc { const uint8_t* d_ptr; int h, w; int64_t pitch; };

struct ResizeTableBlob {
    std::vector<uint8_t> bytes;          // [jobs][area][lin], 256-byte aligned sections
    size_t o_jobs = 0, o_area = 0, o_ln = 0;
    int max_src_w = 0, n_area = 0, n_other = 0;      // launch hints: widest source, images per regime class
    ResizeTables view(const void* device_base) const {
        const uint8_t* b = (const uint8_t*)device_base;
        ResizeTables t;
        t.jobs = (const ResizeJob*)(b + o_jobs);
        t.area = (const AreaDesc*)(b + o_area);
        t.lin = (const LinTap*)(b + o_ln);
        return t;
    }
};

// computeResizeAreaTab, one descriptor per destination index.
inline void area_tab(int ssize, int dsize, double scale, std::vector<AreaDesc>& out) {
    for (int dx = 0; dx < dsize; ++dx) {
        const double fsx1 = dx * scale;
        const double fsx2 = fsx1 + scale;
        const double cell = fmin(scale, ssize - fsx1);
        int sx1 = (int)ceil(fsx1);
        int sx2 = (int)floor(fsx2);
        if (sx2 > ssize - 1) sx2 = ssize - 1;
        if (sx1 > sx2) sx1 = sx2;
        AreaDesc d;
        memset(&d, 0, sizeof d);
        d.s_first = sx1;
        d.n_full = sx2 - sx1;
        d.w_full = (float)(1.0 / cell);
        d.s_left = sx1 > 0 ? sx1 - 1 : 0;
        d.s_right = sx2;
        if (sx1 - fsx1 > 1e-3) d.w_left = (float)((sx1 - fsx1) / cell);
        if (fsx2 - sx2 > 1e-3) d.w_right = (float)(fmin(fmin(fsx2 - sx2, 1.0), cell) / cell);
        out.push_back(d);
    }
}

// INTER_LINEAR tap computation in "area mode" (used by INTER_AREA when an axis is upscaled).
inline void linear_tab(int ssize, int dsize, double inv, double scale, std::vector<LinTap>& lin) {
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

inline ResizeTableBlob build_resize_tables(const std::vector<ResizeSrc>& srcs, int out_h, int out_w) {
    std::vector<ResizeJob> jobs(srcs.size());
    std::vector<AreaDesc> area;
    std::vector<LinTap> lin;
    // a tap table depends only on (source length, target length): images of equal size share theirs, which
    // keeps the blob of a uniform batch small enough to be staged inline by the H2D copy
    std::map<std::pair<int, int>, int> area_at, lin_at;
    auto area_for = [&](int ssize, int dsize, double scale) {
        auto it = area_at.find({ssize, dsize});
        if (it != area_at.end()) return it->second;
        const int off = (int)area.size();
        area_tab(ssize, dsize, scale, area);
        area_at[{ssize, dsize}] = off;
        return off;
    };
    auto lin_for = [&](int ssize, int dsize, double inv, double scale) {
        auto it = lin_at.find({ssize, dsize});
        if (it != lin_at.end()) return it->second;
        const int off = (int)lin.size();
        linear_tab(ssize, dsize, inv, scale, lin);
        lin_at[{ssize, dsize}] = off;
        return off;
    };
    for (size_t i = 0; i < srcs.size(); ++i) {
        ResizeJob& j = jobs[i];
        memset(&j, 0, sizeof j);
        j.src = srcs[i].d_ptr; j.pitch = srcs[i].pitch; j.sh = srcs[i].h; j.sw = srcs[i].w;
        const double inv_x = (double)out_w / j.sw, inv_y = (double)out_h / j.sh;
        const double sx = 1.0 / inv_x, sy = 1.0 / inv_y;      // OpenCV: scale = 1/inv, not ssize/dsize
        if (j.sw == out_w && j.sh == out_h) {
            j.regime = 0;
        } else if (sx >= 1 && sy >= 1) {
            const int isx = (int)nearbyint(sx), isy = (int)nearbyint(sy);
            if (fabs(sx - isx) < 2.220446049250313e-16 && fabs(sy - isy) < 2.220446049250313e-16) {
                j.regime = 1; j.isx = isx; j.isy = isy;
            } else {
                j.regime = 2;
                j.xoff = area_for(j.sw, out_w, sx);
                j.yoff = area_for(j.sh, out_h, sy);
            }
        } else {
            j.regime = 3;
            j.xoff = lin_for(j.sw, out_w, inv_x, sx);
            j.yoff = lin_for(j.sh, out_h, inv_y, sy);
        }
    }
    ResizeTableBlob b;
    for (const ResizeJob& jj : jobs) {
        if (jj.sw > b.max_src_w) b.max_src_w = jj.sw;
        if (jj.regime == 2) ++b.n_area; else ++b.n_other;
    }
    b.o_jobs = 0;
    b.o_area = (size_t)align_up((int64_t)(b.o_jobs + jobs.size() * sizeof(ResizeJob)), 256);
    b.o_ln = (size_t)align_up((int64_t)(b.o_area + area.size() * sizeof(AreaDesc)), 256);
    b.bytes.assign(b.o_ln + lin.size() * sizeof(LinTap) + 256, 0);
    memcpy(b.bytes.data() + b.o_jobs, jobs.data(), jobs.size() * sizeof(ResizeJob));
    if (!area.empty()) memcpy(b.bytes.data() + b.o_area, area.data(), area.size() * sizeof(AreaDesc));
    if (!lin.empty()) memcpy(b.bytes.data() + b.o_ln, lin.data(), lin.size() * sizeof(LinTap));
    return b;
}

}  // namespace wicca
