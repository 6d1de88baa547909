// capi_icon.cu - C ABI for the icon path: HaarCoder.get_small_copy and its multi-depth,
// device-resident and batch-plan variants.  See include/wicca_b200.h for the contract and the
// reference lines each entry point replaces.
#include <string.h>

#include <algorithm>
#include <vector>

#include "host_common.h"
#include "icon_types.h"
#include "kernels.h"
#include "resize_tables.h"

#include <map>
#include <tuple>

using namespace wicca;

namespace {

bool needs_padding(int H, int W, int depth) {
    if (depth <= 0) return false;
    const int64_t r = (int64_t)1 << depth;
    return (H % r) != 0 || (W % r) != 0;
}

// Common argument validation of the icon entry points.  Mirrors the order of checks of the
// reference: validate_image (validation.py:94-99) then get_padded_copy (data_loader.py:93-117).
int validate_icon_args_impl(const void* src, int H, int W, int C, const int* depths, int n_depths, int border_type) {
    if (!src) return fail(WICCA_EINVAL, "image pointer is NULL");
    if (H <= 0 || W <= 0 || C <= 0) return fail(WICCA_EINVAL, "image is empty (H=%d W=%d C=%d)", H, W, C);
    if (!depths || n_depths <= 0) return fail(WICCA_EINVAL, "no depths given");
    for (int i = 0; i < n_depths; ++i) {
        if (depths[i] > WICCA_MAX_DEPTH) return fail(WICCA_EDEPTH, "depth %d exceeds WICCA_MAX_DEPTH=%d", depths[i], WICCA_MAX_DEPTH);
        if (needs_padding(H, W, depths[i])) {
            if (!border_valid(border_type)) return fail(WICCA_EBORDER, "unsupported border type %d", border_type);
            if (C > 4) return fail(WICCA_ECHANNELS, "padding needs C <= 4 (cv2.copyMakeBorder), got C=%d", C);
        }
    }
    return 0;
}

bool fused_eligible(const void* d_src, int64_t pitch, int C) {
    return C == 3 && ((uintptr_t)d_src % 16) == 0 && (pitch % 16) == 0;
}

// Fill one IconImage (including its tensor map).  outs: the fused outputs (depth 1..6) of this image.
int fill_icon_image(IconImage* im, const uint8_t* d_src, int H, int W, int64_t pitch, const IconOut* outs, int n_outs,
                    int item_base, uint32_t* sum6 = nullptr, int sum6_stride = 0) {
    memset(im, 0, sizeof(*im));
    int rc = encode_image_tmap(&im->tmap, d_src, H, pitch);
    if (rc) return rc;
    icon_image_geometry(im, d_src, H, W, pitch, item_base);
    if (sum6) icon_image_add_sum6(im, sum6, sum6_stride);
    for (int i = 0; i < n_outs; ++i) {
        if (outs[i].depth > kMaxFused) continue;     // finished from the level-6 plane by haar_tail_kernel
        icon_image_add_level(im, outs[i].depth, outs[i].d_ptr, outs[i].pitch);
        const int d = outs[i].depth;
        if (d <= 3) {   // levels 1..3 leave the kernel through TMA store
            static const int box_w[3] = {kOut1Row, kOut2Row, kOut3Row};
            static const int box_h[3] = {16, 8, 4};
            rc = encode_icon_tmap(&im->hmap[d - 1], outs[i].d_ptr, outs[i].h, (int64_t)outs[i].w * 3, outs[i].pitch,
                                  box_w[d - 1], box_h[d - 1]);
            if (rc) return rc;
        }
    }
    return 0;
}

int strip_px(int W, int Wp_max) { return Wp_max - (W & ~(kChunkPx - 1)); }
// REPLICATE and CONSTANT borders are resolved from the shared-memory stage; only the mirroring /
// wrapping borders need the pre-built right strip.
bool border_needs_strip(int border_type) { const int b = border_base(border_type); return b == 2 || b == 3 || b == 4; }

// evict_first on the input stream helps only when the output stream is negligible (no level 1 or 2)
int stream_hint_for(const IconOut* outs, int n) {
    for (int i = 0; i < n; ++i)
        if (outs[i].depth >= 1 && outs[i].depth <= 2) return 0;
    return 1;
}

int64_t icon_pitch_for(int w, int C) { return align_up((int64_t)w * C, 128); }

// Enqueue the general path for one depth (1..WICCA_MAX_DEPTH).  f32a/f32b: ping-pong scratch for
// depths > 8 (each >= level8_elems floats).
int enqueue_generic(const uint8_t* d_src, int64_t pitch, int H, int W, int C, int depth, int border_type, int bconst,
                    uint8_t* d_dst, int64_t dst_pitch, float* f32a, float* f32b, cudaStream_t stream) {
    GenericIconArgs a;
    a.src = d_src; a.pitch = pitch; a.H = H; a.W = W; a.C = C;
    a.border_type = border_base(border_type); a.border_const = bconst;
    if (depth <= 8) {
        a.depth = depth;
        a.out_h = icon_dim(H, depth); a.out_w = icon_dim(W, depth);
        a.dst_u8 = d_dst; a.dst_pitch = dst_pitch; a.dst_f32 = nullptr;
        cudaError_t e = launch_icon_generic(a, stream);
        if (e != cudaSuccess) return cuda_fail(e, "generic icon kernel");
        return 0;
    }
    // depth > 8: exact level-8 values as fp32 over the grid padded for the FINAL depth, then
    // replay the remaining levels in fp32 exactly as the reference does.
    const int oh = icon_dim(H, depth), ow = icon_dim(W, depth);
    int lh = oh << (depth - 8), lw = ow << (depth - 8);
    a.depth = 8; a.out_h = lh; a.out_w = lw; a.dst_u8 = nullptr; a.dst_pitch = 0; a.dst_f32 = f32a;
    cudaError_t e = launch_icon_generic(a, stream);
    if (e != cudaSuccess) return cuda_fail(e, "generic icon kernel (level 8)");
    float* in = f32a;
    float* out = f32b;
    for (int l = 9; l <= depth; ++l) {
        lh >>= 1; lw >>= 1;
        const bool last = (l == depth);
        if (last && dst_pitch != (int64_t)lw * C)
            return fail(WICCA_EINVAL, "depth > 8 needs a tight destination");
        e = launch_level_f32(in, out, last ? d_dst : nullptr, lh, lw, C, stream);
        if (e != cudaSuccess) return cuda_fail(e, "fp32 level kernel");
        std::swap(in, out);
    }
    return 0;
}

// Depth 7..WICCA_MAX_DEPTH of a 3-channel image whose exact level-6 block sums the one-pass kernel has left in
// `sum6` ((ceil(H/64), ceil(W/64), 3) uint32): no second pass over the image.  f32a/f32b as in enqueue_generic.
struct TailGeom { int s6_h, s6_w, ext_h, ext_w; };
// level-6 blocks the one-pass kernel produces (s6_*) and those of the padded grid of `max_depth` (ext_*)
TailGeom tail_geom(int H, int W, int max_depth) {
    TailGeom g;
    g.s6_h = icon_dim(H, 6); g.s6_w = icon_dim(W, 6);
    g.ext_h = icon_dim(H, max_depth) << (max_depth - 6); g.ext_w = icon_dim(W, max_depth) << (max_depth - 6);
    return g;
}
size_t sum6_bytes(const TailGeom& g) { return (size_t)g.ext_h * g.ext_w * 3 * sizeof(uint32_t); }
TailArgs tail_base(const uint8_t* d_src, int64_t pitch, int H, int W, int border_type, int bconst, uint32_t* sum6, const TailGeom& g) {
    TailArgs a;
    memset(&a, 0, sizeof a);
    a.src = d_src; a.pitch = pitch; a.H = H; a.W = W;
    a.border_type = border_base(border_type); a.border_const = bconst;
    a.sum6 = sum6; a.s6_h = g.s6_h; a.s6_w = g.s6_w; a.ext_h = g.ext_h; a.ext_w = g.ext_w;
    return a;
}
int enqueue_tail(const TailArgs& base, int depth, uint8_t* d_dst, int64_t dst_pitch, float* f32a, float* f32b, cudaStream_t stream) {
    TailArgs a = base;
    if (depth <= 8) {
        a.depth = depth; a.out_h = icon_dim(a.H, depth); a.out_w = icon_dim(a.W, depth);
        a.dst_u8 = d_dst; a.dst_pitch = dst_pitch; a.dst_f32 = nullptr;
        cudaError_t e = launch_icon_tail(a, stream);
        if (e != cudaSuccess) return cuda_fail(e, "icon tail kernel");
        return 0;
    }
    const int oh = icon_dim(a.H, depth), ow = icon_dim(a.W, depth);
    int lh = oh << (depth - 8), lw = ow << (depth - 8);
    a.depth = 8; a.out_h = lh; a.out_w = lw; a.dst_u8 = nullptr; a.dst_pitch = 0; a.dst_f32 = f32a;
    cudaError_t e = launch_icon_tail(a, stream);
    if (e != cudaSuccess) return cuda_fail(e, "icon tail kernel (level 8)");
    float* in = f32a;
    float* out = f32b;
    for (int l = 9; l <= depth; ++l) {
        lh >>= 1; lw >>= 1;
        const bool last = (l == depth);
        if (last && dst_pitch != (int64_t)lw * 3) return fail(WICCA_EINVAL, "depth > 8 needs a tight destination");
        e = launch_level_f32(in, out, last ? d_dst : nullptr, lh, lw, 3, stream);
        if (e != cudaSuccess) return cuda_fail(e, "fp32 level kernel");
        std::swap(in, out);
    }
    return 0;
}

size_t level8_elems(int H, int W, int C, int depth) {
    if (depth <= 8) return 0;
    const int64_t oh = icon_dim(H, depth), ow = icon_dim(W, depth);
    return (size_t)((oh << (depth - 8)) * (ow << (depth - 8)) * C);
}

void fill_timing(wicca_timing* t, Ctx& c) {
    if (!t) return;
    t->h2d_ms = t->kernel_ms = t->d2h_ms = t->total_ms = 0.f;
    cudaEventElapsedTime(&t->h2d_ms, c.ev[0], c.ev[1]);
    cudaEventElapsedTime(&t->kernel_ms, c.ev[1], c.ev[2]);
    cudaEventElapsedTime(&t->d2h_ms, c.ev[2], c.ev[3]);
    cudaEventElapsedTime(&t->total_ms, c.ev[0], c.ev[3]);
}

}  // namespace

namespace wicca {

int validate_icon_args(const void* src, int H, int W, int C, const int* depths, int n_depths, int border_type) {
    return validate_icon_args_impl(src, H, W, C, depths, n_depths, border_type);
}

// Shared by the one-shot host call and the batch workers: image already on the device in
// c.d_src (pitched); computes every requested depth > 0 and copies the icons to the host
// destinations.  Enqueues on c.stream; records ev[2] after the kernels and ev[3] after D2H.
int enqueue_icons_resident(Ctx& c, int H, int W, int C, int64_t pitch, const int* depths, int n_depths, int border_type,
                           int bconst, std::vector<IconOut>& outs) {
    const uint8_t* d_src = (const uint8_t*)c.d_src.p;
    outs.assign(n_depths, IconOut());
    std::vector<int> fused, tail, generic;
    const bool can_fuse = fused_eligible(d_src, pitch, C);
    size_t icon_bytes = 0, f32_elems = 0;
    for (int i = 0; i < n_depths; ++i) {
        const int d = depths[i];
        if (d <= 0) continue;
        outs[i].depth = d;
        outs[i].h = icon_dim(H, d);
        outs[i].w = icon_dim(W, d);
        outs[i].pitch = d > 8 ? (int64_t)outs[i].w * C : icon_pitch_for(outs[i].w, C);
        icon_bytes = (size_t)align_up((int64_t)icon_bytes, 256);
        icon_bytes += (size_t)outs[i].pitch * outs[i].h;
        bool dup = false;   // the fused kernel has one slot per level: duplicates go the general way
        for (int j : fused) dup |= (depths[j] == d);
        if (can_fuse && d <= kMaxFused && !dup) fused.push_back(i);
        else if (can_fuse && d > kMaxFused) { tail.push_back(i); f32_elems = std::max(f32_elems, level8_elems(H, W, C, d)); }
        else { generic.push_back(i); f32_elems = std::max(f32_elems, level8_elems(H, W, C, d)); }
    }
    WICCA_CUDA(c.d_icons.reserve(icon_bytes + 256));
    if (f32_elems) {
        WICCA_CUDA(c.d_f32a.reserve(f32_elems * sizeof(float)));
        WICCA_CUDA(c.d_f32b.reserve(f32_elems * sizeof(float) / 4 + 16));
    }
    int tail_max = 0;
    for (int i : tail) tail_max = std::max(tail_max, depths[i]);
    const TailGeom tg = tail_geom(H, W, tail.empty() ? 6 : tail_max);
    if (!tail.empty()) WICCA_CUDA(c.d_sum6.reserve(sum6_bytes(tg)));
    size_t off = 0;
    for (int i = 0; i < n_depths; ++i) {
        if (depths[i] <= 0) continue;
        off = (size_t)align_up((int64_t)off, 256);
        outs[i].d_ptr = (uint8_t*)c.d_icons.p + off;
        off += (size_t)outs[i].pitch * outs[i].h;
    }
    if (!fused.empty() || !tail.empty()) {
        std::vector<IconOut> fo;
        for (int i : fused) fo.push_back(outs[i]);
        WICCA_CUDA(c.h_desc.reserve(sizeof(IconImage) + 64));
        WICCA_CUDA(c.d_desc.reserve(sizeof(IconImage) + 64));
        IconImage* him = (IconImage*)c.h_desc.p;
        int rc = fill_icon_image(him, d_src, H, W, pitch, fo.data(), (int)fo.size(), 0, tail.empty() ? nullptr : (uint32_t*)c.d_sum6.p,
                                 tg.ext_w);
        if (rc) return rc;
        uint8_t** h_strip = (uint8_t**)((uint8_t*)c.h_desc.p + sizeof(IconImage));
        *h_strip = nullptr;
        if (border_needs_strip(border_type) && strip_px(W, him->Wp_max) > 0) {
            WICCA_CUDA(c.d_strip.reserve((size_t)H * kStripPitch));
            *h_strip = (uint8_t*)c.d_strip.p;
        }
        WICCA_CUDA(cudaMemcpyAsync(c.d_desc.p, c.h_desc.p, sizeof(IconImage) + sizeof(uint8_t*), cudaMemcpyHostToDevice,
                                   c.stream));
        const IconImage* d_im = (const IconImage*)c.d_desc.p;
        uint8_t* const* d_strips = (uint8_t* const*)((uint8_t*)c.d_desc.p + sizeof(IconImage));
        if (*h_strip) {
            cudaError_t e = launch_edge_strips(d_im, d_strips, 1, H, border_base(border_type), bconst, c.stream);
            if (e != cudaSuccess) return cuda_fail(e, "edge strip kernel");
        }
        cudaError_t e = launch_icon_tma(d_im, d_strips, 1, him->items_x * him->items_y, border_base(border_type), bconst,
                                        device_info(c.device).sm_count, icon_variant_from_env(),
                                        stream_hint_for(fo.data(), (int)fo.size()), c.stream);
        if (e != cudaSuccess) return cuda_fail(e, "fused icon kernel");
        if (!tail.empty()) {
            const TailArgs tb = tail_base(d_src, pitch, H, W, border_type, bconst, (uint32_t*)c.d_sum6.p, tg);
            e = launch_icon_tail_fill(tb, c.stream);
            if (e != cudaSuccess) return cuda_fail(e, "icon tail fill kernel");
            for (int i : tail) {
                rc = enqueue_tail(tb, depths[i], outs[i].d_ptr, outs[i].pitch, (float*)c.d_f32a.p, (float*)c.d_f32b.p, c.stream);
                if (rc) return rc;
            }
        }
    }
    for (int i : generic) {
        int rc = enqueue_generic(d_src, pitch, H, W, C, depths[i], border_type, bconst, outs[i].d_ptr, outs[i].pitch,
                                 (float*)c.d_f32a.p, (float*)c.d_f32b.p, c.stream);
        if (rc) return rc;
    }
    WICCA_CUDA(cudaEventRecord(c.ev[2], c.stream));
    return 0;
}

int icons_from_resident(Ctx& c, int H, int W, int C, int64_t pitch, const int* depths, int n_depths, int border_type,
                        int bconst, uint8_t* const* dsts) {
    std::vector<IconOut> outs;
    int rc0 = enqueue_icons_resident(c, H, W, C, pitch, depths, n_depths, border_type, bconst, outs);
    if (rc0) return rc0;
    // D2H: straight into page-locked destinations; pageable ones go through the pinned bounce buffer
    // (a pageable cudaMemcpyAsync would block the host until the whole stream has drained).
    size_t bounce = 0;
    std::vector<char> direct(n_depths, 1);
    for (int i = 0; i < n_depths; ++i) {
        if (depths[i] <= 0) continue;
        if (!is_pinned_host(dsts[i])) { direct[i] = 0; bounce += (size_t)outs[i].w * C * outs[i].h; }
    }
    if (bounce) {
        c.flush_pending();                           // never reallocate under copies that are still owed
        WICCA_CUDA(c.h_bounce.reserve(bounce));
    }
    size_t boff = 0, pack_bytes = 0, pack_off = 0;
    for (int i = 0; i < n_depths; ++i)             // large icons are packed tight on the device and cross the link flat
        if (c.link_shared && depths[i] > 0 && (size_t)outs[i].w * C * outs[i].h >= kFlatCopyMin && outs[i].pitch != (int64_t)outs[i].w * C)
            pack_bytes += (size_t)align_up((int64_t)outs[i].w * C * outs[i].h, 256);
    if (pack_bytes) WICCA_CUDA(c.d_pack.reserve(pack_bytes));
    for (int i = 0; i < n_depths; ++i) {
        if (depths[i] <= 0) continue;
        const size_t rowb = (size_t)outs[i].w * C;
        uint8_t* target = dsts[i];
        if (!direct[i]) {
            target = (uint8_t*)c.h_bounce.p + boff;
            c.pending.push_back({dsts[i], target, rowb * outs[i].h});
            boff += rowb * outs[i].h;
        }
        int rc = download_rows_async(c, target, outs[i].d_ptr, outs[i].pitch, (int64_t)rowb, outs[i].h, &pack_off);
        if (rc) return rc;
    }
    WICCA_CUDA(cudaEventRecord(c.ev[3], c.stream));
    return 0;
}

}  // namespace wicca

extern "C" {

const char* wicca_version(void) { return "wicca_b200 0.1.0 (sm_100a)"; }
const char* wicca_last_error(void) { return last_error_ref().c_str(); }
int wicca_device_count(void) { return device_count_cached(); }
int wicca_shutdown(void) { resize_table_cache_clear(); destroy_all_ctx(); return 0; }

int64_t wicca_pitch_bytes(int W, int C) { return align_up((int64_t)W * C, 128); }
int wicca_icon_dim(int n, int depth) { return icon_dim(n, depth); }

int wicca_haar_icons_multi_u8(const uint8_t* src, int H, int W, int C, int64_t src_row_stride, const int* depths,
                              int n_depths, int border_type, double border_const, uint8_t* const* dsts, int device,
                              wicca_timing* t) {
    int rc = validate_icon_args(src, H, W, C, depths, n_depths, border_type);
    if (rc) return rc;
    if (!dsts) return fail(WICCA_EINVAL, "dsts is NULL");
    for (int i = 0; i < n_depths; ++i)
        if (!dsts[i]) return fail(WICCA_EINVAL, "dsts[%d] is NULL", i);
    const int64_t rowb = (int64_t)W * C;
    if (src_row_stride == 0) src_row_stride = rowb;
    if (src_row_stride < rowb) return fail(WICCA_EINVAL, "row stride %lld < W*C = %lld", (long long)src_row_stride, (long long)rowb);
    if (t) memset(t, 0, sizeof(*t));

    bool device_work = false;
    for (int i = 0; i < n_depths; ++i) {
        if (depths[i] <= 0) {   // reference: depth <= 0 runs zero levels -> a copy (wavelet_coder.py:61,67)
            for (int y = 0; y < H; ++y) memcpy(dsts[i] + (size_t)y * rowb, src + (size_t)y * src_row_stride, (size_t)rowb);
        } else {
            device_work = true;
        }
    }
    if (!device_work) return 0;

    CtxLease L;
    rc = acquire_ctx(device, &L.c);
    if (rc) return rc;
    Ctx& c = *L.c;
    const int bconst = saturate_u8(border_const);
    const int64_t pitch = wicca_pitch_bytes(W, C);
    WICCA_CUDA(c.d_src.reserve((size_t)pitch * H + 256));
    WICCA_CUDA(cudaEventRecord(c.ev[0], c.stream));
    rc = upload_image_async(c, src, H, rowb, src_row_stride, pitch);
    if (rc) { cudaStreamSynchronize(c.stream); return rc; }
    WICCA_CUDA(cudaEventRecord(c.ev[1], c.stream));
    rc = icons_from_resident(c, H, W, C, pitch, depths, n_depths, border_type, bconst, dsts);
    cudaError_t se = cudaStreamSynchronize(c.stream);
    if (rc || se != cudaSuccess) c.pending.clear();
    if (rc) return rc;
    if (se != cudaSuccess) return cuda_fail(se, "stream synchronize");
    c.flush_pending();
    fill_timing(t, c);
    return 0;
}

int wicca_haar_icon_u8(const uint8_t* src, int H, int W, int C, int64_t src_row_stride, int depth, int border_type,
                       double border_const, uint8_t* dst, int device, wicca_timing* t) {
    uint8_t* dsts[1] = {dst};
    return wicca_haar_icons_multi_u8(src, H, W, C, src_row_stride, &depth, 1, border_type, border_const, dsts, device, t);
}

int wicca_haar_icons_multi_dev(const uint8_t* d_src, int H, int W, int C, int64_t src_pitch, const int* depths,
                               int n_depths, int border_type, double border_const, uint8_t* const* d_dsts,
                               const int64_t* dst_pitches, int device, void* stream_v) {
    int rc = validate_icon_args(d_src, H, W, C, depths, n_depths, border_type);
    if (rc) return rc;
    if (!d_dsts || !dst_pitches) return fail(WICCA_EINVAL, "d_dsts / dst_pitches is NULL");
    if (src_pitch < (int64_t)W * C) return fail(WICCA_EINVAL, "src_pitch < W*C");
    rc = check_device(device);
    if (rc) return rc;
    WICCA_CUDA(cudaSetDevice(device));
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int bconst = saturate_u8(border_const);

    std::vector<IconOut> fo;
    std::vector<int> generic, tail;
    size_t f32_elems = 0;
    const bool can_fuse = fused_eligible(d_src, src_pitch, C);
    for (int i = 0; i < n_depths; ++i) {
        const int d = depths[i];
        if (!d_dsts[i]) return fail(WICCA_EINVAL, "d_dsts[%d] is NULL", i);
        if (d <= 0) return fail(WICCA_EDEPTH, "device-resident call needs depth >= 1");
        const int w = icon_dim(W, d);
        if (dst_pitches[i] < (int64_t)w * C) return fail(WICCA_EINVAL, "dst_pitches[%d] too small", i);
        bool dup = false;
        for (auto& o : fo) dup |= (o.depth == d);
        if (can_fuse && d <= kMaxFused && !dup && ((uintptr_t)d_dsts[i] % 16) == 0 && (dst_pitches[i] % 16) == 0) {
            IconOut o;
            o.depth = d; o.h = icon_dim(H, d); o.w = w; o.d_ptr = d_dsts[i]; o.pitch = dst_pitches[i];
            fo.push_back(o);
        } else if (can_fuse && d > kMaxFused) {
            tail.push_back(i);
            f32_elems = std::max(f32_elems, level8_elems(H, W, C, d));
        } else {
            generic.push_back(i);
            f32_elems = std::max(f32_elems, level8_elems(H, W, C, d));
        }
    }
    // Scratch (descriptor, strip, fp32 ping-pong) is stream-ordered memory: allocated and freed on the
    // caller's stream, so the call only enqueues and never synchronises.
    struct AsyncScratch {
        cudaStream_t stream;
        std::vector<void*> ptrs;
        cudaError_t get(void** p, size_t bytes) {
            cudaError_t e = cudaMallocAsync(p, bytes, stream);
            if (e == cudaSuccess) ptrs.push_back(*p);
            return e;
        }
        ~AsyncScratch() { for (void* p : ptrs) cudaFreeAsync(p, stream); }
    } scratch{stream, {}};
    float* f32a = nullptr;
    float* f32b = nullptr;
    if (f32_elems) {
        WICCA_CUDA(scratch.get((void**)&f32a, f32_elems * sizeof(float)));
        WICCA_CUDA(scratch.get((void**)&f32b, f32_elems * sizeof(float) / 4 + 16));
    }
    if (!fo.empty() || !tail.empty()) {
        uint32_t* sum6 = nullptr;
        int tail_max = 6;
        for (int i : tail) tail_max = std::max(tail_max, depths[i]);
        const TailGeom tg = tail_geom(H, W, tail_max);
        if (!tail.empty()) WICCA_CUDA(scratch.get((void**)&sum6, sum6_bytes(tg)));
        IconImage him;
        rc = fill_icon_image(&him, d_src, H, W, src_pitch, fo.data(), (int)fo.size(), 0, sum6, tg.ext_w);
        if (rc) return rc;
        uint8_t* strip = nullptr;
        if (border_needs_strip(border_type) && strip_px(W, him.Wp_max) > 0)
            WICCA_CUDA(scratch.get((void**)&strip, (size_t)H * kStripPitch));
        uint8_t* d_desc = nullptr;
        WICCA_CUDA(scratch.get((void**)&d_desc, sizeof(IconImage) + 64));
        // pageable sources: the runtime stages them before returning, so the stack copies are safe
        WICCA_CUDA(cudaMemcpyAsync(d_desc, &him, sizeof(IconImage), cudaMemcpyHostToDevice, stream));
        WICCA_CUDA(cudaMemcpyAsync(d_desc + sizeof(IconImage), &strip, sizeof(strip), cudaMemcpyHostToDevice, stream));
        const IconImage* d_im = (const IconImage*)d_desc;
        uint8_t* const* d_strips = (uint8_t* const*)(d_desc + sizeof(IconImage));
        if (strip) {
            cudaError_t e = launch_edge_strips(d_im, d_strips, 1, H, border_base(border_type), bconst, stream);
            if (e != cudaSuccess) return cuda_fail(e, "edge strip kernel");
        }
        cudaError_t e = launch_icon_tma(d_im, d_strips, 1, him.items_x * him.items_y, border_base(border_type), bconst,
                                        device_info(device).sm_count, icon_variant_from_env(),
                                        stream_hint_for(fo.data(), (int)fo.size()), stream);
        if (e != cudaSuccess) return cuda_fail(e, "fused icon kernel");
        if (!tail.empty()) {
            const TailArgs tb = tail_base(d_src, src_pitch, H, W, border_type, bconst, sum6, tg);
            e = launch_icon_tail_fill(tb, stream);
            if (e != cudaSuccess) return cuda_fail(e, "icon tail fill kernel");
            for (int i : tail) {
                rc = enqueue_tail(tb, depths[i], d_dsts[i], dst_pitches[i], f32a, f32b, stream);
                if (rc) return rc;
            }
        }
    }
    for (int i : generic) {
        rc = enqueue_generic(d_src, src_pitch, H, W, C, depths[i], border_type, bconst, d_dsts[i], dst_pitches[i], f32a,
                             f32b, stream);
        if (rc) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// Batch plans
// ------------------------------------------------------------------------------------------
struct wicca_plan {
    int device = 0;
    int n = 0, C = 0, n_depths = 0;
    int border = 1, bconst = 0;
    bool fused = false;
    std::vector<int> depths;
    std::vector<IconOut> outs;                 // n * n_depths
    std::vector<IconImage> h_imgs;
    std::vector<GenericIconArgs> gen;          // general path launches
    std::vector<TailArgs> tails;               // depths 7, 8 of a fused plan, from the level-6 planes
    std::vector<TailArgs> tail_fills;          // one per image whose deeper padding reaches past the depth-6 extents
    DevBuf d_imgs, d_strip_ptrs, d_strips, d_icons, d_sum6;
    int total_items = 0, max_rows = 0, sm_count = 148;
    bool need_strips = false;
    int launches = 0;
    int64_t bytes_read = 0, bytes_written = 0;
    // resize/normalise epilogue: tap tables per (depth index, out_h, out_w), built on first use
    struct Epilogue { DevBuf d_tables; ResizeTableBlob blob; };
    std::map<std::tuple<int, int, int>, Epilogue*> epilogues;
    std::mutex epi_mu;
};

int wicca_plan_create(int device, int n_images, const uint8_t* const* d_srcs, const int* Hs, const int* Ws,
                      const int64_t* src_pitches, int C, const int* depths, int n_depths, int border_type,
                      double border_const, wicca_plan** plan_out) {
    if (!plan_out) return fail(WICCA_EINVAL, "plan is NULL");
    *plan_out = nullptr;
    if (n_images <= 0 || !d_srcs || !Hs || !Ws || !src_pitches) return fail(WICCA_EINVAL, "bad image list");
    if (!depths || n_depths <= 0) return fail(WICCA_EINVAL, "no depths given");
    int rc = check_device(device);
    if (rc) return rc;
    WICCA_CUDA(cudaSetDevice(device));
    bool all_fusable = (C == 3);
    for (int k = 0; k < n_depths; ++k) {
        if (depths[k] < 1 || depths[k] > 8) return fail(WICCA_EDEPTH, "plans support depths 1..8, got %d", depths[k]);
        for (int j = 0; j < k; ++j)
            if (depths[j] == depths[k]) return fail(WICCA_EDEPTH, "duplicate depth %d in plan", depths[k]);
    }
    for (int i = 0; i < n_images; ++i) {
        rc = validate_icon_args(d_srcs[i], Hs[i], Ws[i], C, depths, n_depths, border_type);
        if (rc) return rc;
        if (src_pitches[i] < (int64_t)Ws[i] * C) return fail(WICCA_EINVAL, "src_pitches[%d] < W*C", i);
        if (!fused_eligible(d_srcs[i], src_pitches[i], C)) all_fusable = false;
    }
    wicca_plan* p = new wicca_plan();
    p->device = device; p->n = n_images; p->C = C; p->n_depths = n_depths;
    p->border = border_base(border_type); p->bconst = saturate_u8(border_const);
    p->depths.assign(depths, depths + n_depths);
    p->fused = all_fusable;
    p->sm_count = device_info(device).sm_count;
    p->outs.resize((size_t)n_images * n_depths);

    auto cleanup = [&](int code) {
        p->d_imgs.release(); p->d_strip_ptrs.release(); p->d_strips.release(); p->d_icons.release(); p->d_sum6.release();
        delete p;
        return code;
    };

    size_t icon_bytes = 0;
    for (int i = 0; i < n_images; ++i)
        for (int k = 0; k < n_depths; ++k) {
            IconOut& o = p->outs[(size_t)i * n_depths + k];
            o.depth = depths[k];
            o.h = icon_dim(Hs[i], o.depth);
            o.w = icon_dim(Ws[i], o.depth);
            o.pitch = icon_pitch_for(o.w, C);
            icon_bytes = (size_t)align_up((int64_t)icon_bytes, 256);
            icon_bytes += (size_t)o.pitch * o.h;
            p->bytes_written += (int64_t)o.h * o.w * C;
        }
    for (int i = 0; i < n_images; ++i) p->bytes_read += (int64_t)Hs[i] * Ws[i] * C * (p->fused ? 1 : n_depths);
    cudaError_t e = p->d_icons.reserve(icon_bytes + 256);
    if (e != cudaSuccess) return cleanup(cuda_fail(e, "icon allocation"));
    size_t off = 0;
    for (auto& o : p->outs) {
        off = (size_t)align_up((int64_t)off, 256);
        o.d_ptr = (uint8_t*)p->d_icons.p + off;
        off += (size_t)o.pitch * o.h;
    }

    if (p->fused) {
        p->h_imgs.resize(n_images);
        std::vector<uint8_t*> strip_ptrs(n_images, nullptr);
        size_t strip_bytes = 0;
        int base = 0;
        int tail_max = 6;
        for (int k = 0; k < n_depths; ++k) tail_max = std::max(tail_max, depths[k]);
        const bool deep = tail_max > kMaxFused;
        std::vector<size_t> s6_off(n_images + 1, 0);
        if (deep) {
            for (int i = 0; i < n_images; ++i)
                s6_off[i + 1] = s6_off[i] + (size_t)align_up((int64_t)sum6_bytes(tail_geom(Hs[i], Ws[i], tail_max)), 256);
            e = p->d_sum6.reserve(s6_off[n_images]);
            if (e != cudaSuccess) return cleanup(cuda_fail(e, "level-6 plane allocation"));
        }
        for (int i = 0; i < n_images; ++i) {
            uint32_t* sum6 = deep ? (uint32_t*)((uint8_t*)p->d_sum6.p + s6_off[i]) : nullptr;
            const TailGeom tg = tail_geom(Hs[i], Ws[i], tail_max);
            rc = fill_icon_image(&p->h_imgs[i], d_srcs[i], Hs[i], Ws[i], src_pitches[i], &p->outs[(size_t)i * n_depths],
                                 n_depths, base, sum6, tg.ext_w);
            if (rc) return cleanup(rc);
            if (deep) {
                const TailArgs tb = tail_base(d_srcs[i], src_pitches[i], Hs[i], Ws[i], p->border, p->bconst, sum6, tg);
                if ((int64_t)tg.ext_h * tg.ext_w > (int64_t)tg.s6_h * tg.s6_w) p->tail_fills.push_back(tb);
                for (int k = 0; k < n_depths; ++k) {
                    const IconOut& o = p->outs[(size_t)i * n_depths + k];
                    if (o.depth <= kMaxFused) continue;
                    TailArgs t = tb;
                    t.depth = o.depth; t.out_h = o.h; t.out_w = o.w; t.dst_u8 = o.d_ptr; t.dst_pitch = o.pitch; t.dst_f32 = nullptr;
                    p->tails.push_back(t);
                }
            }
            base += p->h_imgs[i].items_x * p->h_imgs[i].items_y;
            p->max_rows = std::max(p->max_rows, Hs[i]);
            if (border_needs_strip(border_type) && strip_px(Ws[i], p->h_imgs[i].Wp_max) > 0) strip_bytes += (size_t)Hs[i] * kStripPitch;
        }
        p->total_items = base;
        if (strip_bytes) {
            e = p->d_strips.reserve(strip_bytes);
            if (e != cudaSuccess) return cleanup(cuda_fail(e, "strip allocation"));
            size_t so = 0;
            for (int i = 0; i < n_images; ++i)
                if (border_needs_strip(border_type) && strip_px(Ws[i], p->h_imgs[i].Wp_max) > 0) {
                    strip_ptrs[i] = (uint8_t*)p->d_strips.p + so;
                    so += (size_t)Hs[i] * kStripPitch;
                    p->need_strips = true;
                }
        }
        e = p->d_imgs.reserve(sizeof(IconImage) * n_images);
        if (e == cudaSuccess) e = p->d_strip_ptrs.reserve(sizeof(uint8_t*) * n_images);
        if (e == cudaSuccess) e = cudaMemcpy(p->d_imgs.p, p->h_imgs.data(), sizeof(IconImage) * n_images, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(p->d_strip_ptrs.p, strip_ptrs.data(), sizeof(uint8_t*) * n_images, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return cleanup(cuda_fail(e, "plan descriptor upload"));
        p->launches = 1 + (p->need_strips ? 1 : 0) + (int)p->tails.size() + (int)p->tail_fills.size();
    } else {
        for (int i = 0; i < n_images; ++i)
            for (int k = 0; k < n_depths; ++k) {
                const IconOut& o = p->outs[(size_t)i * n_depths + k];
                GenericIconArgs a;
                a.src = d_srcs[i]; a.pitch = src_pitches[i]; a.H = Hs[i]; a.W = Ws[i]; a.C = C;
                a.depth = o.depth; a.border_type = p->border; a.border_const = p->bconst;
                a.out_h = o.h; a.out_w = o.w; a.dst_u8 = o.d_ptr; a.dst_pitch = o.pitch; a.dst_f32 = nullptr;
                p->gen.push_back(a);
            }
        p->launches = (int)p->gen.size();
    }
    *plan_out = p;
    return 0;
}

int wicca_plan_launch(wicca_plan* p, void* stream_v) {
    if (!p) return fail(WICCA_ESTATE, "plan is NULL");
    cudaStream_t stream = (cudaStream_t)stream_v;
    WICCA_CUDA(cudaSetDevice(p->device));
    if (p->fused) {
        const IconImage* d_im = (const IconImage*)p->d_imgs.p;
        uint8_t* const* d_strips = (uint8_t* const*)p->d_strip_ptrs.p;
        if (p->need_strips) {
            cudaError_t e = launch_edge_strips(d_im, d_strips, p->n, p->max_rows, p->border, p->bconst, stream);
            if (e != cudaSuccess) return cuda_fail(e, "edge strip kernel");
        }
        cudaError_t e = launch_icon_tma(d_im, d_strips, p->n, p->total_items, p->border, p->bconst, p->sm_count,
                                        icon_variant_from_env(), stream_hint_for(p->outs.data(), p->n_depths), stream);
        if (e != cudaSuccess) return cuda_fail(e, "fused icon kernel");
        for (const auto& t : p->tail_fills) {
            e = launch_icon_tail_fill(t, stream);
            if (e != cudaSuccess) return cuda_fail(e, "icon tail fill kernel");
        }
        for (const auto& t : p->tails) {
            e = launch_icon_tail(t, stream);
            if (e != cudaSuccess) return cuda_fail(e, "icon tail kernel");
        }
    } else {
        for (const auto& a : p->gen) {
            cudaError_t e = launch_icon_generic(a, stream);
            if (e != cudaSuccess) return cuda_fail(e, "generic icon kernel");
        }
    }
    return 0;
}

int wicca_plan_icon(const wicca_plan* p, int image, int depth_index, uint8_t** d_icon, int* h, int* w, int64_t* pitch) {
    if (!p) return fail(WICCA_ESTATE, "plan is NULL");
    if (image < 0 || image >= p->n || depth_index < 0 || depth_index >= p->n_depths)
        return fail(WICCA_EINVAL, "image/depth index out of range");
    const IconOut& o = p->outs[(size_t)image * p->n_depths + depth_index];
    if (d_icon) *d_icon = o.d_ptr;
    if (h) *h = o.h;
    if (w) *w = o.w;
    if (pitch) *pitch = o.pitch;
    return 0;
}

int wicca_plan_read_icon(const wicca_plan* p, int image, int depth_index, uint8_t* dst) {
    if (!p || !dst) return fail(WICCA_ESTATE, "plan or dst is NULL");
    if (image < 0 || image >= p->n || depth_index < 0 || depth_index >= p->n_depths)
        return fail(WICCA_EINVAL, "image/depth index out of range");
    const IconOut& o = p->outs[(size_t)image * p->n_depths + depth_index];
    WICCA_CUDA(cudaSetDevice(p->device));
    WICCA_CUDA(cudaDeviceSynchronize());
    const size_t rowb = (size_t)o.w * p->C;
    WICCA_CUDA(cudaMemcpy2D(dst, rowb, o.d_ptr, (size_t)o.pitch, rowb, (size_t)o.h, cudaMemcpyDeviceToHost));
    return 0;
}

int wicca_plan_info(const wicca_plan* p, int* launches, int64_t* bytes_read, int64_t* bytes_written) {
    if (!p) return fail(WICCA_ESTATE, "plan is NULL");
    if (launches) *launches = p->launches;
    if (bytes_read) *bytes_read = p->bytes_read;
    if (bytes_written) *bytes_written = p->bytes_written;
    return 0;
}

int wicca_plan_resize_norm(wicca_plan* p, int depth_index, int out_h, int out_w, int norm_mode, float* d_dst,
                           uint8_t* d_dst_u8, void* stream_v) {
    if (!p || !d_dst) return fail(WICCA_ESTATE, "plan or destination is NULL");
    if (depth_index < 0 || depth_index >= p->n_depths) return fail(WICCA_EINVAL, "depth index out of range");
    if (out_h <= 0 || out_w <= 0 || norm_mode < 0 || norm_mode > 3) return fail(WICCA_EINVAL, "bad target size / mode");
    if (p->C != 3) return fail(WICCA_ECHANNELS, "the resize/normalise epilogue needs 3-channel icons");
    WICCA_CUDA(cudaSetDevice(p->device));
    wicca_plan::Epilogue* ep = nullptr;
    {
        std::lock_guard<std::mutex> lk(p->epi_mu);
        auto key = std::make_tuple(depth_index, out_h, out_w);
        auto it = p->epilogues.find(key);
        if (it == p->epilogues.end()) {
            std::vector<ResizeSrc> srcs(p->n);
            for (int i = 0; i < p->n; ++i) {
                const IconOut& o = p->outs[(size_t)i * p->n_depths + depth_index];
                srcs[i] = {o.d_ptr, o.h, o.w, o.pitch};
            }
            ep = new wicca_plan::Epilogue();
            ep->blob = build_resize_tables(srcs, out_h, out_w);
            cudaError_t e = ep->d_tables.reserve(ep->blob.bytes.size());
            if (e == cudaSuccess) e = cudaMemcpy(ep->d_tables.p, ep->blob.bytes.data(), ep->blob.bytes.size(), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { ep->d_tables.release(); delete ep; return cuda_fail(e, "epilogue table upload"); }
            p->epilogues[key] = ep;
        } else {
            ep = it->second;
        }
    }
    cudaError_t e = launch_resize_norm(ep->blob.view(ep->d_tables.p), p->n, out_h, out_w, norm_mode, d_dst, d_dst_u8,
                                       ep->blob.max_src_w, ep->blob.n_area, ep->blob.n_other, (cudaStream_t)stream_v);
    if (e != cudaSuccess) return cuda_fail(e, "resize/normalise kernel");
    return 0;
}

int wicca_plan_destroy(wicca_plan* p) {
    if (!p) return 0;
    cudaSetDevice(p->device);
    for (auto& kv : p->epilogues) { kv.second->d_tables.release(); delete kv.second; }
    p->d_imgs.release(); p->d_strip_ptrs.release(); p->d_strips.release(); p->d_icons.release(); p->d_sum6.release();
    delete p;
    return 0;
}

}  // extern "C"
