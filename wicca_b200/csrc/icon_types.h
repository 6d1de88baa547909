// icon_types.h - descriptors shared by the host side (capi.cu) and the icon kernels.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace wicca {

constexpr int kMaxFused = 6;          // levels produced by the one-pass kernel
constexpr int kItemW = 128;           // pixels per work item (one warp), x
constexpr int kItemH = 64;            // rows per work item = 2^kMaxFused
constexpr int kChunkPx = 16;          // pixels per lane per row
constexpr int kStageRowBytes = kItemW * 3;            // 384
constexpr int kStageBytes = kStageRowBytes * kItemH;  // 24576
// per-warp output tiles of levels 1..3 (dense, handed to TMA store); a warp owns half an item
// (128 x 32 px): L1 16 x 192 B, L2 8 x 96 B, L3 4 x 48 B
constexpr int kOut1Row = 192, kOut2Row = 96, kOut3Row = 48;            // bytes per tile row
constexpr int kHalf1Off = 0, kHalf2Off = 16 * kOut1Row, kHalf3Off = kHalf2Off + 8 * kOut2Row;   // 0, 3072, 3840
constexpr int kHalfStageBytes = 4096;
constexpr int kStripPitch = 256;      // bytes per row of the right-edge strip (<= 78 px * 3)

// One image of a launch.  Lives in global memory (array indexed by image).
struct alignas(128) IconImage {
    CUtensorMap tmap;            // 2-D uint32 view of the pitched image, box 96 x 64
    CUtensorMap hmap[3];         // uint8 views of the level 1..3 icons for TMA store, boxes 192x16, 96x8, 48x4
    const uint8_t* src;          // device, 16-byte aligned
    int64_t pitch;               // bytes, multiple of 16
    int H, W;
    int items_x, items_y;        // ceil(W/128), ceil(H/64)
    int item_base;               // first global work-item index of this image
    int Hp_max, Wp_max;          // padded extents at the deepest emitted level
    uint8_t* icon[kMaxFused];    // per level (index = depth-1); nullptr = not emitted
    int64_t icon_pitch[kMaxFused];
    int icon_h[kMaxFused], icon_w[kMaxFused];
    uint32_t* sum6;              // nullptr, or the plane of exact 64 x 64 block sums for the levels above 6:
    int sum6_h, sum6_w;          // this kernel fills rows < ceil(H/64), columns < ceil(W/64) of it,
    int sum6_stride;             // blocks per plane row (>= sum6_w: the deeper level's padded grid)
};

// Geometry part of an IconImage (everything except the tensor map).  Host side.
inline void icon_image_geometry(IconImage* im, const uint8_t* src, int H, int W, int64_t pitch, int item_base) {
    im->src = src; im->pitch = pitch; im->H = H; im->W = W;
    im->items_x = (W + kItemW - 1) / kItemW;
    im->items_y = (H + kItemH - 1) / kItemH;
    im->item_base = item_base;
    im->Hp_max = 0; im->Wp_max = 0;
    for (int l = 0; l < kMaxFused; ++l) { im->icon[l] = nullptr; im->icon_pitch[l] = 0; im->icon_h[l] = 0; im->icon_w[l] = 0; }
    im->sum6 = nullptr; im->sum6_h = 0; im->sum6_w = 0; im->sum6_stride = 0;
}
// Request the plane of exact level-6 block sums (haar_tail_kernel finishes depths > 6 from it); `stride` blocks per row.
inline void icon_image_add_sum6(IconImage* im, uint32_t* plane, int stride) {
    im->sum6 = plane;
    im->sum6_h = (im->H + 63) >> 6; im->sum6_w = (im->W + 63) >> 6;
    im->sum6_stride = stride < im->sum6_w ? im->sum6_w : stride;
    if (im->sum6_h * 64 > im->Hp_max) im->Hp_max = im->sum6_h * 64;
    if (im->sum6_w * 64 > im->Wp_max) im->Wp_max = im->sum6_w * 64;
}
// Request level `depth` (1..6) of the image; icon rows are `pitch` bytes apart.
inline void icon_image_add_level(IconImage* im, int depth, uint8_t* icon, int64_t pitch) {
    const int r = 1 << depth;
    const int h = (im->H + r - 1) >> depth, w = (im->W + r - 1) >> depth;
    im->icon[depth - 1] = icon; im->icon_pitch[depth - 1] = pitch;
    im->icon_h[depth - 1] = h; im->icon_w[depth - 1] = w;
    if (h * r > im->Hp_max) im->Hp_max = h * r;     // padded extents grow with depth
    if (w * r > im->Wp_max) im->Wp_max = w * r;
}

// Arguments of the general (any C / depth <= 8 / any alignment) kernel.
struct GenericIconArgs {
    const uint8_t* src; int64_t pitch; int H, W, C;
    int depth;                   // 1..8
    int border_type, border_const;
    int out_h, out_w;
    uint8_t* dst_u8; int64_t dst_pitch;      // used when depth is the final depth
    float* dst_f32;                          // else: exact level value as fp32 (tight, out_w*C per row)
};

// Arguments of haar_tail_kernel: depth 7 / 8 (uint8) or the exact level-8 plane (float32) of a 3-channel image from
// the level-6 sums; out_h x out_w may exceed the depth's own icon when the plane feeds still deeper levels.
struct TailArgs {
    const uint8_t* src; int64_t pitch; int H, W;
    int border_type, border_const;
    uint32_t* sum6;              // (ext_h, ext_w, 3): rows < s6_h, columns < s6_w come from the one-pass kernel,
    int s6_h, s6_w, ext_h, ext_w; //                   the rest from haar_tail_fill_kernel
    int depth;                   // 7 or 8
    int out_h, out_w;
    uint8_t* dst_u8; int64_t dst_pitch;
    float* dst_f32;
};

}  // namespace wicca
