// haar_subband.cu - full sub-band forward / inverse 2-D Haar transform (extension, SURVEY.md
// 8(a) row A4; the reference keeps only LL, wicca/wavelet_coder.py:61-65).
//
// Per level, with the 2x2 block  a b / c d  (a = x[2i,2j], b = x[2i,2j+1], c = x[2i+1,2j]):
//   LL = ((a+c)+(b+d))/4   HL = ((a+c)-(b+d))/4   LH = ((a-c)+(b-d))/4   HH = ((a-c)-(b-d))/4
// i.e. row-pair sum/difference first, then column-pair sum/difference, then *0.25 - the order
// of the reference's LL.  Coefficients are stored Mallat-style in one fp32 HWC plane of the
// padded size: HL_l right of LL_l, LH_l below, HH_l diagonal.
//
// Two sets of kernels: per-level ones (any channel count, any depth; one thread per coefficient
// quadruple) and the fused patch kernels below for C <= 4, which do the first four levels in one pass.
#include <cuda_runtime.h>
#include <stdint.h>

#include "haar_math.cuh"
#include "kernels.h"

namespace wicca {

struct SubbandOut {
    float* ll; int64_t ll_stride;      // LL_l destination (row stride in floats)
    float* plane; int64_t pl_stride;   // Mallat plane (row stride in floats)
    int h, w, C;                       // extent of the level's sub-bands
};

__device__ __forceinline__ void analyse_store(const SubbandOut& o, int oy, int ox, int c, float a, float b, float cc,
                                              float d) {
    const float rs0 = __fadd_rn(a, cc), rs1 = __fadd_rn(b, d);     // row-pair sums   (even col, odd col)
    const float rd0 = __fsub_rn(a, cc), rd1 = __fsub_rn(b, d);     // row-pair diffs
    const float ll = __fmul_rn(__fadd_rn(rs0, rs1), 0.25f);
    const float hl = __fmul_rn(__fsub_rn(rs0, rs1), 0.25f);
    const float lh = __fmul_rn(__fadd_rn(rd0, rd1), 0.25f);
    const float hh = __fmul_rn(__fsub_rn(rd0, rd1), 0.25f);
    const int64_t e = (int64_t)ox * o.C + c;
    o.ll[(int64_t)oy * o.ll_stride + e] = ll;
    o.plane[(int64_t)oy * o.pl_stride + (int64_t)o.w * o.C + e] = hl;                      // right
    o.plane[(int64_t)(oy + o.h) * o.pl_stride + e] = lh;                                    // below
    o.plane[(int64_t)(oy + o.h) * o.pl_stride + (int64_t)o.w * o.C + e] = hh;               // diagonal
}

// Level 1 straight from the uint8 image, border rule evaluated on the fly (no padded copy).
__global__ void forward_level1_u8_kernel(const uint8_t* __restrict__ src, int64_t pitch, int H, int W,
                                         int border_type, int border_const, SubbandOut o) {
    const int64_t n = (int64_t)o.h * o.w * o.C;
    const float fc = (float)border_const;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % o.C);
        const int64_t t = i / o.C;
        const int ox = (int)(t % o.w);
        const int oy = (int)(t / o.w);
        const int y0 = border_index(2 * oy, H, border_type), y1 = border_index(2 * oy + 1, H, border_type);
        const int x0 = border_index(2 * ox, W, border_type), x1 = border_index(2 * ox + 1, W, border_type);
        auto px = [&](int y, int x) -> float {
            return (y < 0 || x < 0) ? fc : (float)src[(int64_t)y * pitch + (int64_t)x * o.C + c];
        };
        analyse_store(o, oy, ox, c, px(y0, x0), px(y0, x1), px(y1, x0), px(y1, x1));
    }
}

// Levels >= 2: input is the previous LL (fp32, row stride in floats).
__global__ void forward_level_f32_kernel(const float* __restrict__ in, int64_t in_stride, SubbandOut o) {
    const int64_t n = (int64_t)o.h * o.w * o.C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % o.C);
        const int64_t t = i / o.C;
        const int ox = (int)(t % o.w);
        const int oy = (int)(t / o.w);
        const float* p = in + (int64_t)(2 * oy) * in_stride + (int64_t)(2 * ox) * o.C + c;
        analyse_store(o, oy, ox, c, p[0], p[o.C], p[in_stride], p[in_stride + o.C]);
    }
}

// One synthesis level: LL_l (h x w) + details of level l from the plane -> LL_{l-1} (2h x 2w).
__global__ void inverse_level_f32_kernel(const float* __restrict__ ll, int64_t ll_stride,
                                         const float* __restrict__ plane, int64_t pl_stride, float* __restrict__ out,
                                         int64_t out_stride, int h, int w, int C) {
    const int64_t n = (int64_t)h * w * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t t = i / C;
        const int ox = (int)(t % w);
        const int oy = (int)(t / w);
        const int64_t e = (int64_t)ox * C + c;
        const float vll = ll[(int64_t)oy * ll_stride + e];
        const float vhl = plane[(int64_t)oy * pl_stride + (int64_t)w * C + e];
        const float vlh = plane[(int64_t)(oy + h) * pl_stride + e];
        const float vhh = plane[(int64_t)(oy + h) * pl_stride + (int64_t)w * C + e];
        const float s0 = __fadd_rn(vll, vhl), s1 = __fsub_rn(vll, vhl);    // column sums of the block
        const float d0 = __fadd_rn(vlh, vhh), d1 = __fsub_rn(vlh, vhh);
        float* q = out + (int64_t)(2 * oy) * out_stride + (int64_t)(2 * ox) * C + c;
        q[0] = __fadd_rn(s0, d0);                 // a = LL+HL+LH+HH
        q[C] = __fadd_rn(s1, d1);                 // b = LL-HL+LH-HH
        q[out_stride] = __fsub_rn(s0, d0);        // c = LL+HL-LH-HH
        q[out_stride + C] = __fsub_rn(s1, d1);    // d = LL-HL-LH+HH
    }
}

// Two levels per launch for the levels above the fused ones (a plane at least 256 times smaller than the image, where
// a launch costs more than its work): a thread takes one (block of the upper level, channel) = 4 x 4 values of the
// input LL.  o1: the lower of the two levels (its LL is not stored, nobody reads it), o2: the upper one.
__global__ void forward_level2x_f32_kernel(const float* __restrict__ in, int64_t in_stride, SubbandOut o1, SubbandOut o2) {
    const int64_t n = (int64_t)o2.h * o2.w * o2.C;
    const int C = o2.C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t t = i / C;
        const int ox = (int)(t % o2.w);
        const int oy = (int)(t / o2.w);
        float ll1[2][2];
#pragma unroll
        for (int iy = 0; iy < 2; ++iy)
#pragma unroll
            for (int ix = 0; ix < 2; ++ix) {
                const int by = 2 * oy + iy, bx = 2 * ox + ix;                   // block of the lower level
                const float* p = in + (int64_t)(2 * by) * in_stride + (int64_t)(2 * bx) * C + c;
                const float a = p[0], b = p[C], cc = p[in_stride], d = p[in_stride + C];
                const float rs0 = __fadd_rn(a, cc), rs1 = __fadd_rn(b, d), rd0 = __fsub_rn(a, cc), rd1 = __fsub_rn(b, d);
                ll1[iy][ix] = __fmul_rn(__fadd_rn(rs0, rs1), 0.25f);
                const int64_t e = (int64_t)bx * C + c;
                o1.plane[(int64_t)by * o1.pl_stride + (int64_t)o1.w * C + e] = __fmul_rn(__fsub_rn(rs0, rs1), 0.25f);
                o1.plane[(int64_t)(by + o1.h) * o1.pl_stride + e] = __fmul_rn(__fadd_rn(rd0, rd1), 0.25f);
                o1.plane[(int64_t)(by + o1.h) * o1.pl_stride + (int64_t)o1.w * C + e] = __fmul_rn(__fsub_rn(rd0, rd1), 0.25f);
            }
        analyse_store(o2, oy, ox, c, ll1[0][0], ll1[0][1], ll1[1][0], ll1[1][1]);
    }
}

// Two synthesis levels per launch: LL_l (h x w) + details of levels l and l - 1 -> LL_{l-2} (4h x 4w).
__global__ void inverse_level2x_f32_kernel(const float* __restrict__ ll, int64_t ll_stride, const float* __restrict__ plane,
                                           int64_t pl_stride, float* __restrict__ out, int64_t out_stride, int h, int w, int C) {
    const int64_t n = (int64_t)h * w * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t t = i / C;
        const int ox = (int)(t % w);
        const int oy = (int)(t / w);
        float ll1[2][2];
        {
            const int64_t e = (int64_t)ox * C + c;
            const float vll = ll[(int64_t)oy * ll_stride + e];
            const float vhl = plane[(int64_t)oy * pl_stride + (int64_t)w * C + e];
            const float vlh = plane[(int64_t)(oy + h) * pl_stride + e];
            const float vhh = plane[(int64_t)(oy + h) * pl_stride + (int64_t)w * C + e];
            const float s0 = __fadd_rn(vll, vhl), s1 = __fsub_rn(vll, vhl), d0 = __fadd_rn(vlh, vhh), d1 = __fsub_rn(vlh, vhh);
            ll1[0][0] = __fadd_rn(s0, d0); ll1[0][1] = __fadd_rn(s1, d1); ll1[1][0] = __fsub_rn(s0, d0); ll1[1][1] = __fsub_rn(s1, d1);
        }
        const int h1 = 2 * h, w1 = 2 * w;
#pragma unroll
        for (int iy = 0; iy < 2; ++iy)
#pragma unroll
            for (int ix = 0; ix < 2; ++ix) {
                const int by = 2 * oy + iy, bx = 2 * ox + ix;
                const int64_t e = (int64_t)bx * C + c;
                const float vll = ll1[iy][ix];
                const float vhl = plane[(int64_t)by * pl_stride + (int64_t)w1 * C + e];
                const float vlh = plane[(int64_t)(by + h1) * pl_stride + e];
                const float vhh = plane[(int64_t)(by + h1) * pl_stride + (int64_t)w1 * C + e];
                const float s0 = __fadd_rn(vll, vhl), s1 = __fsub_rn(vll, vhl), d0 = __fadd_rn(vlh, vhh), d1 = __fsub_rn(vlh, vhh);
                float* q = out + (int64_t)(2 * by) * out_stride + (int64_t)(2 * bx) * C + c;
                q[0] = __fadd_rn(s0, d0); q[C] = __fadd_rn(s1, d1); q[out_stride] = __fsub_rn(s0, d0); q[out_stride + C] = __fsub_rn(s1, d1);
            }
    }
}

// ------------------------------------------------------------------------------------------
// Fused kernels: levels 1..min(depth, 4) of one 64 x 64-pixel tile in one pass (the per-level
// kernels above finish depths > 4 on a plane that is 256 times smaller).  Forward: the uint8 tile is
// read once (C B/px), every coefficient is written once (4C B/px).  Inverse: every coefficient is
// read once, the image is written once.  The channel count is a template parameter.
// ------------------------------------------------------------------------------------------
constexpr int kTile = 64;
constexpr int kTileThreads = 256;
constexpr int kFusedLevels = 4;          // inverse: levels expanded inside a tile
constexpr int kFwdFusedLevels = 6;       // forward: a 64 x 64 tile holds one whole level-6 block, so the tile finishes it

struct TileGeom {
    int Hp, Wp;                 // padded extents (multiples of 2^depth)
    int levels;                 // levels done inside the tile: min(depth, kFusedLevels)
    int tiles_x, tiles_y;
    float* plane; int64_t pl_stride;      // Mallat plane
    float* ll; int64_t ll_stride;         // where LL_levels lives (the plane itself when depth <= kFusedLevels)
};

template <int C>
__device__ __forceinline__ void analyse(const float (&a)[C], const float (&b)[C], const float (&cc)[C],
                                        const float (&d)[C], float (&ll)[C], float (&hl)[C], float (&lh)[C],
                                        float (&hh)[C]) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float rs0 = __fadd_rn(a[c], cc[c]), rs1 = __fadd_rn(b[c], d[c]);
        const float rd0 = __fsub_rn(a[c], cc[c]), rd1 = __fsub_rn(b[c], d[c]);
        ll[c] = __fmul_rn(__fadd_rn(rs0, rs1), 0.25f);
        hl[c] = __fmul_rn(__fsub_rn(rs0, rs1), 0.25f);
        lh[c] = __fmul_rn(__fadd_rn(rd0, rd1), 0.25f);
        hh[c] = __fmul_rn(__fsub_rn(rd0, rd1), 0.25f);
    }
}

// ------------------------------------------------------------------------------------------
// Forward: a CTA covers a 64 x 64-pixel tile, a warp a 16-row x 32-pixel region of it, and a lane
// one 4 x 4-pixel patch (one level-2 block; lane = 8*ly + lx inside the warp's 4 x 8 level-2 blocks).
// Levels 1 and 2 are computed in registers; for levels 3 and 4 the warp's level-2 LLs pass through the per-warp
// stage and a lane owns one (block, channel) element (see below).  Up to level 4 nothing crosses a warp: no barrier,
// no CTA-wide pyramid.  Depths 5 and 6 end with ONE barrier, after which warp 0 finishes the tile from its sixteen
// level-4 LLs (all d levels of depth <= 6 in one pass over the image, one launch).  Sub-band rows leave through the per-warp stage so that every store
// instruction writes 128 contiguous bytes of one row.
// ------------------------------------------------------------------------------------------
// Store ROWS staged rows of SEG floats (row pitch PITCH in the stage) to rows of a sub-band.  One store
// instruction covers 32 consecutive floats of one row (or 32/SEG whole rows when SEG is 8 or 16); the row pointer
// is bumped by the stride, every other offset is an immediate.
template <int SEG, int ROWS, int PITCH, bool kCheck>
__device__ __forceinline__ void store_rows_impl(const float* stage, int lane, float* base, int64_t stride, int rows_valid,
                                                int seg_valid) {
    constexpr int RPI = (SEG == 8 || SEG == 16) ? 32 / SEG : 1;       // rows per store instruction
    constexpr int IPR = (SEG + 31) / 32;                               // store instructions per row (RPI == 1)
    const int sub = RPI > 1 ? lane / SEG : 0, j = RPI > 1 ? lane % SEG : lane;
    float* p = base + (int64_t)sub * stride + j;
    const float* q = stage + sub * PITCH + j;
#pragma unroll
    for (int r = 0; r < ROWS; r += RPI) {
#pragma unroll
        for (int i = 0; i < IPR; ++i) {
            bool ok = (SEG % 32 == 0 || RPI > 1) ? true : (j + 32 * i < SEG);
            if (kCheck) ok = ok && (r + sub < rows_valid) && (j + 32 * i < seg_valid);
            if (ok) p[32 * i] = q[r * PITCH + 32 * i];
        }
        p += RPI * stride;
    }
}
template <int SEG, int ROWS, int PITCH>
__device__ __forceinline__ void store_rows(const float* stage, int lane, float* base, int64_t stride, bool full,
                                           int rows_valid, int seg_valid) {
    if (full) store_rows_impl<SEG, ROWS, PITCH, false>(stage, lane, base, stride, rows_valid, seg_valid);
    else store_rows_impl<SEG, ROWS, PITCH, true>(stage, lane, base, stride, rows_valid, seg_valid);
}

// Store that asks L2 to keep the line for a while (evict_last): the level-3/4 sub-band rows are written as 48- and
// 24-byte runs by different warps and CTAs, and should meet their neighbours in L2 before they travel to HBM.
__device__ __forceinline__ void store_keep(float* p, float v) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

template <int C>
__global__ void __launch_bounds__(kTileThreads, C <= 3 ? 4 : 3)
forward_patch_kernel(const uint8_t* __restrict__ src, int64_t pitch, int H, int W, int border_type, int border_const,
                     TileGeom g) {
    constexpr int kPitch1 = 16 * C + 8;                        // staged level-1 row (+8: float2 writes hit every bank once)
    __shared__ __align__(16) float s_st[kTileThreads / 32][8 * kPitch1];   // per-warp staging of one sub-band of one level
    __shared__ float s_ll4[4][4 * C];                                      // the tile's level-4 LLs (levels 5 and 6)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy = warp >> 1, wx = warp & 1, ly = lane >> 3, lx = lane & 7;
    const int ty = blockIdx.y, tx = blockIdx.x;
    const int b2y = 4 * wy + ly, b2x = 8 * wx + lx;                       // level-2 block inside the tile
    const int py = ty * kTile + 4 * b2y, px = tx * kTile + 4 * b2x;       // top-left pixel of the lane's patch
    const bool interior = (ty * kTile + kTile <= H) && (tx * kTile + kTile <= W);
    const bool full = (ty * kTile + kTile <= g.Hp) && (tx * kTile + kTile <= g.Wp);   // no clipped sub-band rows
    const bool wide = ((uintptr_t)src % 4 == 0) && (pitch % 4 == 0);
    const float fc = (float)border_const;
    float* stage = s_st[warp];
    const int64_t stride = g.pl_stride;

    // ---- the 4 x 4 patch as floats
    float pix[4][4][C];
    if (interior && wide) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(src + (int64_t)(py + r) * pitch + (int64_t)px * C);
            uint32_t wds[C];
#pragma unroll
            for (int k = 0; k < C; ++k) wds[k] = p[k];
#pragma unroll
            for (int b = 0; b < 4 * C; ++b) pix[r][b / C][b % C] = (float)((wds[b >> 2] >> (8 * (b & 3))) & 0xFFu);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int ym = border_index(py + r, H, border_type);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int xm = border_index(px + q, W, border_type);
#pragma unroll
                for (int c = 0; c < C; ++c)
                    pix[r][q][c] = (ym < 0 || xm < 0) ? fc : (float)src[(int64_t)ym * pitch + (int64_t)xm * C + c];
            }
        }
    }
    // ---- level 1: four 2 x 2 blocks per lane
    float ll1[2][2][C], hl1[2][2][C], lh1[2][2][C], hh1[2][2][C];
#pragma unroll
    for (int iy = 0; iy < 2; ++iy)
#pragma unroll
        for (int ix = 0; ix < 2; ++ix)
            analyse<C>(pix[2 * iy][2 * ix], pix[2 * iy][2 * ix + 1], pix[2 * iy + 1][2 * ix], pix[2 * iy + 1][2 * ix + 1],
                       ll1[iy][ix], hl1[iy][ix], lh1[iy][ix], hh1[iy][ix]);
    {
        const int h1 = g.Hp >> 1, w1 = g.Wp >> 1;
        const int gy0 = ty * 32 + 8 * wy, gx0 = tx * 32 + 16 * wx;       // first level-1 block row / column of the warp
        const int rows_valid = h1 - gy0, seg_valid = (w1 - gx0) * C;
        float* base_hl = g.plane + (int64_t)gy0 * g.pl_stride + (int64_t)(w1 + gx0) * C;
        float* base_lh = g.plane + (int64_t)(gy0 + h1) * g.pl_stride + (int64_t)gx0 * C;
        float* base_hh = base_lh + (int64_t)w1 * C;
        auto put = [&](const float (&v)[2][2][C]) {
#pragma unroll
            for (int iy = 0; iy < 2; ++iy) {
                float2* q = reinterpret_cast<float2*>(stage + (2 * ly + iy) * kPitch1 + 2 * lx * C);
#pragma unroll
                for (int e = 0; e < C; ++e)
                    q[e] = make_float2(v[iy][(2 * e) / C][(2 * e) % C], v[iy][(2 * e + 1) / C][(2 * e + 1) % C]);
            }
            __syncwarp();
        };
        put(hl1); store_rows<16 * C, 8, kPitch1>(stage, lane, base_hl, stride, full, rows_valid, seg_valid); __syncwarp();
        put(lh1); store_rows<16 * C, 8, kPitch1>(stage, lane, base_lh, stride, full, rows_valid, seg_valid); __syncwarp();
        put(hh1); store_rows<16 * C, 8, kPitch1>(stage, lane, base_hh, stride, full, rows_valid, seg_valid); __syncwarp();
        if (g.levels == 1) {
            put(ll1);
            store_rows<16 * C, 8, kPitch1>(stage, lane, g.ll + (int64_t)gy0 * g.ll_stride + (int64_t)gx0 * C,
                                               g.ll_stride, full, rows_valid, seg_valid);
            return;
        }
    }
    // ---- level 2: one block per lane, straight from the level-1 LLs in registers
    float ll2[C], hl2[C], lh2[C], hh2[C];
    analyse<C>(ll1[0][0], ll1[0][1], ll1[1][0], ll1[1][1], ll2, hl2, lh2, hh2);
    {
        const int h2 = g.Hp >> 2, w2 = g.Wp >> 2;
        const int gy0 = ty * 16 + 4 * wy, gx0 = tx * 16 + 8 * wx;
        const int rows_valid = h2 - gy0, seg_valid = (w2 - gx0) * C;
        constexpr int seg = 8 * C;                               // 4 rows of 8C floats = C store instructions
        float* base_hl = g.plane + (int64_t)gy0 * g.pl_stride + (int64_t)(w2 + gx0) * C;
        float* base_lh = g.plane + (int64_t)(gy0 + h2) * g.pl_stride + (int64_t)gx0 * C;
        float* base_hh = base_lh + (int64_t)w2 * C;
        auto put = [&](const float (&v)[C]) {
#pragma unroll
            for (int c = 0; c < C; ++c) stage[ly * seg + lx * C + c] = v[c];
            __syncwarp();
        };
        put(hl2); store_rows<seg, 4, seg>(stage, lane, base_hl, stride, full, rows_valid, seg_valid); __syncwarp();
        put(lh2); store_rows<seg, 4, seg>(stage, lane, base_lh, stride, full, rows_valid, seg_valid); __syncwarp();
        put(hh2); store_rows<seg, 4, seg>(stage, lane, base_hh, stride, full, rows_valid, seg_valid); __syncwarp();
        if (g.levels == 2) {
            put(ll2);
            store_rows<seg, 4, seg>(stage, lane, g.ll + (int64_t)gy0 * g.ll_stride + (int64_t)gx0 * C, g.ll_stride,
                                          full, rows_valid, seg_valid);
            return;
        }
    }
    // ---- levels 3 and 4: the level-2 LLs of the warp's 4 x 8 blocks go through the stage once, then a lane owns one
    // (level-3 block, channel) element, in the order of the sub-band rows: lane = row * 4C + block * C + channel.  Its
    // coefficients go straight from registers to 4C-float runs of two rows (no per-lane scalar stores: those cost 6 x
    // the L2 write transactions, and no redundant arithmetic on all 32 lanes: the kernel is close to issue bound).
    {
        constexpr int seg2 = 8 * C, seg = 4 * C;                  // floats of a level-2 / level-3 row of the warp
#pragma unroll
        for (int c = 0; c < C; ++c) stage[ly * seg2 + lx * C + c] = ll2[c];
        __syncwarp();
        const int row = lane / seg, j = lane % seg;               // lanes >= 2 * seg idle (C < 4)
        float ll3 = 0.f;
        if (lane < 2 * seg) {
            const float* p = stage + (2 * row) * seg2 + 2 * (j / C) * C + (j % C);
            const float a = p[0], b = p[C], cc = p[seg2], d = p[seg2 + C];
            const float rs0 = __fadd_rn(a, cc), rs1 = __fadd_rn(b, d), rd0 = __fsub_rn(a, cc), rd1 = __fsub_rn(b, d);
            ll3 = __fmul_rn(__fadd_rn(rs0, rs1), 0.25f);
            const int h3 = g.Hp >> 3, w3 = g.Wp >> 3;
            const int gy = ty * 8 + 2 * wy + row, gx0 = tx * 8 + 4 * wx;
            if (gy < h3 && gx0 + j / C < w3) {
                float* q_hl = g.plane + (int64_t)gy * g.pl_stride + (int64_t)(w3 + gx0) * C + j;
                float* q_lh = g.plane + (int64_t)(gy + h3) * g.pl_stride + (int64_t)gx0 * C + j;
                store_keep(q_hl, __fmul_rn(__fsub_rn(rs0, rs1), 0.25f));
                store_keep(q_lh, __fmul_rn(__fadd_rn(rd0, rd1), 0.25f));
                store_keep(q_lh + (int64_t)w3 * C, __fmul_rn(__fsub_rn(rd0, rd1), 0.25f));
                if (g.levels == 3) store_keep(g.ll + (int64_t)gy * g.ll_stride + (int64_t)gx0 * C + j, ll3);
            }
        }
        if (g.levels == 3) return;
        // ---- level 4: the warp's 1 x 2 blocks; lane = block * C + channel takes its four LL_3 from the lanes above
        const int t = lane % (2 * C);
        const int s0 = (t / C) * 2 * C + (t % C);                 // lane holding LL_3 of (row 0, block 2 * (t / C), channel)
        const float a = __shfl_sync(0xFFFFFFFFu, ll3, s0), b = __shfl_sync(0xFFFFFFFFu, ll3, s0 + C);
        const float cc = __shfl_sync(0xFFFFFFFFu, ll3, s0 + seg), d = __shfl_sync(0xFFFFFFFFu, ll3, s0 + seg + C);
        if (lane < 2 * C) {
            const float rs0 = __fadd_rn(a, cc), rs1 = __fadd_rn(b, d), rd0 = __fsub_rn(a, cc), rd1 = __fsub_rn(b, d);
            const float ll4 = __fmul_rn(__fadd_rn(rs0, rs1), 0.25f);
            const int h4 = g.Hp >> 4, w4 = g.Wp >> 4;
            const int gy = ty * 4 + wy, gx0 = tx * 4 + 2 * wx;
            if (gy < h4 && gx0 + lane / C < w4) {
                float* q_hl = g.plane + (int64_t)gy * g.pl_stride + (int64_t)(w4 + gx0) * C + lane;
                float* q_lh = g.plane + (int64_t)(gy + h4) * g.pl_stride + (int64_t)gx0 * C + lane;
                store_keep(q_hl, __fmul_rn(__fsub_rn(rs0, rs1), 0.25f));
                store_keep(q_lh, __fmul_rn(__fadd_rn(rd0, rd1), 0.25f));
                store_keep(q_lh + (int64_t)w4 * C, __fmul_rn(__fsub_rn(rd0, rd1), 0.25f));
                if (g.levels == 4) store_keep(g.ll + (int64_t)gy * g.ll_stride + (int64_t)gx0 * C + lane, ll4);
            }
            s_ll4[wy][2 * wx * C + lane] = ll4;                   // lane = block * C + channel
        }
    }
    if (g.levels == 4) return;
    // ---- levels 5 and 6: the tile's 4 x 4 level-4 LLs meet in shared memory - the only CTA-wide step, a few dozen
    // values - and warp 0 finishes the tile's 2 x 2 level-5 blocks and its one level-6 block (a 64 x 64 tile IS a
    // level-6 block), so depths 5 and 6 need no second launch over a scratch plane.
    __syncthreads();
    if (warp != 0) return;
    float ll5 = 0.f;
    {
        const int blk = (lane / C) & 3, c = lane % C, by = blk >> 1, bx = blk & 1;
        const float a = s_ll4[2 * by][(2 * bx) * C + c], b = s_ll4[2 * by][(2 * bx + 1) * C + c];
        const float cc = s_ll4[2 * by + 1][(2 * bx) * C + c], d = s_ll4[2 * by + 1][(2 * bx + 1) * C + c];
        const float rs0 = __fadd_rn(a, cc), rs1 = __fadd_rn(b, d), rd0 = __fsub_rn(a, cc), rd1 = __fsub_rn(b, d);
        ll5 = __fmul_rn(__fadd_rn(rs0, rs1), 0.25f);
        const int h5 = g.Hp >> 5, w5 = g.Wp >> 5;
        const int gy = ty * 2 + by, gx = tx * 2 + bx;
        if (lane < 4 * C && gy < h5 && gx < w5) {
            float* q_hl = g.plane + (int64_t)gy * g.pl_stride + (int64_t)(w5 + gx) * C + c;
            float* q_lh = g.plane + (int64_t)(gy + h5) * g.pl_stride + (int64_t)gx * C + c;
            store_keep(q_hl, __fmul_rn(__fsub_rn(rs0, rs1), 0.25f));
            store_keep(q_lh, __fmul_rn(__fadd_rn(rd0, rd1), 0.25f));
            store_keep(q_lh + (int64_t)w5 * C, __fmul_rn(__fsub_rn(rd0, rd1), 0.25f));
            if (g.levels == 5) store_keep(g.ll + (int64_t)gy * g.ll_stride + (int64_t)gx * C + c, ll5);
        }
    }
    if (g.levels == 5) return;
    {
        const int c = lane % C;
        const float a = __shfl_sync(0xFFFFFFFFu, ll5, c), b = __shfl_sync(0xFFFFFFFFu, ll5, C + c);
        const float cc = __shfl_sync(0xFFFFFFFFu, ll5, 2 * C + c), d = __shfl_sync(0xFFFFFFFFu, ll5, 3 * C + c);
        const float rs0 = __fadd_rn(a, cc), rs1 = __fadd_rn(b, d), rd0 = __fsub_rn(a, cc), rd1 = __fsub_rn(b, d);
        const int h6 = g.Hp >> 6, w6 = g.Wp >> 6;
        if (lane < C && ty < h6 && tx < w6) {
            float* q_hl = g.plane + (int64_t)ty * g.pl_stride + (int64_t)(w6 + tx) * C + c;
            float* q_lh = g.plane + (int64_t)(ty + h6) * g.pl_stride + (int64_t)tx * C + c;
            store_keep(q_hl, __fmul_rn(__fsub_rn(rs0, rs1), 0.25f));
            store_keep(q_lh, __fmul_rn(__fadd_rn(rd0, rd1), 0.25f));
            store_keep(q_lh + (int64_t)w6 * C, __fmul_rn(__fsub_rn(rd0, rd1), 0.25f));
            store_keep(g.ll + (int64_t)ty * g.ll_stride + (int64_t)tx * C + c, __fmul_rn(__fadd_rn(rs0, rs1), 0.25f));   // levels == 6
        }
    }
}

// ------------------------------------------------------------------------------------------
// Inverse, the mirror image: a lane owns one level-2 block = a 4 x 4 patch of the output.  The LL of that block is
// walked down from the top fused level inside the warp (levels 4 and 3: one (block, channel) element per lane,
// handed down through the per-warp stage), while the level-1 details - most of the input - are already on their
// way to shared memory (cp.async issued first thing); every global load of the walk is issued before the first
// wait, because a second serial trip to HBM per warp is what this kernel cannot afford.  Levels 2 and 1 are
// expanded in registers and the 4 x 4 x C patch leaves through the per-warp stage as whole rows.  No barrier.
// ------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void load_run(const float* p, bool vec, bool inside, float (&v)[N]) {
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = 0.f;
    if (!inside) return;
    if (vec && N % 2 == 0) {
#pragma unroll
        for (int k = 0; k < N / 2; ++k) { const float2 t = reinterpret_cast<const float2*>(p)[k]; v[2 * k] = t.x; v[2 * k + 1] = t.y; }
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] = p[k];
    }
}

// Shared memory of the inverse kernel, per warp: the level-1 details of its region (3 sub-bands x 8 rows x 16C floats;
// +8 floats of row pitch so that the lanes' float2 reads - 2C floats per lane, rows 2 ly + iy - hit every bank once)
// followed by the 4-row output stage.
template <int C>
struct InvSmem {
    static constexpr int kDetPitch = 16 * C + 8;
    static constexpr int kDetRows = 24;
    static constexpr int kWarpFloats = kDetRows * kDetPitch + 4 * 32 * C;
    static constexpr size_t kBytes = (size_t)(kTileThreads / 32) * kWarpFloats * sizeof(float);
};

// cp.async of the 24 detail row segments (HL, LH, HH x 8 rows, 64C bytes each) in chunks of CB bytes.
template <int C, int CB>
__device__ __forceinline__ void fetch_details(float* det, const float* b_hl, const float* b_lh, int64_t hh_off, int64_t stride,
                                              int lane) {
    constexpr int kChunksPerRow = 64 * C / CB, kPerLane = 24 * kChunksPerRow / 32, kF = CB / 4;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
        const int q = lane + 32 * i, seg = q / kChunksPerRow, ch = q % kChunksPerRow;
        const int sb = seg >> 3, r = seg & 7;
        const float* src = (sb == 0 ? b_hl : b_lh) + (int64_t)r * stride + (sb == 2 ? hh_off : 0) + ch * kF;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(det + seg * InvSmem<C>::kDetPitch + ch * kF);
        if (CB == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        else if (CB == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
        else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
    }
}

template <int C>
__global__ void __launch_bounds__(kTileThreads, C <= 3 ? 4 : 3)
inverse_patch_kernel(TileGeom g, float* __restrict__ out, int64_t out_stride) {
    constexpr int kSeg = 32 * C;                                  // floats of one output row of the warp's region
    constexpr int kDetPitch = InvSmem<C>::kDetPitch, kDetRows = InvSmem<C>::kDetRows;
    extern __shared__ __align__(16) float s_inv[];                // per warp: [level-1 details][output stage]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy = warp >> 1, wx = warp & 1, ly = lane >> 3, lx = lane & 7;
    const int ty = blockIdx.y, tx = blockIdx.x;
    const int gy2 = ty * 16 + 4 * wy + ly, gx2 = tx * 16 + 8 * wx + lx;      // the lane's level-2 block (global)
    const int L = g.levels;
    const int h1 = g.Hp >> 1, w1 = g.Wp >> 1;
    const bool full = (ty * kTile + kTile <= g.Hp) && (tx * kTile + kTile <= g.Wp);
    float* det = s_inv + warp * InvSmem<C>::kWarpFloats;          // 3 sub-bands x 8 rows x 16C floats (row pitch kDetPitch)
    float* stage = det + kDetRows * kDetPitch;

    // ---- the bulk of the tile's input - the level-1 details of the warp's 8 x 16 blocks - starts its way to shared
    // memory now (cp.async: no registers, no stall), so that it is in flight while the LL of the patch is walked
    // down from the top level; issued after that walk it would cost every warp a second, serial trip to HBM.
    if (full) {
        const int gy0 = ty * 32 + 8 * wy, gx0 = tx * 32 + 16 * wx;
        const float* b_hl = g.plane + (int64_t)gy0 * g.pl_stride + (int64_t)(w1 + gx0) * C;
        const float* b_lh = g.plane + (int64_t)(gy0 + h1) * g.pl_stride + (int64_t)gx0 * C;
        const int al = (int)(((uintptr_t)g.plane | (uintptr_t)(g.pl_stride * 4) | (uintptr_t)((int64_t)w1 * C * 4)) & 15);
        if (al == 0) fetch_details<C, 16>(det, b_hl, b_lh, (int64_t)w1 * C, g.pl_stride, lane);
        else if ((al & 7) == 0) fetch_details<C, 8>(det, b_hl, b_lh, (int64_t)w1 * C, g.pl_stride, lane);
        else fetch_details<C, 4>(det, b_hl, b_lh, (int64_t)w1 * C, g.pl_stride, lane);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // ---- LL of the lane's level-2 block (levels >= 2).  Levels 4 and 3 are expanded by one lane per (block, channel)
    // element - lane = row * 4C + block * C + channel, the order of the sub-band rows, so the detail loads are whole
    // runs - and handed down through the stage; a lane then picks up the LL_2 of its own block.  Every global load of
    // levels 4, 3 and 2 is issued here, before the first wait: they do not depend on each other, only the sums do.
    constexpr int seg3 = 4 * C, seg2 = 8 * C;                     // floats of a level-3 / level-2 row of the warp
    const bool in2 = (2 * gy2 < h1) && (2 * gx2 < w1);            // L >= 2: extents are multiples of 4, all or nothing
    // v = {LL, HL, LH, HH} of element e of the run at (gy, gx0) of level T (LL only when take_ll)
    auto load_quad = [&](int T, int gy, int gx0, int e, bool take_ll, float (&v)[4]) {
        const int hT = g.Hp >> T, wT = g.Wp >> T;
        const float* q_hl = g.plane + (int64_t)gy * g.pl_stride + (int64_t)(wT + gx0) * C + e;
        const float* q_lh = g.plane + (int64_t)(gy + hT) * g.pl_stride + (int64_t)gx0 * C + e;
        v[0] = take_ll ? g.ll[(int64_t)gy * g.ll_stride + (int64_t)gx0 * C + e] : 0.f;
        v[1] = q_hl[0]; v[2] = q_lh[0]; v[3] = q_lh[(int64_t)wT * C];
    };
    auto expand = [](const float (&v)[4], float& a, float& b, float& cc, float& d) {
        const float s0 = __fadd_rn(v[0], v[1]), s1 = __fsub_rn(v[0], v[1]);
        const float d0 = __fadd_rn(v[2], v[3]), d1 = __fsub_rn(v[2], v[3]);
        a = __fadd_rn(s0, d0); b = __fadd_rn(s1, d1); cc = __fsub_rn(s0, d0); d = __fsub_rn(s1, d1);
    };
    float v4[4] = {0.f, 0.f, 0.f, 0.f}, v3[4] = {0.f, 0.f, 0.f, 0.f}, ll2[C], hl2[C], lh2[C], hh2[C];
    const int row3 = lane / seg3, j3 = lane % seg3;
    if (L == 4 && lane < 2 * C && ty * 4 + wy < (g.Hp >> 4) && tx * 4 + 2 * wx + lane / C < (g.Wp >> 4))
        load_quad(4, ty * 4 + wy, tx * 4 + 2 * wx, lane, true, v4);
    if (L >= 3 && lane < 2 * seg3 && ty * 8 + 2 * wy + row3 < (g.Hp >> 3) && tx * 8 + 4 * wx + j3 / C < (g.Wp >> 3))
        load_quad(3, ty * 8 + 2 * wy + row3, tx * 8 + 4 * wx, j3, L == 3, v3);
    if (L >= 2) {
        const int h2 = g.Hp >> 2, w2 = g.Wp >> 2;
        const float* q_hl = g.plane + (int64_t)gy2 * g.pl_stride + (int64_t)(w2 + gx2) * C;
        const float* q_lh = g.plane + (int64_t)(gy2 + h2) * g.pl_stride + (int64_t)gx2 * C;
        const float* q_hh = q_lh + (int64_t)w2 * C;
        const float* q_ll = g.ll + (int64_t)gy2 * g.ll_stride + (int64_t)gx2 * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            hl2[c] = in2 ? q_hl[c] : 0.f; lh2[c] = in2 ? q_lh[c] : 0.f; hh2[c] = in2 ? q_hh[c] : 0.f;
            ll2[c] = (in2 && L == 2) ? q_ll[c] : 0.f;
        }
    }
    if (L > 2) {
        float* st3 = stage;                                       // LL_3 of the warp's 2 x 4 blocks
        float* st2 = stage + 2 * seg3;                            // LL_2 of the warp's 4 x 8 blocks
        if (L == 4) {
            if (lane < 2 * C) {
                float a, b, cc, d;
                expand(v4, a, b, cc, d);
                float* q = st3 + 2 * (lane / C) * C + (lane % C);
                q[0] = a; q[C] = b; q[seg3] = cc; q[seg3 + C] = d;
            }
            __syncwarp();
        }
        if (lane < 2 * seg3) {
            if (L == 4) v3[0] = st3[lane];
            float a, b, cc, d;
            expand(v3, a, b, cc, d);
            float* q = st2 + 2 * row3 * seg2 + 2 * (j3 / C) * C + (j3 % C);
            q[0] = a; q[C] = b; q[seg2] = cc; q[seg2 + C] = d;
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < C; ++c) ll2[c] = st2[ly * seg2 + lx * C + c];
        __syncwarp();                                             // the stage carries output rows from here on
    }
    // ---- level 2 -> the four LL_1 values of the patch (or LL_1 itself when only one level is fused)
    float ll1[2][2][C];
    bool in1[2][2];
#pragma unroll
    for (int iy = 0; iy < 2; ++iy)
#pragma unroll
        for (int ix = 0; ix < 2; ++ix) in1[iy][ix] = (2 * gy2 + iy < h1) && (2 * gx2 + ix < w1);
    if (L >= 2) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float s0 = __fadd_rn(ll2[c], hl2[c]), s1 = __fsub_rn(ll2[c], hl2[c]);
            const float d0 = __fadd_rn(lh2[c], hh2[c]), d1 = __fsub_rn(lh2[c], hh2[c]);
            ll1[0][0][c] = __fadd_rn(s0, d0); ll1[0][1][c] = __fadd_rn(s1, d1);
            ll1[1][0][c] = __fsub_rn(s0, d0); ll1[1][1][c] = __fsub_rn(s1, d1);
        }
    } else {
#pragma unroll
        for (int iy = 0; iy < 2; ++iy)
#pragma unroll
            for (int ix = 0; ix < 2; ++ix) {
                const float* q = g.ll + (int64_t)(2 * gy2 + iy) * g.ll_stride + (int64_t)(2 * gx2 + ix) * C;
#pragma unroll
                for (int c = 0; c < C; ++c) ll1[iy][ix][c] = in1[iy][ix] ? q[c] : 0.f;
            }
    }
    // ---- level 1: two block rows; each output row of the patch goes through the stage as soon as it exists
    const bool vec = ((w1 * C) % 2 == 0) && ((uintptr_t)g.plane % 8 == 0);
    const int row0 = ty * kTile + 16 * wy;                        // first output row of the warp's region
    const int rows_left = g.Hp - row0;                            // valid rows of the region (may exceed 16)
    const int seg_valid = (g.Wp - (tx * kTile + 32 * wx)) * C;
    float* obase = out + (int64_t)row0 * out_stride + (int64_t)(tx * kTile + 32 * wx) * C;
#pragma unroll
    for (int iy = 0; iy < 2; ++iy) {
        const int gy1 = 2 * gy2 + iy;
        float hl[2 * C], lh[2 * C], hh[2 * C];
        const bool rin = in1[iy][0];
        const float* q_hl = g.plane + (int64_t)gy1 * g.pl_stride + (int64_t)(w1 + 2 * gx2) * C;
        const float* q_lh = g.plane + (int64_t)(gy1 + h1) * g.pl_stride + (int64_t)(2 * gx2) * C;
        const float* q_hh = q_lh + (int64_t)w1 * C;
        if (full) {
            if (iy == 0) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();                                     // every lane's copies have landed
            }
            const float* d = det + (2 * ly + iy) * kDetPitch + lx * 2 * C;
            load_run<2 * C>(d, true, true, hl); load_run<2 * C>(d + 8 * kDetPitch, true, true, lh);
            load_run<2 * C>(d + 16 * kDetPitch, true, true, hh);
        } else if (in1[iy][0] && in1[iy][1]) {
            load_run<2 * C>(q_hl, vec, rin, hl); load_run<2 * C>(q_lh, vec, rin, lh); load_run<2 * C>(q_hh, vec, rin, hh);
        } else {
#pragma unroll
            for (int k = 0; k < 2 * C; ++k) {
                const bool ok = in1[iy][k / C];
                hl[k] = ok ? q_hl[k] : 0.f; lh[k] = ok ? q_lh[k] : 0.f; hh[k] = ok ? q_hh[k] : 0.f;
            }
        }
        float top[4 * C], bot[4 * C];                             // the two pixel rows this block row expands to
#pragma unroll
        for (int ix = 0; ix < 2; ++ix)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float vll = ll1[iy][ix][c], vhl = hl[ix * C + c], vlh = lh[ix * C + c], vhh = hh[ix * C + c];
                const float s0 = __fadd_rn(vll, vhl), s1 = __fsub_rn(vll, vhl);
                const float d0 = __fadd_rn(vlh, vhh), d1 = __fsub_rn(vlh, vhh);
                top[(2 * ix) * C + c] = __fadd_rn(s0, d0); top[(2 * ix + 1) * C + c] = __fadd_rn(s1, d1);
                bot[(2 * ix) * C + c] = __fsub_rn(s0, d0); bot[(2 * ix + 1) * C + c] = __fsub_rn(s1, d1);
            }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = 2 * iy + half;                          // patch row: region rows r, 4 + r, 8 + r, 12 + r
            float4* q = reinterpret_cast<float4*>(stage + ly * kSeg + lx * 4 * C);
#pragma unroll
            for (int e = 0; e < C; ++e)
                q[e] = half ? make_float4(bot[4 * e], bot[4 * e + 1], bot[4 * e + 2], bot[4 * e + 3])
                            : make_float4(top[4 * e], top[4 * e + 1], top[4 * e + 2], top[4 * e + 3]);
            __syncwarp();
            store_rows<kSeg, 4, kSeg>(stage, lane, obase + (int64_t)r * out_stride, 4 * out_stride, full,
                                      (rows_left - r + 3) >> 2, seg_valid);
            __syncwarp();
        }
    }
}

template <int C>
static cudaError_t launch_forward_tiles(const uint8_t* d_src, int64_t pitch, int H, int W, int border_type,
                                        int border_const, const TileGeom& g, cudaStream_t stream) {
    forward_patch_kernel<C><<<dim3(g.tiles_x, g.tiles_y), kTileThreads, 0, stream>>>(d_src, pitch, H, W, border_type, border_const, g);
    return cudaGetLastError();
}
template <int C>
static cudaError_t launch_inverse_tiles(const TileGeom& g, float* out, int64_t out_stride, cudaStream_t stream) {
    static thread_local int configured_dev = -1;                  // > 48 KB of dynamic shared memory: opt in once per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(inverse_patch_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)InvSmem<C>::kBytes);
        if (e != cudaSuccess) return e;
        configured_dev = dev;
    }
    inverse_patch_kernel<C><<<dim3(g.tiles_x, g.tiles_y), kTileThreads, InvSmem<C>::kBytes, stream>>>(g, out, out_stride);
    return cudaGetLastError();
}

static int grid_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    return (int)b;
}

// d_coeffs: Mallat plane (Hp, Wp, C); d_work: scratch of >= Hp*Wp*C/4 + Hp*Wp*C/16 floats.
cudaError_t launch_forward(const uint8_t* d_src, int64_t pitch, int H, int W, int C, int Hp, int Wp, int depth,
                           int border_type, int border_const, float* d_coeffs, float* d_work, cudaStream_t stream) {
    const int64_t pl_stride = (int64_t)Wp * C;
    float* workA = d_work;                                          // LL of odd levels
    float* workB = d_work + ((int64_t)Hp / 2) * ((int64_t)Wp / 2) * C;   // LL of even levels
    const float* in = nullptr;
    int64_t in_stride = 0;
    int first = 1;
    if (C <= 4) {
        // levels 1..min(depth, kFwdFusedLevels) in one pass per 64 x 64 tile
        TileGeom g;
        g.Hp = Hp; g.Wp = Wp; g.levels = depth < kFwdFusedLevels ? depth : kFwdFusedLevels;
        g.tiles_x = (Wp + kTile - 1) / kTile; g.tiles_y = (Hp + kTile - 1) / kTile;
        g.plane = d_coeffs; g.pl_stride = pl_stride;
        if (depth <= g.levels) { g.ll = d_coeffs; g.ll_stride = pl_stride; }
        else { g.ll = (g.levels & 1) ? workA : workB; g.ll_stride = (int64_t)(Wp >> g.levels) * C; }
        cudaError_t e;
        switch (C) {
            case 1: e = launch_forward_tiles<1>(d_src, pitch, H, W, border_type, border_const, g, stream); break;
            case 2: e = launch_forward_tiles<2>(d_src, pitch, H, W, border_type, border_const, g, stream); break;
            case 3: e = launch_forward_tiles<3>(d_src, pitch, H, W, border_type, border_const, g, stream); break;
            default: e = launch_forward_tiles<4>(d_src, pitch, H, W, border_type, border_const, g, stream); break;
        }
        if (e != cudaSuccess) return e;
        in = g.ll; in_stride = g.ll_stride;
        first = g.levels + 1;
    }
    for (int l = first; l <= depth; ++l) {
        SubbandOut o;
        o.h = Hp >> l; o.w = Wp >> l; o.C = C;
        o.plane = d_coeffs; o.pl_stride = pl_stride;
        if (first > kFwdFusedLevels && l + 1 <= depth) {
            // above the fused levels: two levels per launch (the planes are tiny, both scratch buffers hold any of them)
            SubbandOut o2 = o;
            o2.h = Hp >> (l + 1); o2.w = Wp >> (l + 1);
            if (l + 1 == depth) { o2.ll = d_coeffs; o2.ll_stride = pl_stride; }
            else { o2.ll = (in == workA) ? workB : workA; o2.ll_stride = (int64_t)o2.w * C; }
            o.ll = nullptr; o.ll_stride = 0;
            forward_level2x_f32_kernel<<<grid_for((int64_t)o2.h * o2.w * C), 256, 0, stream>>>(in, in_stride, o, o2);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            in = o2.ll; in_stride = o2.ll_stride;
            ++l;
            continue;
        }
        if (l == depth) { o.ll = d_coeffs; o.ll_stride = pl_stride; }
        else if (first > kFwdFusedLevels) { o.ll = (in == workA) ? workB : workA; o.ll_stride = (int64_t)o.w * C; }
        else { o.ll = (l & 1) ? workA : workB; o.ll_stride = (int64_t)o.w * C; }
        const int64_t n = (int64_t)o.h * o.w * C;
        if (l == 1) forward_level1_u8_kernel<<<grid_for(n), 256, 0, stream>>>(d_src, pitch, H, W, border_type, border_const, o);
        else forward_level_f32_kernel<<<grid_for(n), 256, 0, stream>>>(in, in_stride, o);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        in = o.ll; in_stride = o.ll_stride;
    }
    return cudaSuccess;
}

// d_image: (Hp, Wp, C) fp32 output; d_work: scratch of >= Hp*Wp*C/4 + Hp*Wp*C/16 floats.
cudaError_t launch_inverse(const float* d_coeffs, int Hp, int Wp, int C, int depth, float* d_image, float* d_work,
                           cudaStream_t stream) {
    const int64_t pl_stride = (int64_t)Wp * C;
    float* workA = d_work;
    float* workB = d_work + ((int64_t)Hp / 2) * ((int64_t)Wp / 2) * C;
    const float* ll = d_coeffs;
    int64_t ll_stride = pl_stride;
    const int fused_levels = (C <= 4) ? (depth < kFusedLevels ? depth : kFusedLevels) : 0;
    // levels depth .. fused_levels+1 one by one (only when depth > kFusedLevels or C > 4)
    for (int l = depth; l > fused_levels; --l) {
        const int h = Hp >> l, w = Wp >> l;
        const int64_t n = (int64_t)h * w * C;
        if (fused_levels == kFusedLevels && l - 1 > fused_levels) {
            // above the fused levels: two levels per launch (tiny planes: either scratch buffer holds any of them)
            float* out = (ll == workA) ? workB : workA;
            const int64_t out_stride = (int64_t)(4 * w) * C;
            inverse_level2x_f32_kernel<<<grid_for(n), 256, 0, stream>>>(ll, ll_stride, d_coeffs, pl_stride, out, out_stride, h, w, C);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            ll = out; ll_stride = out_stride;
            --l;
            continue;
        }
        float* out; int64_t out_stride;
        if (l == 1) { out = d_image; out_stride = pl_stride; }
        else if (fused_levels == kFusedLevels) { out = (ll == workA) ? workB : workA; out_stride = (int64_t)(2 * w) * C; }
        else { out = ((l - 1) & 1) ? workA : workB; out_stride = (int64_t)(2 * w) * C; }
        inverse_level_f32_kernel<<<grid_for(n), 256, 0, stream>>>(ll, ll_stride, d_coeffs, pl_stride, out, out_stride, h, w, C);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        ll = out; ll_stride = out_stride;
    }
    if (fused_levels > 0) {
        TileGeom g;
        g.Hp = Hp; g.Wp = Wp; g.levels = fused_levels;
        g.tiles_x = (Wp + kTile - 1) / kTile; g.tiles_y = (Hp + kTile - 1) / kTile;
        g.plane = const_cast<float*>(d_coeffs); g.pl_stride = pl_stride;
        g.ll = const_cast<float*>(ll); g.ll_stride = ll_stride;
        cudaError_t e;
        switch (C) {
            case 1: e = launch_inverse_tiles<1>(g, d_image, pl_stride, stream); break;
            case 2: e = launch_inverse_tiles<2>(g, d_image, pl_stride, stream); break;
            case 3: e = launch_inverse_tiles<3>(g, d_image, pl_stride, stream); break;
            default: e = launch_inverse_tiles<4>(g, d_image, pl_stride, stream); break;
        }
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace wicca
