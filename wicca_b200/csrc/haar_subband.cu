// haar_subband.cu - full sub-band forward / inverse 2-D Haar transform (extension, SURVEY.md
// 8(a) row A4; the reference keeps only LL, wicca/wavelet_coder.py:61-65).
//
// Per level, with the 2x2 block  a b / c d  (a = x[2i,2j], b = x[2i,2j+1], c = x[2i+1,2j]):
//   LL = ((a+c)+(b+d))/4   HL = ((a+c)-(b+d))/4   LH = ((a-c)+(b-d))/4   HH = ((a-c)-(b-d))/4
// i.e. row-pair sum/difference first, then column-pair sum/difference, then *0.25 - the order
// of the reference's LL.  Coefficients are stored Mallat-style in one fp32 HWC plane of the
// padded size: HL_l right of LL_l, LH_l below, HH_l diagonal.
//
// Every thread produces one coefficient quadruple; float4-free scalar stores are coalesced
// because consecutive threads own consecutive (x, c) elements of an output row.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

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

// ------------------------------------------------------------------------------------------
// Fused tile kernels: all levels 1..min(depth, 6) of one 64 x 64-pixel tile in one pass.
// Forward: the uint8 tile is read once (3 B/px), every coefficient is written once (12 B/px);
// the shrinking LL pyramid stays in shared memory.  Inverse: every coefficient is read once,
// the image is written once.  One thread owns one 2 x 2 block with all its channels, and
// consecutive threads own consecutive blocks of a row, so a warp reads/writes runs of whole
// sectors; the channel count is a template parameter and every index is a shift.
// ------------------------------------------------------------------------------------------
constexpr int kTile = 64;
constexpr int kTileThreads = 256;

struct TileGeom {
    int Hp, Wp;                 // padded extents (multiples of 2^depth)
    int levels;                 // levels done inside the tile: min(depth, 6)
    int tiles_x, tiles_y;
    float* plane; int64_t pl_stride;      // Mallat plane
    float* ll; int64_t ll_stride;         // where LL_levels lives (the plane itself when depth <= 6)
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

// 2*C consecutive bytes (two horizontally adjacent pixels) -> floats, with the widest loads the
// alignment allows (p is 4-byte aligned when `wide`).
template <int C>
__device__ __forceinline__ void load_pixel_pair(const uint8_t* p, bool wide, float (&a)[C], float (&b)[C]) {
    uint8_t v[2 * C];
    if (wide && (2 * C) % 4 == 0) {
#pragma unroll
        for (int k = 0; k < 2 * C / 4; ++k) {
            const uint32_t w = reinterpret_cast<const uint32_t*>(p)[k];
            v[4 * k] = (uint8_t)w; v[4 * k + 1] = (uint8_t)(w >> 8); v[4 * k + 2] = (uint8_t)(w >> 16); v[4 * k + 3] = (uint8_t)(w >> 24);
        }
    } else if (wide) {                       // 2*C is even: 16-bit loads (p is even because x is even)
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const uint16_t w = reinterpret_cast<const uint16_t*>(p)[k];
            v[2 * k] = (uint8_t)w; v[2 * k + 1] = (uint8_t)(w >> 8);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 2 * C; ++k) v[k] = p[k];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) { a[c] = (float)v[c]; b[c] = (float)v[C + c]; }
}

template <int C>
__global__ void __launch_bounds__(kTileThreads)
forward_tile_kernel(const uint8_t* __restrict__ src, int64_t pitch, int H, int W, int border_type, int border_const,
                    TileGeom g) {
    __shared__ float s_a[32 * 32 * C];
    __shared__ float s_b[16 * 16 * C];
    // level-1 rows are handed from "one thread = one block, all channels" to "one lane = one float of
    // the row" through this per-warp buffer, so that every store instruction covers 128 contiguous bytes
    __shared__ float s_st[kTileThreads / 32][4][32 * C];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
    const int y0 = ty * kTile, x0 = tx * kTile;
    const bool interior = (y0 + kTile <= H) && (x0 + kTile <= W);
    const float fc = (float)border_const;
    const bool wide = ((uintptr_t)src % 4 == 0) && (pitch % 4 == 0);      // x0*C and 2*bx*C are multiples of 2*C
    const float* in_f = nullptr;
#pragma unroll
    for (int l = 1; l <= 6; ++l) {
        if (l > g.levels) break;
        const int sh = 6 - l;                           // log2(blocks per tile side)
        const int n = 1 << sh;
        const int hl_ = g.Hp >> l, wl_ = g.Wp >> l;     // sub-band extents
        float* out_f = (l & 1) ? s_a : s_b;
        const bool last = (l == g.levels);
        for (int blk = threadIdx.x; blk < n * n; blk += kTileThreads) {
            const int by = blk >> sh, bx = blk & (n - 1);
            float a[C], b[C], cc[C], d[C];
            if (l == 1) {
                if (interior) {
                    const uint8_t* p = src + (int64_t)(y0 + 2 * by) * pitch + (int64_t)(x0 + 2 * bx) * C;
                    load_pixel_pair<C>(p, wide, a, b);
                    load_pixel_pair<C>(p + pitch, wide, cc, d);
                } else {
                    const int ya = border_index(y0 + 2 * by, H, border_type), yb = border_index(y0 + 2 * by + 1, H, border_type);
                    const int xa = border_index(x0 + 2 * bx, W, border_type), xb = border_index(x0 + 2 * bx + 1, W, border_type);
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        a[c] = (ya < 0 || xa < 0) ? fc : (float)src[(int64_t)ya * pitch + (int64_t)xa * C + c];
                        b[c] = (ya < 0 || xb < 0) ? fc : (float)src[(int64_t)ya * pitch + (int64_t)xb * C + c];
                        cc[c] = (yb < 0 || xa < 0) ? fc : (float)src[(int64_t)yb * pitch + (int64_t)xa * C + c];
                        d[c] = (yb < 0 || xb < 0) ? fc : (float)src[(int64_t)yb * pitch + (int64_t)xb * C + c];
                    }
                }
            } else {
                const int in_row = 2 * n * C;
                const float* p = in_f + (2 * by) * in_row + (2 * bx) * C;
#pragma unroll
                for (int c = 0; c < C; ++c) { a[c] = p[c]; b[c] = p[C + c]; cc[c] = p[in_row + c]; d[c] = p[in_row + C + c]; }
            }
            float vll[C], vhl[C], vlh[C], vhh[C];
            analyse<C>(a, b, cc, d, vll, vhl, vlh, vhh);
            const int gy = ty * n + by, gx = tx * n + bx;
            if (l == 1) {
                // the warp owns block row `by` (bx == lane): stage, then store lane-contiguous rows
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    s_st[warp][0][lane * C + c] = vhl[c]; s_st[warp][1][lane * C + c] = vlh[c];
                    s_st[warp][2][lane * C + c] = vhh[c]; s_st[warp][3][lane * C + c] = vll[c];
                }
                __syncwarp();
                if (gy < hl_) {
                    const int gx0 = tx * n;
                    int valid = (wl_ - gx0) * C;                 // floats of this row segment inside the sub-band
                    if (valid > 32 * C) valid = 32 * C;
                    float* r_hl = g.plane + (int64_t)gy * g.pl_stride + (int64_t)(wl_ + gx0) * C;
                    float* r_lh = g.plane + (int64_t)(gy + hl_) * g.pl_stride + (int64_t)gx0 * C;
                    float* r_hh = r_lh + (int64_t)wl_ * C;
                    float* r_ll = g.ll + (int64_t)gy * g.ll_stride + (int64_t)gx0 * C;
#pragma unroll
                    for (int k = 0; k < C; ++k) {
                        const int jj = lane + 32 * k;
                        if (jj < valid) {
                            r_hl[jj] = s_st[warp][0][jj]; r_lh[jj] = s_st[warp][1][jj]; r_hh[jj] = s_st[warp][2][jj];
                            if (last) r_ll[jj] = s_st[warp][3][jj];
                        }
                    }
                }
                __syncwarp();
            } else if (gy < hl_ && gx < wl_) {
                float* q_hl = g.plane + (int64_t)gy * g.pl_stride + (int64_t)(wl_ + gx) * C;
                float* q_lh = g.plane + (int64_t)(gy + hl_) * g.pl_stride + (int64_t)gx * C;
                float* q_hh = q_lh + (int64_t)wl_ * C;
#pragma unroll
                for (int c = 0; c < C; ++c) { q_hl[c] = vhl[c]; q_lh[c] = vlh[c]; q_hh[c] = vhh[c]; }
                if (last) {
                    float* q_ll = g.ll + (int64_t)gy * g.ll_stride + (int64_t)gx * C;
#pragma unroll
                    for (int c = 0; c < C; ++c) q_ll[c] = vll[c];
                }
            }
            if (!last) {
                float* q = out_f + (by * n + bx) * C;
#pragma unroll
                for (int c = 0; c < C; ++c) q[c] = vll[c];
            }
        }
        __syncthreads();
        in_f = out_f;
    }
}

template <int C>
__global__ void __launch_bounds__(kTileThreads)
inverse_tile_kernel(TileGeom g, float* __restrict__ out, int64_t out_stride) {
    __shared__ float s_a[32 * 32 * C];
    __shared__ float s_b[16 * 16 * C];
    __shared__ float s_st[kTileThreads / 32][2][64 * C];    // per-warp staging of two output rows (see forward kernel)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
    const float* in_f = nullptr;
    for (int l = g.levels; l >= 1; --l) {
        const int sh = 6 - l;
        const int n = 1 << sh;                          // LL_l tile is n x n
        const int hl_ = g.Hp >> l, wl_ = g.Wp >> l;
        float* out_f = ((l - 1) & 1) ? s_a : s_b;
        for (int blk = threadIdx.x; blk < n * n; blk += kTileThreads) {
            const int by = blk >> sh, bx = blk & (n - 1);
            const int gy = ty * n + by, gx = tx * n + bx;
            const bool inside = gy < hl_ && gx < wl_;
            float vll[C], vhl[C], vlh[C], vhh[C];
#pragma unroll
            for (int c = 0; c < C; ++c) vll[c] = vhl[c] = vlh[c] = vhh[c] = 0.f;
            if (inside) {
                const float* q_ll = (l == g.levels) ? g.ll + (int64_t)gy * g.ll_stride + (int64_t)gx * C
                                                    : in_f + (by * n + bx) * C;
                const float* q_hl = g.plane + (int64_t)gy * g.pl_stride + (int64_t)(wl_ + gx) * C;
                const float* q_lh = g.plane + (int64_t)(gy + hl_) * g.pl_stride + (int64_t)gx * C;
                const float* q_hh = q_lh + (int64_t)wl_ * C;
#pragma unroll
                for (int c = 0; c < C; ++c) { vll[c] = q_ll[c]; vhl[c] = q_hl[c]; vlh[c] = q_lh[c]; vhh[c] = q_hh[c]; }
            }
            float a[C], b[C], cc[C], d[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float s0 = __fadd_rn(vll[c], vhl[c]), s1 = __fsub_rn(vll[c], vhl[c]);
                const float d0 = __fadd_rn(vlh[c], vhh[c]), d1 = __fsub_rn(vlh[c], vhh[c]);
                a[c] = __fadd_rn(s0, d0); b[c] = __fadd_rn(s1, d1); cc[c] = __fsub_rn(s0, d0); d[c] = __fsub_rn(s1, d1);
            }
            if (l == 1) {
                // the warp owns block row `by` (bx == lane): two output rows of 64*C floats, staged and then
                // stored with lane-contiguous addresses (128 contiguous bytes per store instruction)
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    s_st[warp][0][lane * 2 * C + c] = a[c]; s_st[warp][0][lane * 2 * C + C + c] = b[c];
                    s_st[warp][1][lane * 2 * C + c] = cc[c]; s_st[warp][1][lane * 2 * C + C + c] = d[c];
                }
                __syncwarp();
                if (gy < hl_) {
                    const int gx0 = tx * n;
                    int valid = (wl_ - gx0) * 2 * C;
                    if (valid > 64 * C) valid = 64 * C;
                    float* r0 = out + (int64_t)(2 * gy) * out_stride + (int64_t)(2 * gx0) * C;
                    float* r1 = r0 + out_stride;
#pragma unroll
                    for (int k = 0; k < 2 * C; ++k) {
                        const int jj = lane + 32 * k;
                        if (jj < valid) { r0[jj] = s_st[warp][0][jj]; r1[jj] = s_st[warp][1][jj]; }
                    }
                }
                __syncwarp();
            } else {
                const int out_row = 2 * n * C;
                float* q = out_f + (2 * by) * out_row + (2 * bx) * C;
#pragma unroll
                for (int c = 0; c < C; ++c) { q[c] = a[c]; q[C + c] = b[c]; q[out_row + c] = cc[c]; q[out_row + C + c] = d[c]; }
            }
        }
        __syncthreads();
        in_f = out_f;
    }
}

// ------------------------------------------------------------------------------------------
// Forward, depth >= 2: the same 64 x 64 tile, but each lane owns a whole 4 x 4-pixel patch (one
// level-2 block): levels 1 and 2 are computed in registers, levels 3 and 4 with warp shuffles, and
// only the sixteen LL_4 values of the tile cross warps (one barrier instead of one per level).
// Warp w covers level-2 blocks rows 4*(w>>1)..+3, columns 8*(w&1)..+7 (lane = 8*ly + lx).
// ------------------------------------------------------------------------------------------
// Store R x (UNIT*UPR) staged floats as rows of a sub-band: store instruction k writes the 32 consecutive
// staged floats 32k..32k+31, i.e. 32/UNIT whole UNIT-float pieces, each inside one row.
template <int R, int UNIT, int UPR, int PITCH, bool kCheck>
__device__ __forceinline__ void store_staged_impl(const float* stage, int lane, float* base, int stride, int rows_valid,
                                                  int seg_valid) {
    constexpr int K = R * UNIT * UPR / 32;               // store instructions
    constexpr int PER = 32 / UNIT;                       // pieces per instruction
    const int piece = lane / UNIT, lo = lane % UNIT;
    float* lbase = base + lo;
    const float* lstage = stage + lo;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        // piece 0 of the instruction is at a compile-time position; the others are written as lane-dependent
        // deltas that repeat across k, so the compiler keeps them in a few registers
        const int r0 = (k * PER) / UPR, j0 = ((k * PER) % UPR) * UNIT;
        int so = r0 * PITCH + j0, go = r0 * stride + j0, r = r0, j = j0;
#pragma unroll
        for (int q = 1; q < PER; ++q) {
            const int rq = (k * PER + q) / UPR, jq = ((k * PER + q) % UPR) * UNIT;
            const int sel = (piece == q) ? 1 : 0;
            so += sel * ((rq - r0) * PITCH + (jq - j0));
            go += sel * ((rq - r0) * stride + (jq - j0));
            if (kCheck) { r += sel * (rq - r0); j += sel * (jq - j0); }
        }
        if (!kCheck || (r < rows_valid && j + lo < seg_valid)) lbase[go] = lstage[so];
    }
}
template <int C, int R, int UNIT, int UPR, int PITCH>
__device__ __forceinline__ void store_staged(const float* stage, int lane, float* base, int stride, bool full,
                                             int rows_valid, int seg_valid) {
    if (full) store_staged_impl<R, UNIT, UPR, PITCH, false>(stage, lane, base, stride, rows_valid, seg_valid);
    else store_staged_impl<R, UNIT, UPR, PITCH, true>(stage, lane, base, stride, rows_valid, seg_valid);
}

template <int C>
__global__ void __launch_bounds__(kTileThreads, C <= 3 ? 4 : 3)
forward_patch_kernel(const uint8_t* __restrict__ src, int64_t pitch, int H, int W, int border_type, int border_const,
                     TileGeom g) {
    constexpr int kPitch1 = 16 * C + 8;                        // staged level-1 row (+8: float2 writes hit every bank once)
    __shared__ __align__(16) float s_st[kTileThreads / 32][8 * kPitch1];   // per-warp staging of one sub-band of one level
    __shared__ float s_ll2[16][16][C], s_ll3[8][8][C];
    __shared__ int s_arrived;                                  // warps that have published their LL_2 values
    if (threadIdx.x == 0) s_arrived = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy = warp >> 1, wx = warp & 1, ly = lane >> 3, lx = lane & 7;
    const int tile = blockIdx.x;
    const int ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
    const int b2y = 4 * wy + ly, b2x = 8 * wx + lx;                       // level-2 block inside the tile
    const int py = ty * kTile + 4 * b2y, px = tx * kTile + 4 * b2x;       // top-left pixel of the lane's patch
    const bool interior = (ty * kTile + kTile <= H) && (tx * kTile + kTile <= W);
    const bool full = (ty * kTile + kTile <= g.Hp) && (tx * kTile + kTile <= g.Wp);   // no clipped sub-band rows
    const bool wide = ((uintptr_t)src % 4 == 0) && (pitch % 4 == 0);
    const float fc = (float)border_const;
    float* stage = s_st[warp];
    const int stride = (int)g.pl_stride;

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
        put(hl1); store_staged<C, 8, 16, C, kPitch1>(stage, lane, base_hl, stride, full, rows_valid, seg_valid); __syncwarp();
        put(lh1); store_staged<C, 8, 16, C, kPitch1>(stage, lane, base_lh, stride, full, rows_valid, seg_valid); __syncwarp();
        put(hh1); store_staged<C, 8, 16, C, kPitch1>(stage, lane, base_hh, stride, full, rows_valid, seg_valid); __syncwarp();
        if (g.levels == 1) {
            put(ll1);
            store_staged<C, 8, 16, C, kPitch1>(stage, lane, g.ll + (int64_t)gy0 * g.ll_stride + (int64_t)gx0 * C,
                                               (int)g.ll_stride, full, rows_valid, seg_valid);
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
        put(hl2); store_staged<C, 4, 8, C, seg>(stage, lane, base_hl, stride, full, rows_valid, seg_valid); __syncwarp();
        put(lh2); store_staged<C, 4, 8, C, seg>(stage, lane, base_lh, stride, full, rows_valid, seg_valid); __syncwarp();
        put(hh2); store_staged<C, 4, 8, C, seg>(stage, lane, base_hh, stride, full, rows_valid, seg_valid); __syncwarp();
        if (g.levels == 2) {
            put(ll2);
            store_staged<C, 4, 8, C, seg>(stage, lane, g.ll + (int64_t)gy0 * g.ll_stride + (int64_t)gx0 * C, (int)g.ll_stride,
                                          full, rows_valid, seg_valid);
            return;
        }
    }
    // ---- levels 3..6 need LL_2 of other warps: every warp publishes its 4 x 8 LL_2 values, and whichever warp
    //      publishes last finishes the tile alone (nobody waits at a barrier)
#pragma unroll
    for (int c = 0; c < C; ++c) s_ll2[b2y][b2x][c] = ll2[c];
    __syncwarp();
    int is_last = 0;
    if (lane == 0) { __threadfence_block(); is_last = (atomicAdd(&s_arrived, 1) == kTileThreads / 32 - 1); }
    is_last = __shfl_sync(0xFFFFFFFFu, is_last, 0);
    if (!is_last) return;
    __threadfence_block();
    auto write_details = [&](int level, int gy, int gx, const float (&hl)[C], const float (&lh)[C], const float (&hh)[C],
                             const float (&ll)[C]) {
        const int hL = g.Hp >> level, wL = g.Wp >> level;
        if (gy >= hL || gx >= wL) return;
        float* q_hl = g.plane + (int64_t)gy * g.pl_stride + (int64_t)(wL + gx) * C;
        float* q_lh = g.plane + (int64_t)(gy + hL) * g.pl_stride + (int64_t)gx * C;
        float* q_hh = q_lh + (int64_t)wL * C;
#pragma unroll
        for (int c = 0; c < C; ++c) { q_hl[c] = hl[c]; q_lh[c] = lh[c]; q_hh[c] = hh[c]; }
        if (g.levels == level) {
            float* q_ll = g.ll + (int64_t)gy * g.ll_stride + (int64_t)gx * C;
#pragma unroll
            for (int c = 0; c < C; ++c) q_ll[c] = ll[c];
        }
    };
    // level 3: 8 x 8 blocks, two per lane
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int by = (lane >> 3) + 4 * i, bx = lane & 7;
        float ll[C], hl[C], lh[C], hh[C];
        analyse<C>(s_ll2[2 * by][2 * bx], s_ll2[2 * by][2 * bx + 1], s_ll2[2 * by + 1][2 * bx], s_ll2[2 * by + 1][2 * bx + 1], ll,
                   hl, lh, hh);
        write_details(3, ty * 8 + by, tx * 8 + bx, hl, lh, hh, ll);
#pragma unroll
        for (int c = 0; c < C; ++c) s_ll3[by][bx][c] = ll[c];
    }
    if (g.levels == 3) return;
    __syncwarp();
    // level 4: 4 x 4 blocks on lanes 0..15 (lane = 4 * by + bx; the upper half-warp mirrors it and stores nothing)
    float ll4[C];
    {
        const int by = (lane >> 2) & 3, bx = lane & 3;
        float hl[C], lh[C], hh[C];
        analyse<C>(s_ll3[2 * by][2 * bx], s_ll3[2 * by][2 * bx + 1], s_ll3[2 * by + 1][2 * bx], s_ll3[2 * by + 1][2 * bx + 1], ll4,
                   hl, lh, hh);
        if (lane < 16) write_details(4, ty * 4 + by, tx * 4 + bx, hl, lh, hh, ll4);
    }
    if (g.levels == 4) return;
    // level 5: lane groups {m, m+1, m+4, m+5}
    float ll5[C];
    {
        const int base = lane & 10;
        float a[C], b[C], cc[C], d[C], hl[C], lh[C], hh[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            a[c] = __shfl_sync(0xFFFFFFFFu, ll4[c], base); b[c] = __shfl_sync(0xFFFFFFFFu, ll4[c], base + 1);
            cc[c] = __shfl_sync(0xFFFFFFFFu, ll4[c], base + 4); d[c] = __shfl_sync(0xFFFFFFFFu, ll4[c], base + 5);
        }
        analyse<C>(a, b, cc, d, ll5, hl, lh, hh);
        if (lane == base) write_details(5, ty * 2 + (lane >> 3), tx * 2 + ((lane >> 1) & 1), hl, lh, hh, ll5);
    }
    if (g.levels == 5) return;
    // level 6: the four level-5 owners are lanes 0, 2, 8, 10
    {
        float a[C], b[C], cc[C], d[C], ll6[C], hl[C], lh[C], hh[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            a[c] = __shfl_sync(0xFFFFFFFFu, ll5[c], 0); b[c] = __shfl_sync(0xFFFFFFFFu, ll5[c], 2);
            cc[c] = __shfl_sync(0xFFFFFFFFu, ll5[c], 8); d[c] = __shfl_sync(0xFFFFFFFFu, ll5[c], 10);
        }
        analyse<C>(a, b, cc, d, ll6, hl, lh, hh);
        if (lane == 0) write_details(6, ty, tx, hl, lh, hh, ll6);
    }
}

template <int C>
static cudaError_t launch_forward_tiles(const uint8_t* d_src, int64_t pitch, int H, int W, int border_type,
                                        int border_const, const TileGeom& g, cudaStream_t stream) {
    // depth 1: one warp per full tile row (384-byte row segments); deeper: one 4 x 4 patch per lane, one barrier
    if (getenv("WICCA_FORWARD_TILE"))
        forward_tile_kernel<C><<<g.tiles_x * g.tiles_y, kTileThreads, 0, stream>>>(d_src, pitch, H, W, border_type, border_const, g);
    else
        forward_patch_kernel<C><<<g.tiles_x * g.tiles_y, kTileThreads, 0, stream>>>(d_src, pitch, H, W, border_type, border_const, g);
    return cudaGetLastError();
}
template <int C>
static cudaError_t launch_inverse_tiles(const TileGeom& g, float* out, int64_t out_stride, cudaStream_t stream) {
    inverse_tile_kernel<C><<<g.tiles_x * g.tiles_y, kTileThreads, 0, stream>>>(g, out, out_stride);
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
        // levels 1..min(depth,6) in one pass per 64 x 64 tile
        TileGeom g;
        g.Hp = Hp; g.Wp = Wp; g.levels = depth < 6 ? depth : 6;
        g.tiles_x = (Wp + kTile - 1) / kTile; g.tiles_y = (Hp + kTile - 1) / kTile;
        g.plane = d_coeffs; g.pl_stride = pl_stride;
        if (depth <= 6) { g.ll = d_coeffs; g.ll_stride = pl_stride; }
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
        if (l == depth) { o.ll = d_coeffs; o.ll_stride = pl_stride; }
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
    const int fused_levels = (C <= 4) ? (depth < 6 ? depth : 6) : 0;
    // levels depth .. fused_levels+1 one by one (only when depth > 6 or C > 4)
    for (int l = depth; l > fused_levels; --l) {
        const int h = Hp >> l, w = Wp >> l;
        float* out; int64_t out_stride;
        if (l == 1) { out = d_image; out_stride = pl_stride; }
        else { out = ((l - 1) & 1) ? workA : workB; out_stride = (int64_t)(2 * w) * C; }
        const int64_t n = (int64_t)h * w * C;
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
