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

template <int C>
static cudaError_t launch_forward_tiles(const uint8_t* d_src, int64_t pitch, int H, int W, int border_type,
                                        int border_const, const TileGeom& g, cudaStream_t stream) {
    forward_tile_kernel<C><<<g.tiles_x * g.tiles_y, kTileThreads, 0, stream>>>(d_src, pitch, H, W, border_type, border_const, g);
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
