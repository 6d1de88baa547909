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
// the image is written once.  Consecutive threads own consecutive (x, c) elements of a sub-band
// row, so every global access is a run of whole 32-byte sectors.
// ------------------------------------------------------------------------------------------
constexpr int kTile = 64;
constexpr int kTileThreads = 256;

struct TileGeom {
    int Hp, Wp, C;              // padded extents (multiples of 2^depth), channels (1..4)
    int levels;                 // levels done inside the tile: min(depth, 6)
    int tiles_x, tiles_y;
    float* plane; int64_t pl_stride;      // Mallat plane
    float* ll; int64_t ll_stride;         // where LL_levels lives (the plane itself when depth <= 6)
};

__global__ void __launch_bounds__(kTileThreads)
forward_tile_kernel(const uint8_t* __restrict__ src, int64_t pitch, int H, int W, int border_type, int border_const,
                    TileGeom g) {
    __shared__ __align__(16) uint8_t s_u8[kTile * kTile * 4];
    __shared__ float s_a[32 * 32 * 4];
    __shared__ float s_b[16 * 16 * 4];
    const int C = g.C;
    const int row_bytes = kTile * C;
    const bool vec_ok = ((uintptr_t)src % 16 == 0) && (pitch % 16 == 0);
    const int n_tiles = g.tiles_x * g.tiles_y;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
        const int y0 = ty * kTile, x0 = tx * kTile;
        // ---- load the tile (border-extended) into shared memory
        if (vec_ok && y0 + kTile <= H && x0 + kTile <= W) {
            const int per_row = row_bytes / 16;                       // 4*C uint4 per row
            for (int i = threadIdx.x; i < kTile * per_row; i += kTileThreads) {
                const int r = i / per_row, q = i - r * per_row;
                const uint4 v = *reinterpret_cast<const uint4*>(src + (int64_t)(y0 + r) * pitch + (int64_t)x0 * C + q * 16);
                *reinterpret_cast<uint4*>(s_u8 + r * row_bytes + q * 16) = v;
            }
        } else {
            for (int i = threadIdx.x; i < kTile * row_bytes; i += kTileThreads) {
                const int r = i / row_bytes, b = i - r * row_bytes;
                const int px = b / C, c = b - px * C;
                uint8_t v = 0;
                if (y0 + r < g.Hp && x0 + px < g.Wp) {
                    const int ym = border_index(y0 + r, H, border_type), xm = border_index(x0 + px, W, border_type);
                    v = (ym < 0 || xm < 0) ? (uint8_t)border_const : src[(int64_t)ym * pitch + (int64_t)xm * C + c];
                }
                s_u8[i] = v;
            }
        }
        __syncthreads();
        // ---- levels
        const float* in_f = nullptr;
        for (int l = 1; l <= g.levels; ++l) {
            const int n = kTile >> l;                       // blocks per tile side at this level
            const int hl = g.Hp >> l, wl = g.Wp >> l;       // sub-band extents
            const int in_row = 2 * n * C;                   // elements per row of this level's input
            float* out_f = (l & 1) ? s_a : s_b;
            const bool last = (l == g.levels);
            for (int e = threadIdx.x; e < n * n * C; e += kTileThreads) {
                const int by = e / (n * C), r = e - by * (n * C);
                const int bx = r / C, c = r - bx * C;
                float a, b, cc, d;
                if (l == 1) {
                    const uint8_t* p = s_u8 + (2 * by) * row_bytes + (2 * bx) * C + c;
                    a = (float)p[0]; b = (float)p[C]; cc = (float)p[row_bytes]; d = (float)p[row_bytes + C];
                } else {
                    const float* p = in_f + (2 * by) * in_row + (2 * bx) * C + c;
                    a = p[0]; b = p[C]; cc = p[in_row]; d = p[in_row + C];
                }
                const float rs0 = __fadd_rn(a, cc), rs1 = __fadd_rn(b, d);
                const float rd0 = __fsub_rn(a, cc), rd1 = __fsub_rn(b, d);
                const float vll = __fmul_rn(__fadd_rn(rs0, rs1), 0.25f);
                const int gy = ty * n + by, gx = tx * n + bx;
                if (gy < hl && gx < wl) {
                    const int64_t ecol = (int64_t)gx * C + c;
                    g.plane[(int64_t)gy * g.pl_stride + (int64_t)wl * C + ecol] = __fmul_rn(__fsub_rn(rs0, rs1), 0.25f);
                    g.plane[(int64_t)(gy + hl) * g.pl_stride + ecol] = __fmul_rn(__fadd_rn(rd0, rd1), 0.25f);
                    g.plane[(int64_t)(gy + hl) * g.pl_stride + (int64_t)wl * C + ecol] = __fmul_rn(__fsub_rn(rd0, rd1), 0.25f);
                    if (last) g.ll[(int64_t)gy * g.ll_stride + ecol] = vll;
                }
                if (!last) out_f[by * (n * C) + r] = vll;
            }
            __syncthreads();
            in_f = out_f;
        }
    }
}

__global__ void __launch_bounds__(kTileThreads)
inverse_tile_kernel(TileGeom g, float* __restrict__ out, int64_t out_stride) {
    extern __shared__ __align__(16) float s_dyn[];          // final 64 x 64 x C tile
    __shared__ float s_a[32 * 32 * 4];
    __shared__ float s_b[16 * 16 * 4];
    const int C = g.C;
    const int n_tiles = g.tiles_x * g.tiles_y;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
        const float* in_f = nullptr;
        for (int l = g.levels; l >= 1; --l) {
            const int n = kTile >> l;                       // LL_l tile is n x n
            const int hl = g.Hp >> l, wl = g.Wp >> l;
            float* out_f = (l == 1) ? s_dyn : (((l - 1) & 1) ? s_a : s_b);
            const int out_row = 2 * n * C;
            for (int e = threadIdx.x; e < n * n * C; e += kTileThreads) {
                const int by = e / (n * C), r = e - by * (n * C);
                const int bx = r / C, c = r - bx * C;
                const int gy = ty * n + by, gx = tx * n + bx;
                float vll = 0.f, vhl = 0.f, vlh = 0.f, vhh = 0.f;
                if (gy < hl && gx < wl) {
                    const int64_t ecol = (int64_t)gx * C + c;
                    vll = (l == g.levels) ? g.ll[(int64_t)gy * g.ll_stride + ecol] : in_f[by * (n * C) + r];
                    vhl = g.plane[(int64_t)gy * g.pl_stride + (int64_t)wl * C + ecol];
                    vlh = g.plane[(int64_t)(gy + hl) * g.pl_stride + ecol];
                    vhh = g.plane[(int64_t)(gy + hl) * g.pl_stride + (int64_t)wl * C + ecol];
                }
                const float s0 = __fadd_rn(vll, vhl), s1 = __fsub_rn(vll, vhl);
                const float d0 = __fadd_rn(vlh, vhh), d1 = __fsub_rn(vlh, vhh);
                float* q = out_f + (2 * by) * out_row + (2 * bx) * C + c;
                q[0] = __fadd_rn(s0, d0);
                q[C] = __fadd_rn(s1, d1);
                q[out_row] = __fsub_rn(s0, d0);
                q[out_row + C] = __fsub_rn(s1, d1);
            }
            __syncthreads();
            in_f = out_f;
        }
        // ---- write the reconstructed 64 x 64 x C tile, whole rows of consecutive floats
        const int row_f = kTile * C;
        for (int i = threadIdx.x; i < kTile * row_f; i += kTileThreads) {
            const int r = i / row_f, b = i - r * row_f;
            const int y = ty * kTile + r, xf = tx * row_f + b;
            if (y < g.Hp && xf < g.Wp * C) out[(int64_t)y * out_stride + xf] = s_dyn[i];
        }
        __syncthreads();
    }
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
        g.Hp = Hp; g.Wp = Wp; g.C = C; g.levels = depth < 6 ? depth : 6;
        g.tiles_x = (Wp + kTile - 1) / kTile; g.tiles_y = (Hp + kTile - 1) / kTile;
        g.plane = d_coeffs; g.pl_stride = pl_stride;
        if (depth <= 6) { g.ll = d_coeffs; g.ll_stride = pl_stride; }
        else { g.ll = (g.levels & 1) ? workA : workB; g.ll_stride = (int64_t)(Wp >> g.levels) * C; }
        const int64_t tiles = (int64_t)g.tiles_x * g.tiles_y;
        const int grid = (int)(tiles < 148 * 8 ? tiles : 148 * 8);
        forward_tile_kernel<<<grid, kTileThreads, 0, stream>>>(d_src, pitch, H, W, border_type, border_const, g);
        cudaError_t e = cudaGetLastError();
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
        g.Hp = Hp; g.Wp = Wp; g.C = C; g.levels = fused_levels;
        g.tiles_x = (Wp + kTile - 1) / kTile; g.tiles_y = (Hp + kTile - 1) / kTile;
        g.plane = const_cast<float*>(d_coeffs); g.pl_stride = pl_stride;
        g.ll = const_cast<float*>(ll); g.ll_stride = ll_stride;
        const size_t smem = (size_t)kTile * kTile * C * sizeof(float);
        static thread_local int configured_dev = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (configured_dev != dev) {
            cudaError_t e = cudaFuncSetAttribute(inverse_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            if (e != cudaSuccess) return e;
            configured_dev = dev;
        }
        const int64_t tiles = (int64_t)g.tiles_x * g.tiles_y;
        const int grid = (int)(tiles < 148 * 3 ? tiles : 148 * 3);
        inverse_tile_kernel<<<grid, kTileThreads, smem, stream>>>(g, d_image, pl_stride);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace wicca
