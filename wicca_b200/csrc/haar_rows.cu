// haar_rows.cu - the icon path outside the one-pass TMA kernel's domain (haar_icon.cu: C == 3, 16-byte aligned,
// depths 1..6): any channel count, any alignment, depth 1..8, and the levels above 6 of a fused run.
// Same arithmetic as everywhere: icon_d = (sum of the 2^d x 2^d block of the border-extended image) >> 2d, the
// integer identity of HaarCoder.get_small_copy (wicca/wavelet_coder.py:56-67 with the padding of
// wicca/data_loader.py:107-117 as an index map).
//
//   haar_icon_rows_kernel   one depth per launch.  A CTA owns a run of whole output pixels of one output row
//                           (about 1 KB of every input row): each thread streams ONE 32-bit word per input row down the
//                           2^d rows (coalesced 128-byte warp requests, column sums in two packed 16-bit lanes:
//                           255 * 256 < 2^16), the byte-column sums meet in shared memory, and the 2^d pixels of a
//                           group are summed channel by channel with a few threads per output.  HBM bound by design
//                           (replaces a kernel that gave every output element to one thread with r^2 serial byte loads).
//   haar_tail_kernel        depths 7, 8 (and the exact level-8 plane from which depths > 8 continue in float32) from
//                           the exact uint32 level-6 block sums the one-pass kernel leaves in a scratch plane, so
//                           depths 1..8 of an RGB image still cost ONE pass over it.  Level-6 blocks that lie entirely
//                           in the padding of the deeper level (beyond the depth-6 extents) are summed directly from
//                           the image through the border index map, a warp per output pixel.
#include <cuda_runtime.h>
#include <stdint.h>

#include "haar_math.cuh"
#include "icon_types.h"
#include "kernels.h"

namespace wicca {

namespace {

constexpr int kRowsThreads = 256;
constexpr int kRowsMaxTileBytes = 8192;         // 8 words per thread

// One border-extended byte of image row ym (ym >= 0) at byte offset b of the padded row.
__device__ __forceinline__ uint32_t padded_byte(const GenericIconArgs& a, const uint8_t* row, int64_t b) {
    const int x = (int)(b / a.C);
    const int c = (int)(b - (int64_t)x * a.C);
    const int xm = border_index(x, a.W, a.border_type);
    return xm < 0 ? (uint32_t)a.border_const : (uint32_t)row[(int64_t)xm * a.C + c];
}

template <int WPT>
__global__ void __launch_bounds__(kRowsThreads)
haar_icon_rows_kernel(GenericIconArgs a, int groups, int rows_per_cta, int aligned) {
    __shared__ __align__(16) uint16_t colsum[kRowsMaxTileBytes];
    __shared__ uint32_t outsum[1024];
    const int tid = threadIdx.x;
    const int r = 1 << a.depth;
    const int gb = a.C << a.depth;                    // bytes of one output pixel's input row segment
    const int tile_bytes = groups * gb;               // multiple of 4 (host)
    const int n_words = tile_bytes >> 2;
    const int n_out = groups * a.C;
    const int64_t row_bytes = (int64_t)a.W * a.C;     // bytes of an image row that exist
    const int64_t b0 = (int64_t)blockIdx.x * tile_bytes;
    const uint32_t fill = (uint32_t)a.border_const * 0x01010101u;
    // horizontal stage: nseg threads share one output
    int nseg = 1;
    while (nseg * 2 * n_out <= kRowsThreads && nseg * 2 <= r) nseg *= 2;
    const int seg_len = r / nseg;

    for (int rr = 0; rr < rows_per_cta; ++rr) {
        const int oy = blockIdx.y * rows_per_cta + rr;
        if (oy >= a.out_h) break;
        uint32_t accE[WPT], accO[WPT];
#pragma unroll
        for (int k = 0; k < WPT; ++k) accE[k] = accO[k] = 0u;
        for (int o = tid; o < n_out; o += kRowsThreads) outsum[o] = 0u;
#pragma unroll 4
        for (int dy = 0; dy < r; ++dy) {
            const int ym = border_index((oy << a.depth) + dy, a.H, a.border_type);       // uniform over the CTA
            const uint8_t* row = ym < 0 ? nullptr : a.src + (int64_t)ym * a.pitch;
#pragma unroll
            for (int k = 0; k < WPT; ++k) {
                const int w = tid + k * kRowsThreads;
                if (w >= n_words) break;
                const int64_t b = b0 + 4 * (int64_t)w;
                uint32_t v;
                if (row == nullptr) {
                    v = fill;
                } else if (b + 4 <= row_bytes) {
                    if (aligned) v = __ldg(reinterpret_cast<const uint32_t*>(row + b));
                    else v = (uint32_t)row[b] | ((uint32_t)row[b + 1] << 8) | ((uint32_t)row[b + 2] << 16) | ((uint32_t)row[b + 3] << 24);
                } else {
                    v = padded_byte(a, row, b) | (padded_byte(a, row, b + 1) << 8) | (padded_byte(a, row, b + 2) << 16) |
                        (padded_byte(a, row, b + 3) << 24);
                }
                accE[k] += prmt(v, 0u, 0x4240u);      // bytes 0 and 2 in 16-bit lanes
                accO[k] += prmt(v, 0u, 0x4341u);      // bytes 1 and 3
            }
        }
#pragma unroll
        for (int k = 0; k < WPT; ++k) {
            const int w = tid + k * kRowsThreads;
            if (w >= n_words) break;
            uint2 pk;
            pk.x = prmt(accE[k], accO[k], 0x5410u);   // (col 4w, col 4w+1)
            pk.y = prmt(accE[k], accO[k], 0x7632u);   // (col 4w+2, col 4w+3)
            *reinterpret_cast<uint2*>(&colsum[4 * w]) = pk;
        }
        __syncthreads();
        const int total = n_out * nseg;
        for (int i = tid; i < total; i += kRowsThreads) {
            const int o = i % n_out, seg = i / n_out;
            const int g = o / a.C, c = o - g * a.C;
            const uint16_t* p = colsum + g * gb + (seg * seg_len) * a.C + c;
            uint32_t s = 0;
            for (int q = 0; q < seg_len; ++q) s += p[q * a.C];
            if (nseg == 1) outsum[o] = s;
            else atomicAdd(&outsum[o], s);            // integer: exact and order-independent
        }
        __syncthreads();
        for (int o = tid; o < n_out; o += kRowsThreads) {
            const int g = o / a.C, c = o - g * a.C;
            const int ox = blockIdx.x * groups + g;
            if (ox >= a.out_w) continue;
            const uint32_t s = outsum[o];
            if (a.dst_u8 != nullptr) a.dst_u8[(int64_t)oy * a.dst_pitch + (int64_t)ox * a.C + c] = (uint8_t)(s >> (2 * a.depth));
            else a.dst_f32[((int64_t)oy * a.out_w + ox) * a.C + c] = __uint2float_rn(s) * (1.0f / (float)(1u << (2 * a.depth)));
        }
        __syncthreads();
    }
}

// ---- levels above 6 of a fused run --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) haar_tail_kernel(TailArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nb = 1 << (a.depth - 6);                  // level-6 blocks per output pixel and axis
    for (int64_t o = warp; o < (int64_t)a.out_h * a.out_w; o += n_warps) {
        const int oy = (int)(o / a.out_w), ox = (int)(o - (int64_t)oy * a.out_w);
        uint32_t s[3] = {0u, 0u, 0u};
        for (int j = 0; j < nb * nb; ++j) {
            const int by = oy * nb + j / nb, bx = ox * nb + j % nb;
            if (by < a.s6_h && bx < a.s6_w) {           // inside the depth-6 extents: the one-pass kernel's exact sums
                if (lane == 0) {
                    const uint32_t* p = a.sum6 + ((int64_t)by * a.s6_w + bx) * 3;
                    s[0] += p[0]; s[1] += p[1]; s[2] += p[2];
                }
                continue;
            }
            // a 64 x 64 block entirely in the padding of the deeper level: straight from the image
            for (int i = lane; i < 64 * 64; i += 32) {
                const int ym = border_index(by * 64 + (i >> 6), a.H, a.border_type);
                const int xm = border_index(bx * 64 + (i & 63), a.W, a.border_type);
                if (ym < 0 || xm < 0) { s[0] += a.border_const; s[1] += a.border_const; s[2] += a.border_const; continue; }
                const uint8_t* p = a.src + (int64_t)ym * a.pitch + (int64_t)xm * 3;
                s[0] += p[0]; s[1] += p[1]; s[2] += p[2];
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s[c] += __shfl_xor_sync(0xFFFFFFFFu, s[c], off);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (a.dst_u8 != nullptr) a.dst_u8[(int64_t)oy * a.dst_pitch + (int64_t)ox * 3 + c] = (uint8_t)(s[c] >> (2 * a.depth));
                else a.dst_f32[((int64_t)oy * a.out_w + ox) * 3 + c] = __uint2float_rn(s[c]) * (1.0f / (float)(1u << (2 * a.depth)));
            }
        }
    }
}

}  // namespace

// groups of whole output pixels per CTA tile so that a tile is about 1 KB (<= 8 KB) and a multiple of 4 bytes;
// 0 when one output pixel's row segment does not fit (absurd channel counts): the caller uses the scalar kernel
int rows_kernel_groups(int C, int depth) {
    const int64_t gb = (int64_t)C << depth;            // even, because depth >= 1
    if (depth < 1 || gb > kRowsMaxTileBytes) return 0;
    int g = (int)(1024 / gb);
    if (gb % 4 == 0) return g < 1 ? 1 : g;
    g &= ~1;                                           // depth 1, odd channel count: pairs of output pixels
    return g < 2 ? 2 : g;
}

cudaError_t launch_icon_rows(const GenericIconArgs& a, cudaStream_t stream) {
    if ((int64_t)a.out_h * a.out_w <= 0) return cudaSuccess;
    const int groups = rows_kernel_groups(a.C, a.depth);
    if (groups <= 0 || a.depth > 8) return cudaErrorInvalidValue;
    const int64_t tile_bytes = (int64_t)groups * ((int64_t)a.C << a.depth);
    const int wpt = (int)((tile_bytes / 4 + kRowsThreads - 1) / kRowsThreads);
    int rows_per_cta = 32 >> a.depth;
    if (rows_per_cta < 1) rows_per_cta = 1;
    const int aligned = (((uintptr_t)a.src & 3) == 0 && (a.pitch & 3) == 0) ? 1 : 0;
    while ((a.out_h + rows_per_cta - 1) / rows_per_cta > 65535) rows_per_cta *= 2;
    dim3 grid((unsigned)((a.out_w + groups - 1) / groups), (unsigned)((a.out_h + rows_per_cta - 1) / rows_per_cta));
    if (wpt <= 1) haar_icon_rows_kernel<1><<<grid, kRowsThreads, 0, stream>>>(a, groups, rows_per_cta, aligned);
    else if (wpt <= 2) haar_icon_rows_kernel<2><<<grid, kRowsThreads, 0, stream>>>(a, groups, rows_per_cta, aligned);
    else if (wpt <= 4) haar_icon_rows_kernel<4><<<grid, kRowsThreads, 0, stream>>>(a, groups, rows_per_cta, aligned);
    else haar_icon_rows_kernel<8><<<grid, kRowsThreads, 0, stream>>>(a, groups, rows_per_cta, aligned);
    return cudaGetLastError();
}

cudaError_t launch_icon_tail(const TailArgs& a, cudaStream_t stream) {
    const int64_t n = (int64_t)a.out_h * a.out_w;
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 7) / 8;                    // a warp per output pixel, 8 warps per CTA
    if (blocks > 148 * 8) blocks = 148 * 8;
    haar_tail_kernel<<<(int)blocks, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace wicca
