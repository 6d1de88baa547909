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
//                           depths 1..8 of an RGB image still cost ONE pass over it.
//   haar_tail_fill_kernel   the level-6 blocks that lie entirely in the padding of the deeper level (beyond the depth-6
//                           extents), summed directly from the image through the border index map, a CTA per block.
#include <cuda_runtime.h>
#include <stdint.h>

#include "haar_math.cuh"
#include "icon_types.h"
#include "kernels.h"

namespace wicca {

namespace {

constexpr int kRowsThreads = 256;
constexpr int kRowsMaxTileBytes = 8192;         // 8 words per thread

// Source byte offset inside an image row of byte b of the border-extended row (b >= W*C is where it matters);
// -1 when the border rule supplies the constant.  32-bit arithmetic: a padded row is far below 2 GB (launcher).
__device__ __forceinline__ int padded_offset(const GenericIconArgs& a, int b) {
    const int x = b / a.C;
    const int c = b - x * a.C;
    const int xm = border_index(x, a.W, a.border_type);
    return xm < 0 ? -1 : xm * a.C + c;
}

// WPT words per thread and row; U = 8 / WPT rows are fetched before the first of them is consumed, so every thread keeps
// eight independent 32-bit loads in flight (a row loop that consumes each word as it arrives serialises on the HBM
// latency: 0.9 TB/s measured; the batched form is what makes the kernel bandwidth bound).
// BORDER = false: the CTAs whose whole tile lies inside the image (the streaming loop with no selects, few registers);
// BORDER = true: the others (last tile column, last row block, unaligned sources).  Both are launched over the same
// grid and each CTA returns at once from the instantiation that is not its own.
template <int WPT, bool BORDER>
__global__ void __launch_bounds__(kRowsThreads)
haar_icon_rows_kernel(GenericIconArgs a, int groups, int rows_per_cta, int aligned) {
    constexpr int U = WPT >= 8 ? 1 : 8 / WPT;
    __shared__ __align__(16) uint16_t colsum[kRowsMaxTileBytes];
    __shared__ uint32_t outsum[kRowsThreads / 2];   // used only when nseg > 1, i.e. n_out <= 128
    const int tid = threadIdx.x;
    const int r = 1 << a.depth;
    const int gb = a.C << a.depth;                    // bytes of one output pixel's input row segment
    const int tile_bytes = groups * gb;               // multiple of 4 (host)
    const int n_words = tile_bytes >> 2;
    const int n_out = groups * a.C;
    const int64_t row_bytes = (int64_t)a.W * a.C;     // bytes of an image row that exist
    const int64_t padded_bytes = ((int64_t)a.out_w << a.depth) * a.C;   // bytes of a border-extended row that feed an output
    const int64_t b0 = (int64_t)blockIdx.x * tile_bytes;
    const uint32_t fill = (uint32_t)a.border_const * 0x01010101u;
    const bool interior = aligned && (b0 + tile_bytes <= row_bytes) && (r % U) == 0 &&
                          ((((int64_t)blockIdx.y + 1) * rows_per_cta) << a.depth) <= a.H;    // CTA-uniform: no border in this tile
    if (interior == BORDER) return;
    // horizontal stage: nseg threads share one output when there are few outputs and many pixels per output
    int nseg = 1;
    while (nseg * 2 * n_out <= kRowsThreads && nseg * 2 <= r) nseg *= 2;
    const int seg_len = r / nseg;
    const int64_t ob0 = (int64_t)blockIdx.x * n_out;  // first output byte (of the icon row) of this tile
    const int64_t out_row_bytes = (int64_t)a.out_w * a.C;
    // Where this thread's outputs live in the column sums.  The divisions by the (run-time) channel count were most of
    // the kernel's instructions: the few-outputs case (nseg > 1) resolves its one (output, segment) here, once per CTA;
    // the many-outputs case divides by multiplication (exact for the 13-bit indices of a tile) instead of keeping
    // sixteen offsets in registers - at 52 registers the kernel lost more to occupancy than the divisions had cost.
    const uint32_t inv_c = a.C > 1 ? 0xFFFFFFFFu / (uint32_t)a.C + 1u : 0u;
    int hs_off = -1, hs_o = 0;
    if (nseg > 1 && tid < n_out * nseg) {
        const int o = tid % n_out, seg = tid / n_out;
        const int g = o / a.C, c = o - g * a.C;
        hs_off = g * gb + (seg * seg_len) * a.C + c;
        hs_o = o;
    }

    for (int rr = 0; rr < rows_per_cta; ++rr) {
        const int oy = blockIdx.y * rows_per_cta + rr;
        if (oy >= a.out_h) break;
        uint32_t accE[WPT], accO[WPT];
#pragma unroll
        for (int k = 0; k < WPT; ++k) accE[k] = accO[k] = 0u;
        if (nseg > 1)
            for (int o = tid; o < n_out; o += kRowsThreads) outsum[o] = 0u;
        if (!BORDER) {
            // no border anywhere in this tile: a plain strided stream, U x WPT loads issued before the first use
            const uint32_t* p = reinterpret_cast<const uint32_t*>(a.src + ((int64_t)oy << a.depth) * a.pitch + b0) + tid;
            const int64_t pw = a.pitch >> 2;
            for (int dy0 = 0; dy0 < r; dy0 += U) {
                uint32_t v[U][WPT];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int k = 0; k < WPT; ++k)
                        v[u][k] = (tid + k * kRowsThreads < n_words) ? __ldg(p + (int64_t)(dy0 + u) * pw + k * kRowsThreads) : 0u;
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int k = 0; k < WPT; ++k) {
                        accE[k] += prmt(v[u][k], 0u, 0x4240u);      // bytes 0 and 2 in 16-bit lanes
                        accO[k] += prmt(v[u][k], 0u, 0x4341u);      // bytes 1 and 3
                    }
            }
        } else {
            // A tile that touches the bottom and / or the right border.  The rows of this output row that exist in
            // the image are streamed exactly like an interior tile (only the words that lie inside the image
            // horizontally); the rows below the image follow with the border row map applied by selects (no branches
            // between the loads); the few words that touch the right border are walked word-major further down.
            int n_fast = 0;
            if (aligned && (r % U) == 0) {
                const int64_t left = (int64_t)a.H - ((int64_t)oy << a.depth);      // image rows at and below this output row's first
                n_fast = left >= r ? r : (left <= 0 ? 0 : (int)left / U * U);
                const uint32_t* p = reinterpret_cast<const uint32_t*>(a.src + ((int64_t)oy << a.depth) * a.pitch + b0) + tid;
                const int64_t pw = a.pitch >> 2;
                for (int dy0 = 0; dy0 < n_fast; dy0 += U) {
                    uint32_t v[U][WPT];
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int k = 0; k < WPT; ++k) {
                            const int w = tid + k * kRowsThreads;
                            const bool in_image = w < n_words && b0 + 4 * (int64_t)w + 4 <= row_bytes;
                            v[u][k] = in_image ? __ldg(p + (int64_t)(dy0 + u) * pw + k * kRowsThreads) : 0u;
                        }
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int k = 0; k < WPT; ++k) {
                            accE[k] += prmt(v[u][k], 0u, 0x4240u);
                            accO[k] += prmt(v[u][k], 0u, 0x4341u);
                        }
                }
            }
            for (int dy0 = n_fast; dy0 < r; dy0 += U) {
                int ymv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) ymv[u] = dy0 + u < r ? border_index((oy << a.depth) + dy0 + u, a.H, a.border_type) : -2;
                uint32_t v[U][WPT];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint8_t* row = a.src + (int64_t)(ymv[u] < 0 ? 0 : ymv[u]) * a.pitch;
#pragma unroll
                    for (int k = 0; k < WPT; ++k) {
                        const int64_t b = b0 + 4 * (int64_t)(tid + k * kRowsThreads);
                        const bool in_image = tid + k * kRowsThreads < n_words && b + 4 <= row_bytes;
                        const int64_t bs = in_image ? b : 0;
                        uint32_t x;
                        if (aligned) x = __ldg(reinterpret_cast<const uint32_t*>(row + bs));
                        else x = (uint32_t)row[bs] | ((uint32_t)row[bs + 1] << 8) | ((uint32_t)row[bs + 2] << 16) | ((uint32_t)row[bs + 3] << 24);
                        v[u][k] = (!in_image || ymv[u] == -2) ? 0u : (ymv[u] < 0 ? fill : x);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int k = 0; k < WPT; ++k) {
                        accE[k] += prmt(v[u][k], 0u, 0x4240u);
                        accO[k] += prmt(v[u][k], 0u, 0x4341u);
                    }
            }
            if (b0 + tile_bytes > row_bytes) {
#pragma unroll
                for (int k = 0; k < WPT; ++k) {
                    const int64_t b = b0 + 4 * (int64_t)(tid + k * kRowsThreads);
                    if (tid + k * kRowsThreads >= n_words || b + 4 <= row_bytes || b >= padded_bytes) continue;
                    int off[4];                       // source byte of each of the word's four columns (-1: the constant)
#pragma unroll
                    for (int jb = 0; jb < 4; ++jb) off[jb] = padded_offset(a, (int)b + jb);
                    for (int dy0 = 0; dy0 < r; dy0 += 8) {
                        uint32_t by[8][4];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int ym = dy0 + u < r ? border_index((oy << a.depth) + dy0 + u, a.H, a.border_type) : -2;
                            const uint8_t* row = a.src + (int64_t)(ym < 0 ? 0 : ym) * a.pitch;
#pragma unroll
                            for (int jb = 0; jb < 4; ++jb) {
                                const uint32_t x = row[off[jb] < 0 ? 0 : off[jb]];
                                by[u][jb] = ym == -2 ? 0u : ((ym < 0 || off[jb] < 0) ? (uint32_t)a.border_const : x);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            accE[k] += by[u][0] | (by[u][2] << 16);
                            accO[k] += by[u][1] | (by[u][3] << 16);
                        }
                    }
                }
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
        if (nseg > 1) {
            if (hs_off >= 0) {
                const uint16_t* p = colsum + hs_off;
                uint32_t s = 0;
                for (int q = 0; q < seg_len; ++q) s += p[q * a.C];
                atomicAdd(&outsum[hs_o], s);          // integer: exact and order-independent
            }
            __syncthreads();
        }
        // four consecutive output bytes (or floats) per thread and group
#pragma unroll 1
        for (int o4 = 4 * tid; o4 < n_out; o4 += 4 * kRowsThreads) {
            uint32_t sv[4];
            int g = a.C > 1 ? (int)__umulhi((uint32_t)o4, inv_c) : o4, c = o4 - g * a.C;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sv[j] = 0u;
                if (o4 + j < n_out) {
                    if (nseg > 1) {
                        sv[j] = outsum[o4 + j];
                    } else {
                        const uint16_t* p = colsum + g * gb + c;
                        uint32_t s = 0;
#pragma unroll 2
                        for (int q = 0; q < r; ++q) s += p[q * a.C];
                        sv[j] = s;
                    }
                }
                if (++c == a.C) { c = 0; ++g; }
            }
            const int64_t ob = ob0 + o4;
            if (a.dst_u8 != nullptr) {
                uint8_t* q = a.dst_u8 + (int64_t)oy * a.dst_pitch + ob;
                const uint32_t word = (sv[0] >> (2 * a.depth)) | ((sv[1] >> (2 * a.depth)) << 8) | ((sv[2] >> (2 * a.depth)) << 16) |
                                      ((sv[3] >> (2 * a.depth)) << 24);
                if (o4 + 4 <= n_out && ob + 4 <= out_row_bytes && ((uintptr_t)q & 3) == 0) {
                    *reinterpret_cast<uint32_t*>(q) = word;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (o4 + j < n_out && ob + j < out_row_bytes) q[j] = (uint8_t)(word >> (8 * j));
                }
            } else {
                const float scale = 1.0f / (float)(1u << (2 * a.depth));
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (o4 + j < n_out && ob + j < out_row_bytes) a.dst_f32[(int64_t)oy * out_row_bytes + ob + j] = __uint2float_rn(sv[j]) * scale;
            }
        }
        __syncthreads();
    }
}

// ---- levels above 6 of a fused run --------------------------------------------------------------------------
// Level-6 blocks of the deeper level's padded grid that the one-pass kernel did not produce (they lie entirely in the
// padding beyond the depth-6 extents): a CTA per block, straight from the image through the border index map.
__global__ void __launch_bounds__(256) haar_tail_fill_kernel(TailArgs a) {
    __shared__ uint32_t tot[3];
    const int right_w = a.ext_w - a.s6_w;                               // columns right of the valid region
    const int64_t n_right = (int64_t)right_w * a.s6_h;
    const int64_t n_missing = n_right + (int64_t)(a.ext_h - a.s6_h) * a.ext_w;
    for (int64_t m = blockIdx.x; m < n_missing; m += gridDim.x) {       // a CTA per block: 16 pixels per thread
        int by, bx;
        if (m < n_right) { by = (int)(m / right_w); bx = a.s6_w + (int)(m - (int64_t)by * right_w); }
        else { const int64_t q = m - n_right; by = a.s6_h + (int)(q / a.ext_w); bx = (int)(q - (int64_t)(by - a.s6_h) * a.ext_w); }
        if (threadIdx.x < 3) tot[threadIdx.x] = 0u;
        __syncthreads();
        uint32_t s[3] = {0u, 0u, 0u};
        const int xm = border_index(bx * 64 + (threadIdx.x & 63), a.W, a.border_type);
#pragma unroll 4
        for (int i = threadIdx.x >> 6; i < 64; i += 4) {
            const int ym = border_index(by * 64 + i, a.H, a.border_type);
            if (ym < 0 || xm < 0) { s[0] += a.border_const; s[1] += a.border_const; s[2] += a.border_const; continue; }
            const uint8_t* p = a.src + (int64_t)ym * a.pitch + (int64_t)xm * 3;
            s[0] += p[0]; s[1] += p[1]; s[2] += p[2];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s[c] += __shfl_xor_sync(0xFFFFFFFFu, s[c], off);
            if ((threadIdx.x & 31) == 0) atomicAdd(&tot[c], s[c]);      // integer: exact, order-independent
        }
        __syncthreads();
        if (threadIdx.x < 3) a.sum6[((int64_t)by * a.ext_w + bx) * 3 + threadIdx.x] = tot[threadIdx.x];
        __syncthreads();
    }
}

// One thread per output element: the 2^(d-6) x 2^(d-6) level-6 block sums of its pixel, shifted once.
__global__ void __launch_bounds__(256) haar_tail_kernel(TailArgs a) {
    const int nb = 1 << (a.depth - 6);
    const int64_t n = (int64_t)a.out_h * a.out_w * 3;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % 3);
        const int64_t t = i / 3;
        const int ox = (int)(t % a.out_w), oy = (int)(t / a.out_w);
        const uint32_t* p = a.sum6 + ((int64_t)oy * nb * a.ext_w + (int64_t)ox * nb) * 3 + c;
        uint32_t s = 0;
        for (int y = 0; y < nb; ++y)
            for (int x = 0; x < nb; ++x) s += p[((int64_t)y * a.ext_w + x) * 3];
        if (a.dst_u8 != nullptr) a.dst_u8[(int64_t)oy * a.dst_pitch + (int64_t)ox * 3 + c] = (uint8_t)(s >> (2 * a.depth));
        else a.dst_f32[i] = __uint2float_rn(s) * (1.0f / (float)(1u << (2 * a.depth)));
    }
}

}  // namespace

// groups of whole output pixels per CTA tile so that a tile is about 1 KB (<= 8 KB) and a multiple of 4 bytes;
// 0 when one output pixel's row segment does not fit (absurd channel counts): the caller uses the scalar kernel
int rows_kernel_groups(int C, int depth) {
    const int64_t gb = (int64_t)C << depth;            // even, because depth >= 1
    if (depth < 1 || gb > kRowsMaxTileBytes) return 0;
    // eight loads in flight per thread: 1 KB tiles when an output row has >= 8 input rows, wider tiles below that
    const int target = depth >= 3 ? 1024 : (depth == 2 ? 2048 : 4096);
    int g = (int)(target / gb);
    if (gb % 4 == 0) return g < 1 ? 1 : g;
    g &= ~1;                                           // depth 1, odd channel count: pairs of output pixels
    return g < 2 ? 2 : g;
}

// One non-blocking side stream per device and host thread, created on first use (nullptr if that fails: the border
// instantiation then simply follows the interior one on the caller's stream).
static cudaStream_t rows_side_stream() {
    static thread_local cudaStream_t streams[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return nullptr; }
    if (streams[dev] == nullptr && cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        streams[dev] = nullptr;
    }
    return streams[dev];
}

cudaError_t launch_icon_rows(const GenericIconArgs& a, cudaStream_t stream) {
    if ((int64_t)a.out_h * a.out_w <= 0) return cudaSuccess;
    const int groups = rows_kernel_groups(a.C, a.depth);
    if (groups <= 0 || a.depth > 8 || (((int64_t)a.out_w << a.depth) + 8) * a.C >= 0x7FFFFFFF) return cudaErrorInvalidValue;
    const int64_t tile_bytes = (int64_t)groups * ((int64_t)a.C << a.depth);
    const int wpt = (int)((tile_bytes / 4 + kRowsThreads - 1) / kRowsThreads);
    int rows_per_cta = 32 >> a.depth;
    if (rows_per_cta < 1) rows_per_cta = 1;
    const int aligned = (((uintptr_t)a.src & 3) == 0 && (a.pitch & 3) == 0) ? 1 : 0;
    while ((a.out_h + rows_per_cta - 1) / rows_per_cta > 65535) rows_per_cta *= 2;
    dim3 grid((unsigned)((a.out_w + groups - 1) / groups), (unsigned)((a.out_h + rows_per_cta - 1) / rows_per_cta));
    // The border instantiation is a few dozen long-running CTAs (one output row's 2^d input rows each): it runs next to
    // the interior one on a side stream, forked from and joined back into the caller's stream with events (the pattern
    // is capture-safe), instead of after it.
    cudaStream_t side = rows_side_stream();
    cudaEvent_t fork = nullptr, join = nullptr;
    if (side != nullptr) {
        if (cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&join, cudaEventDisableTiming) != cudaSuccess) {
            if (fork) cudaEventDestroy(fork);
            cudaGetLastError();
            fork = join = nullptr;
            side = nullptr;
        }
    }
    cudaStream_t bstream = side ? side : stream;
    if (side) { cudaEventRecord(fork, stream); cudaStreamWaitEvent(side, fork, 0); }
#define WICCA_ROWS_LAUNCH(N)                                                                                              \
    do {                                                                                                                  \
        haar_icon_rows_kernel<N, true><<<grid, kRowsThreads, 0, bstream>>>(a, groups, rows_per_cta, aligned);                 \
        haar_icon_rows_kernel<N, false><<<grid, kRowsThreads, 0, stream>>>(a, groups, rows_per_cta, aligned);                 \
    } while (0)
    if (wpt <= 1) WICCA_ROWS_LAUNCH(1);
    else if (wpt <= 2) WICCA_ROWS_LAUNCH(2);
    else if (wpt <= 4) WICCA_ROWS_LAUNCH(4);
    else WICCA_ROWS_LAUNCH(8);
#undef WICCA_ROWS_LAUNCH
    const cudaError_t launched = cudaGetLastError();
    if (side) {
        cudaEventRecord(join, side);
        cudaStreamWaitEvent(stream, join, 0);
        cudaEventDestroy(fork);                       // destruction is deferred until the events have completed
        cudaEventDestroy(join);
    }
    if (launched != cudaSuccess) return launched;
    return cudaGetLastError();
}

cudaError_t launch_icon_tail_fill(const TailArgs& a, cudaStream_t stream) {
    const int64_t missing = (int64_t)a.ext_h * a.ext_w - (int64_t)a.s6_h * a.s6_w;
    if (missing <= 0) return cudaSuccess;
    int64_t blocks = missing;                        // a CTA per missing block
    if (blocks > 148 * 8) blocks = 148 * 8;
    haar_tail_fill_kernel<<<(int)blocks, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_icon_tail(const TailArgs& a, cudaStream_t stream) {
    const int64_t n = (int64_t)a.out_h * a.out_w * 3;
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    haar_tail_kernel<<<(int)blocks, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace wicca
