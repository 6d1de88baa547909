// wavelet_fir.cu - LL pyramid of an orthogonal wavelet with longer filters (Daubechies, Coiflet) behind the
// WaveletCoder interface (row N4; the reference lists them as its roadmap, README.md:25 and :222, and implements
// only Haar).  Definition in oracle/fir_oracle.py: pad bottom/right to a multiple of 2^depth like get_padded_copy,
// then per level a separable low-pass with periodic wrap-around, taps g = dec_lo / sqrt(2), every product and sum
// rounded to float32 separately (no FMA) so that the CPU restatement and these kernels agree bit for bit; with the
// taps [1/2, 1/2] the result is the reference's Haar icon.  One pass along rows and one along columns per level.
#include <cuda_runtime.h>
#include <stdint.h>

#include "haar_math.cuh"
#include "kernels.h"

namespace wicca {

namespace {

// rows: t[y, j, ch] = sum_n g[n] * x[y, (2 j + n - c) mod w, ch]; x is the uint8 source seen through the border rule
// (level 1) or the previous float32 plane
template <bool kFromU8>
__global__ void __launch_bounds__(256)
fir_rows_kernel(const uint8_t* __restrict__ src, int64_t pitch, int H, int W, int border_type, int border_const,
                const float* __restrict__ in, int h, int w, int C, FirTaps taps, float* __restrict__ t) {
    const int w2 = w >> 1;
    const int64_t total = (int64_t)h * w2 * C;
    const float fc = (float)border_const;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % C);
        const int64_t r = i / C;
        const int j = (int)(r % w2), y = (int)(r / w2);
        const int ym = kFromU8 ? (y < H ? y : border_index(y, H, border_type)) : y;
        const int x0 = 2 * j - taps.c;
        float acc = 0.0f;
        if (x0 >= 0 && x0 + taps.n <= (kFromU8 ? W : w) && ym >= 0) {
            // interior: no wrap-around, no border rule - the taps walk over consecutive pixels
            if (kFromU8) {
                const uint8_t* p = src + (int64_t)ym * pitch + (int64_t)x0 * C + ch;
                for (int n = 0; n < taps.n; ++n) {
                    const float q = __fmul_rn((float)p[n * C], taps.g[n]);
                    acc = n == 0 ? q : __fadd_rn(acc, q);
                }
            } else {
                const float* p = in + ((int64_t)y * w + x0) * C + ch;
                for (int n = 0; n < taps.n; ++n) {
                    const float q = __fmul_rn(p[n * C], taps.g[n]);
                    acc = n == 0 ? q : __fadd_rn(acc, q);
                }
            }
        } else {
            for (int n = 0; n < taps.n; ++n) {
                int x = (x0 + n) % w;                   // deep levels can be narrower than the filter: a true modulo
                x += x < 0 ? w : 0;
                float v;
                if (kFromU8) {
                    const int xm = x < W ? x : border_index(x, W, border_type);
                    v = (ym < 0 || xm < 0) ? fc : (float)src[(int64_t)ym * pitch + (int64_t)xm * C + ch];
                } else {
                    v = in[((int64_t)y * w + x) * C + ch];
                }
                const float q = __fmul_rn(v, taps.g[n]);
                acc = n == 0 ? q : __fadd_rn(acc, q);
            }
        }
        t[i] = acc;
    }
}

// columns: ll[i, j, ch] = sum_m g[m] * t[(2 i + m - c) mod h, j, ch]; the last level clips and truncates to uint8
__global__ void __launch_bounds__(256)
fir_cols_kernel(const float* __restrict__ t, int h, int w2, int C, FirTaps taps, float* __restrict__ out,
                uint8_t* __restrict__ icon, int64_t icon_pitch) {
    const int h2 = h >> 1;
    const int64_t row = (int64_t)w2 * C;
    const int64_t total = (int64_t)h2 * row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i % row;
        const int y2 = (int)(i / row);
        float acc = 0.0f;
        const int y0 = 2 * y2 - taps.c;
        if (y0 >= 0 && y0 + taps.n <= h) {
            const float* p = t + (int64_t)y0 * row + e;
            for (int m = 0; m < taps.n; ++m) {
                const float q = __fmul_rn(p[(int64_t)m * row], taps.g[m]);
                acc = m == 0 ? q : __fadd_rn(acc, q);
            }
        } else {
            for (int m = 0; m < taps.n; ++m) {
                int y = (y0 + m) % h;
                y += y < 0 ? h : 0;
                const float q = __fmul_rn(t[(int64_t)y * row + e], taps.g[m]);
                acc = m == 0 ? q : __fadd_rn(acc, q);
            }
        }
        if (icon) icon[(int64_t)y2 * icon_pitch + e] = (uint8_t)fminf(fmaxf(acc, 0.0f), 255.0f);      // clip, then truncate
        else out[i] = acc;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Tiled level kernel (C <= 4): a CTA produces kTileH x kTileW low-pass samples.
//   stage   the (2*kTileH + L - 2) x (2*kTileW + L - 2) input pixels the tile needs go to shared memory; the
//           periodic wrap-around and (level 1) the border rule are resolved here, once per pixel;
//   rows    each thread filters one staged row for four neighbouring outputs and all channels: the 2*4 + L - 2
//           pixels it loads are shared by the four outputs, the sums run tap by tap as in the oracle;
//   cols    each thread filters one output element down the column of row results.
// ------------------------------------------------------------------------------------------------------------
constexpr int kTileThreads = 256, kOutPerThread = 4;

// Tile geometry by input type: float planes (levels >= 2) take half the width, so that four CTAs fit an SM.
template <typename TIn, int C, int L>
struct FirTile {
    static constexpr int kW = sizeof(TIn) == 1 ? 64 : 32;                   // low-pass samples per tile row
    static constexpr int kH = sizeof(TIn) == 1 ? 16 : 14;                   // tile rows
    static constexpr int rows = 2 * kH + L - 2, cols = 2 * kW + L - 2;      // input window
    // staged row: up to 15 bytes of alignment slack, the window, rounded up to whole 16-byte chunks; float rows are
    // padded to a pitch of 4 words mod 8 (a warp of the row pass reads 4 rows x 8 runs 24 words apart: 8-way bank
    // conflicts on a pitch of 0 mod 8 words, 4-way - the best a multiple of 16 bytes allows - on 4 mod 8)
    static constexpr int kRowBytesMin = (15 + cols * C * (int)sizeof(TIn) + 15) / 16 * 16;
    static constexpr int kRowBytes = (sizeof(TIn) == 1 || kRowBytesMin % 32 == 16) ? kRowBytesMin : kRowBytesMin + 16;
    static constexpr size_t kInBytes = (size_t)rows * kRowBytes;
    static constexpr int kTPitch = kW * C + 4;                               // row results: rows start 4 banks apart, 16-byte stores stay aligned
    static constexpr size_t kSmem = kInBytes + (size_t)rows * kTPitch * sizeof(float);
};

template <int CB>
__device__ __forceinline__ void fir_cp_async(uint32_t dst, const void* src) {
    if (CB == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else if (CB == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

// All rows of an interior window, as CB-byte chunks from the aligned-down start of each row: fire and forget.
template <int CB>
__device__ __forceinline__ void fir_fetch_window(unsigned char* s_raw, int row_pitch_smem, const unsigned char* g_first,
                                                 int64_t g_pitch_bytes, int rows, int chunks) {
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(s_raw);
    for (int q = threadIdx.x; q < rows * chunks; q += kTileThreads) {
        const int r = q / chunks, ch = q - r * chunks;
        fir_cp_async<CB>(s0 + (uint32_t)(r * row_pitch_smem + ch * CB), g_first + (int64_t)r * g_pitch_bytes + ch * CB);
    }
}

template <typename TIn, int C, int L>
__global__ void __launch_bounds__(kTileThreads)
fir_tile_kernel(const TIn* __restrict__ src, int64_t pitch_elems, int H, int W, int border_type, int border_const, int h, int w,
                FirTaps taps, float* __restrict__ out, uint8_t* __restrict__ icon, int64_t icon_pitch) {
    using G = FirTile<TIn, C, L>;
    constexpr int kTileW = G::kW, kTileH = G::kH, rows = G::rows, cols = G::cols, kRowBytes = G::kRowBytes;
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ float s_g[16];
    if (threadIdx.x < 16) s_g[threadIdx.x] = taps.g[threadIdx.x];
    float* s_t = reinterpret_cast<float*>(s_raw + G::kInBytes);                      // [rows][kTPitch]
    const int h2 = h >> 1, w2 = w >> 1;
    const int oy0 = blockIdx.y * kTileH, ox0 = blockIdx.x * kTileW;
    constexpr bool kU8 = sizeof(TIn) == 1;
    // ---- stage: row r of the window lives at s_raw + r * kRowBytes + lead
    const int y0 = 2 * oy0 - taps.c, x0 = 2 * ox0 - taps.c;
    bool interior = y0 >= 0 && x0 >= 0 && y0 + rows <= (kU8 ? min(h, H) : h) && x0 + cols <= (kU8 ? min(w, W) : w);
    int lead = 0;
    if (interior) {
        // no wrap-around and no border rule anywhere in the window: whole rows are contiguous in the source and go to
        // shared memory as aligned chunks (cp.async: every copy of the tile is in flight before the first wait)
        const int64_t pitch_bytes = pitch_elems * (int64_t)sizeof(TIn);
        const uintptr_t al = (uintptr_t)src | (uintptr_t)pitch_bytes;
        const int cb = (al & 15) == 0 ? 16 : (al & 7) == 0 ? 8 : (al & 3) == 0 ? 4 : 0;
        const int64_t first = (int64_t)x0 * C * (int64_t)sizeof(TIn);               // byte offset of the window in its row
        if (cb) {
            lead = (int)(first & (cb - 1));
            const int chunks = (lead + cols * C * (int)sizeof(TIn) + cb - 1) / cb;
            // the rounded-up chunk range must stay inside the row (pitch padding included)
            if (first - lead + (int64_t)chunks * cb <= pitch_bytes) {
                const unsigned char* g0 = reinterpret_cast<const unsigned char*>(src) + (int64_t)y0 * pitch_bytes + (first - lead);
                if (cb == 16) fir_fetch_window<16>(s_raw, kRowBytes, g0, pitch_bytes, rows, chunks);
                else if (cb == 8) fir_fetch_window<8>(s_raw, kRowBytes, g0, pitch_bytes, rows, chunks);
                else fir_fetch_window<4>(s_raw, kRowBytes, g0, pitch_bytes, rows, chunks);
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            } else {
                interior = false; lead = 0;
            }
        } else {
            interior = false;
        }
    }
    if (!interior) {
        for (int e = threadIdx.x; e < rows * cols; e += kTileThreads) {
            const int r = e / cols, k = e - r * cols;
            int y = (y0 + r) % h;
            y += y < 0 ? h : 0;
            int x = (x0 + k) % w;
            x += x < 0 ? w : 0;
            const int ym = kU8 ? (y < H ? y : border_index(y, H, border_type)) : y;
            const int xm = kU8 ? (x < W ? x : border_index(x, W, border_type)) : x;
            TIn* d = reinterpret_cast<TIn*>(s_raw + (size_t)r * kRowBytes) + (size_t)k * C;
            if (ym < 0 || xm < 0) {
#pragma unroll
                for (int ch = 0; ch < C; ++ch) d[ch] = (TIn)border_const;
            } else {
                const TIn* p = src + (int64_t)ym * pitch_elems + (int64_t)xm * C;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) d[ch] = p[ch];
            }
        }
    }
    __syncthreads();
    // ---- rows: (row r, outputs 4 jg .. 4 jg + 3, all channels) per work item
    constexpr int groups = kTileW / kOutPerThread;
    for (int e = threadIdx.x; e < rows * groups; e += kTileThreads) {
        const int r = e / groups, jg = e - r * groups;
        const TIn* p = reinterpret_cast<const TIn*>(s_raw + (size_t)r * kRowBytes + lead) + (size_t)(2 * kOutPerThread * jg) * C;
        // the 2*4 + L - 2 pixels the four outputs share, fetched once
        constexpr int kPix = 2 * kOutPerThread + L - 2;
        float px[kPix][C];
        if (kU8) {
            // Bytes: the run is read as aligned 32-bit words and byte k becomes the float 2^23 + byte with one PRMT (no
            // LDS.U8, no I2F - the conversion pipe runs at a quarter of the rate and was as busy as the issue port).  The
            // product is then fma(2^23 + byte, g, -2^23 * g): (2^23 + byte) * g is exact inside the fma and 2^23 * g is
            // exact, so the one rounding is that of byte * g - bit-identical to __fmul_rn((float)byte, g).
            constexpr int kBytes = kPix * C, kWords = ((kBytes - 1) >> 2) + 2;      // words that can hold a byte of the run
            const uintptr_t a = reinterpret_cast<uintptr_t>(p);
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
            const uint32_t sh = (uint32_t)(a & 3) * 8;
            uint32_t wd[kWords];
#pragma unroll
            for (int k = 0; k < kWords; ++k) wd[k] = wp[k];
#pragma unroll
            for (int b = 0; b < kBytes; ++b) {
                const uint32_t x = __funnelshift_r(wd[b >> 2], wd[(b >> 2) + 1], sh);          // hoisted: one per word
                px[b / C][b % C] = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7440 | (b & 3)));
            }
        } else {
#pragma unroll
            for (int k = 0; k < kPix; ++k)
#pragma unroll
                for (int ch = 0; ch < C; ++ch) px[k][ch] = (float)p[k * C + ch];
        }
        float acc[kOutPerThread][C];
#pragma unroll
        for (int n = 0; n < L; ++n) {
            const float g = s_g[n], ng = -8388608.0f * g;
#pragma unroll
            for (int o = 0; o < kOutPerThread; ++o)
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    const float v = kU8 ? __fmaf_rn(px[2 * o + n][ch], g, ng) : __fmul_rn(px[2 * o + n][ch], g);
                    acc[o][ch] = n == 0 ? v : __fadd_rn(acc[o][ch], v);
                }
        }
        float* d = s_t + (size_t)r * G::kTPitch + kOutPerThread * jg * C;
#pragma unroll
        for (int o = 0; o < kOutPerThread; ++o)
#pragma unroll
            for (int ch = 0; ch < C; ++ch) d[o * C + ch] = acc[o][ch];
    }
    __syncthreads();
    // ---- columns: two neighbouring elements of an output row per work item - an 8-byte shared load feeds the pair, the
    // two products are scalar (mul.rn.f32) and the running sums one packed add.rn.f32x2 (two IEEE additions per issue slot,
    // each half bit-identical to the scalar instruction).  The products must NOT be a mul.rn.f32x2: ptxas 12.9 contracts
    // mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 in spite of the .rn modifiers and of --fmad=false (seen in the SASS; also
    // when the product is written fma(v, g, -0)), which skips the rounding of the product - a handful of icon bytes per
    // million then differ from the definition (found by tools/bench_rows.py on a 1024 x 1024 image).
    constexpr int seg = kTileW * C, seg2 = seg / 2;                              // kTileW is even
    for (int e = threadIdx.x; e < kTileH * seg2; e += kTileThreads) {
        const int i = e / seg2, q = 2 * (e - i * seg2);
        const int oi = oy0 + i;
        if (oi >= h2) continue;
        const float* p = s_t + (size_t)(2 * i) * G::kTPitch + q;               // 8-byte aligned: kTPitch and q are even
        uint64_t acc = 0;
#pragma unroll
        for (int m = 0; m < L; ++m) {
            const float2 v = *reinterpret_cast<const float2*>(p + m * G::kTPitch);
            const float g = s_g[m];
            uint64_t pr;
            asm("mov.b64 %0, {%1, %2};" : "=l"(pr) : "f"(__fmul_rn(v.x, g)), "f"(__fmul_rn(v.y, g)));
            if (m == 0) acc = pr;
            else asm("add.rn.f32x2 %0, %1, %2;" : "=l"(acc) : "l"(acc), "l"(pr));
        }
        float r0, r1;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(acc));
        const bool in0 = ox0 + q / C < w2, in1 = ox0 + (q + 1) / C < w2;
        const int64_t o = (int64_t)ox0 * C + q;
        if (icon) {
            uint8_t* d = icon + (int64_t)oi * icon_pitch + o;
            if (in0) d[0] = (uint8_t)fminf(fmaxf(r0, 0.0f), 255.0f);          // clip, then truncate
            if (in1) d[1] = (uint8_t)fminf(fmaxf(r1, 0.0f), 255.0f);
        } else {
            float* d = out + (int64_t)oi * w2 * C + o;
            if (in0) d[0] = r0;
            if (in1) d[1] = r1;
        }
    }
}

template <typename TIn, int C, int L>
cudaError_t launch_tile_l(const TIn* src, int64_t pitch_elems, int H, int W, int border_type, int border_const, int h, int w,
                          const FirTaps& taps, float* out, uint8_t* icon, int64_t icon_pitch, cudaStream_t stream) {
    using G = FirTile<TIn, C, L>;
    constexpr size_t smem = G::kSmem;
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && configured_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(fir_tile_kernel<TIn, C, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured_dev = dev;
    }
    const dim3 grid((w / 2 + G::kW - 1) / G::kW, (h / 2 + G::kH - 1) / G::kH);
    fir_tile_kernel<TIn, C, L><<<grid, kTileThreads, smem, stream>>>(src, pitch_elems, H, W, border_type, border_const, h, w, taps,
                                                                    out, icon, icon_pitch);
    return cudaGetLastError();
}

// The filter length is a template parameter (the pixel window lives in registers): 2, 4, 6 and 8 taps are
// instantiated - Haar, db2, db3 / coif1, db4; longer filters take the per-element kernels below.
constexpr bool tiled_taps(int n) { return n == 2 || n == 4 || n == 6 || n == 8; }

template <typename TIn, int C>
cudaError_t launch_tile(const TIn* src, int64_t pitch_elems, int H, int W, int border_type, int border_const, int h, int w,
                        const FirTaps& taps, float* out, uint8_t* icon, int64_t icon_pitch, cudaStream_t stream) {
    switch (taps.n) {
        case 2: return launch_tile_l<TIn, C, 2>(src, pitch_elems, H, W, border_type, border_const, h, w, taps, out, icon, icon_pitch, stream);
        case 4: return launch_tile_l<TIn, C, 4>(src, pitch_elems, H, W, border_type, border_const, h, w, taps, out, icon, icon_pitch, stream);
        case 6: return launch_tile_l<TIn, C, 6>(src, pitch_elems, H, W, border_type, border_const, h, w, taps, out, icon, icon_pitch, stream);
        default: return launch_tile_l<TIn, C, 8>(src, pitch_elems, H, W, border_type, border_const, h, w, taps, out, icon, icon_pitch, stream);
    }
}

int grid_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

// d_rows: scratch of Hp * (Wp/2) * C floats; d_ll: scratch of (Hp/2) * (Wp/2) * C floats (both for level 1, the largest).
cudaError_t launch_wavelet_fir(const uint8_t* d_src, int64_t pitch, int H, int W, int C, int depth, int border_type,
                               int border_const, const FirTaps& taps, uint8_t* d_icon, int64_t icon_pitch, float* d_rows,
                               float* d_ll, cudaStream_t stream) {
    const int ratio = 1 << depth;
    int h = (H + ratio - 1) / ratio * ratio, w = (W + ratio - 1) / ratio * ratio;
    if ((C == 1 || C == 3 || C == 4) && tiled_taps(taps.n)) {
        // tiled path: one launch per level; LL planes ping-pong between the two scratch buffers (level 1 -> d_rows)
        const float* in = nullptr;
        for (int level = 1; level <= depth; ++level) {
            float* out = (level & 1) ? d_rows : d_ll;
            uint8_t* icon = level == depth ? d_icon : nullptr;
            cudaError_t e;
            if (level == 1) {
                e = C == 1 ? launch_tile<uint8_t, 1>(d_src, pitch, H, W, border_type, border_const, h, w, taps, out, icon, icon_pitch, stream)
                  : C == 3 ? launch_tile<uint8_t, 3>(d_src, pitch, H, W, border_type, border_const, h, w, taps, out, icon, icon_pitch, stream)
                           : launch_tile<uint8_t, 4>(d_src, pitch, H, W, border_type, border_const, h, w, taps, out, icon, icon_pitch, stream);
            } else {
                const int64_t pe = (int64_t)w * C;
                e = C == 1 ? launch_tile<float, 1>(in, pe, h, w, 0, 0, h, w, taps, out, icon, icon_pitch, stream)
                  : C == 3 ? launch_tile<float, 3>(in, pe, h, w, 0, 0, h, w, taps, out, icon, icon_pitch, stream)
                           : launch_tile<float, 4>(in, pe, h, w, 0, 0, h, w, taps, out, icon, icon_pitch, stream);
            }
            if (e != cudaSuccess) return e;
            in = out;
            h /= 2; w /= 2;
        }
        return cudaSuccess;
    }
    for (int level = 1; level <= depth; ++level) {
        const int64_t n_rows = (int64_t)h * (w / 2) * C;
        if (level == 1) fir_rows_kernel<true><<<grid_for(n_rows), 256, 0, stream>>>(d_src, pitch, H, W, border_type, border_const, nullptr, h, w, C, taps, d_rows);
        else fir_rows_kernel<false><<<grid_for(n_rows), 256, 0, stream>>>(nullptr, 0, 0, 0, 0, 0, d_ll, h, w, C, taps, d_rows);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const int64_t n_cols = (int64_t)(h / 2) * (w / 2) * C;
        const bool last = level == depth;
        fir_cols_kernel<<<grid_for(n_cols), 256, 0, stream>>>(d_rows, h, w / 2, C, taps, d_ll, last ? d_icon : nullptr, icon_pitch);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        h /= 2; w /= 2;
    }
    return cudaSuccess;
}

}  // namespace wicca
