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
