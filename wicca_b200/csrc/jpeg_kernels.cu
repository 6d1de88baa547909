// jpeg_kernels.cu - the data-parallel half of JPEG decoding (row N2), bit-exact with libjpeg-turbo's defaults
// (what cv2.imread runs; restated and pinned in oracle/jpeg_oracle.py):
//   jpeg_idct_kernel   dequantise + jpeg_idct_islow (jidctint.c: CONST_BITS 13, PASS1_BITS 2), one thread per
//                      8 x 8 block, both passes in registers, range-limit to uint8 component planes
//   jpeg_color_kernel  fancy (triangle) / box chroma upsampling (jdsample.c) evaluated per output pixel from the
//                      component planes, YCbCr -> RGB with the 16-bit fixed-point constants of jdcolor.c, written
//                      as a pitched RGB image - the layout every other kernel of this library reads.
// Bound: HBM (2 B/coefficient in, 1 B/sample out; then <= 3 B/px in, 3 B/px out); both are far cheaper than the
// Huffman decoding that feeds them, which is serial per scan and stays on the CPU.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace wicca {

namespace {

constexpr int F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270,
              F_0_899976223 = 7373, F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137,
              F_1_961570560 = 16069, F_2_053119869 = 16819, F_2_562915447 = 20995, F_3_072711026 = 25172;

// One 8-point pass of the LL&M inverse DCT; out = (x + 2^(kShift-1)) >> kShift.
template <int kShift>
__device__ __forceinline__ void idct8(const int (&v)[8], int (&o)[8]) {
    int z2 = v[2], z3 = v[6];
    int z1 = (z2 + z3) * F_0_541196100;
    const int tmp2 = z1 - z3 * F_1_847759065;
    const int tmp3 = z1 + z2 * F_0_765366865;
    const int tmp0 = (v[0] + v[4]) << 13, tmp1 = (v[0] - v[4]) << 13;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    int t0 = v[7], t1 = v[5], t2 = v[3], t3 = v[1];
    z1 = t0 + t3; z2 = t1 + t2; z3 = t0 + t2;
    int z4 = t1 + t3;
    const int z5 = (z3 + z4) * F_1_175875602;
    t0 *= F_0_298631336; t1 *= F_2_053119869; t2 *= F_3_072711026; t3 *= F_1_501321110;
    z1 *= -F_0_899976223; z2 *= -F_2_562915447;
    z3 = z3 * -F_1_961570560 + z5;
    z4 = z4 * -F_0_390180644 + z5;
    t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
    constexpr int half = 1 << (kShift - 1);
    o[0] = (tmp10 + t3 + half) >> kShift; o[7] = (tmp10 - t3 + half) >> kShift;
    o[1] = (tmp11 + t2 + half) >> kShift; o[6] = (tmp11 - t2 + half) >> kShift;
    o[2] = (tmp12 + t1 + half) >> kShift; o[5] = (tmp12 - t1 + half) >> kShift;
    o[3] = (tmp13 + t0 + half) >> kShift; o[4] = (tmp13 - t0 + half) >> kShift;
}

// sample_range_limit + CENTERJSAMPLE indexed with (x & RANGE_MASK): clamp(x + 128) for x in [-512, 511] and the
// table's wrap-around beyond (jdmaster.c prepare_range_limit_table).
__device__ __forceinline__ uint32_t range_limit(int x) {
    const int idx = x & 1023;
    return idx < 128 ? idx + 128 : (idx < 512 ? 255 : (idx < 896 ? 0 : idx - 896));
}

__global__ void __launch_bounds__(128, 4)
jpeg_idct_kernel(JpegImageDesc d) {
    __shared__ int s_qt[3][64];
    for (int i = threadIdx.x; i < 64 * d.ncomp; i += blockDim.x) s_qt[i >> 6][i & 63] = d.comp[i >> 6].qt[i & 63];
    __syncthreads();
    const int c = blockIdx.y;
    const JpegPlaneDesc& p = d.comp[c];
    const int64_t n_blocks = (int64_t)p.blocks_w * p.blocks_h;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (int64_t)gridDim.x * blockDim.x) {
        const uint4* src = reinterpret_cast<const uint4*>(p.coefs + b * 64);
        int ws[8][8];                                       // [row][col]
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint4 q = src[r];                         // one row of 8 coefficients
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int coef = (int)(int16_t)(w[k >> 1] >> (16 * (k & 1)));
                ws[r][k] = coef * s_qt[c][r * 8 + k];
            }
        }
        // pass 1: columns
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int v[8], o[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = ws[r][k];
            idct8<13 - 2>(v, o);
#pragma unroll
            for (int r = 0; r < 8; ++r) ws[r][k] = o[r];
        }
        // pass 2: rows, + range limit; 8 bytes per row
        const int by = (int)(b / p.blocks_w), bx = (int)(b - (int64_t)by * p.blocks_w);
        uint8_t* out = p.plane + (int64_t)(by * 8) * p.plane_pitch + bx * 8;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int o[8];
            idct8<13 + 2 + 3>(ws[r], o);
            uint2 px;
            px.x = range_limit(o[0]) | (range_limit(o[1]) << 8) | (range_limit(o[2]) << 16) | (range_limit(o[3]) << 24);
            px.y = range_limit(o[4]) | (range_limit(o[5]) << 8) | (range_limit(o[6]) << 16) | (range_limit(o[7]) << 24);
            *reinterpret_cast<uint2*>(out + (int64_t)r * p.plane_pitch) = px;
        }
    }
}

// The component's sample at image position (x, y) after libjpeg's upsampling.
__device__ __forceinline__ int upsampled(const JpegPlaneDesc& p, int x, int y) {
    const uint8_t* P = p.plane;
    const int pitch = p.plane_pitch;
    switch (p.mode) {
        case 0:
            return P[(int64_t)y * pitch + x];
        case 1: {                                            // h2v1 fancy: 3/4 nearer + 1/4 farther, biases 1 / 2
            const uint8_t* row = P + (int64_t)y * pitch;
            const int cx = x >> 1, cur = row[cx];
            if (x & 1) return cx == p.dw - 1 ? cur : (3 * cur + row[cx + 1] + 2) >> 2;
            return cx == 0 ? cur : (3 * cur + row[cx - 1] + 1) >> 2;
        }
        case 2: {                                            // h2v2 fancy: rows 3:1, then columns 3:1 on the sums
            const int cy = y >> 1, cx = x >> 1;
            const int fy = (y & 1) ? min(cy + 1, p.dh - 1) : max(cy - 1, 0);
            const uint8_t* r0 = P + (int64_t)cy * pitch;     // nearer row
            const uint8_t* r1 = P + (int64_t)fy * pitch;     // farther row (edge rows replicate)
            const int cur = 3 * r0[cx] + r1[cx];
            if (x & 1) return cx == p.dw - 1 ? (cur * 4 + 7) >> 4 : (3 * cur + 3 * r0[cx + 1] + r1[cx + 1] + 7) >> 4;
            return cx == 0 ? (cur * 4 + 8) >> 4 : (3 * cur + 3 * r0[cx - 1] + r1[cx - 1] + 8) >> 4;
        }
        case 3: {                                            // h1v2 fancy
            const int cy = y >> 1;
            const int fy = (y & 1) ? min(cy + 1, p.dh - 1) : max(cy - 1, 0);
            return (3 * P[(int64_t)cy * pitch + x] + P[(int64_t)fy * pitch + x] + ((y & 1) ? 2 : 1)) >> 2;
        }
        default:                                             // box replication
            return P[(int64_t)(y / p.vf) * pitch + x / p.hf];
    }
}

constexpr int FIX_1_40200 = 91881, FIX_1_77200 = 116130, FIX_0_71414 = 46802, FIX_0_34414 = 22554;   // x * 65536 + 0.5

__device__ __forceinline__ int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// Eight upsampled samples of one component for the image positions x0 .. x0+7 (x0 a multiple of 8) of row y.
// The planes are whole blocks wide (a multiple of 8 samples), so aligned word loads never leave a row; columns
// beyond the real samples only feed pixels beyond the image, which are not stored.
__device__ __forceinline__ void upsampled8(const JpegPlaneDesc& p, int x0, int y, int (&out)[8]) {
    const uint8_t* P = p.plane;
    const int pitch = p.plane_pitch;
    auto load4 = [&](const uint8_t* row, int c, int (&v)[4]) {       // row[c .. c+3], c a multiple of 4; zeros outside
        uint32_t w = 0;
        if (c >= 0 && c < pitch) w = *reinterpret_cast<const uint32_t*>(row + c);
        v[0] = w & 255; v[1] = (w >> 8) & 255; v[2] = (w >> 16) & 255; v[3] = w >> 24;
    };
    if (p.mode == 0 || p.mode == 3) {
        const uint8_t* r0 = P + (int64_t)(p.mode == 0 ? y : (y >> 1)) * pitch + x0;
        int a[4], b[4];
        load4(r0, 0, a); load4(r0, 4, b);
        int near8[8] = {a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3]};
        if (p.mode == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) out[i] = near8[i];
            return;
        }
        const int cy = y >> 1;
        const int fy = (y & 1) ? min(cy + 1, p.dh - 1) : max(cy - 1, 0);
        const uint8_t* r1 = P + (int64_t)fy * pitch + x0;
        load4(r1, 0, a); load4(r1, 4, b);
        const int far8[8] = {a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3]};
        const int bias = (y & 1) ? 2 : 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = (3 * near8[i] + far8[i] + bias) >> 2;
        return;
    }
    if (p.mode == 1 || p.mode == 2) {
        const int c0 = x0 >> 1;                                      // first chroma column, a multiple of 4
        int v[6];                                                    // columns c0-1 .. c0+4 (row sums for h2v2)
        int a[4], b[4], c[4];
        if (p.mode == 1) {
            const uint8_t* row = P + (int64_t)y * pitch;
            load4(row, c0 - 4, a); load4(row, c0, b); load4(row, c0 + 4, c);
            v[0] = a[3]; v[1] = b[0]; v[2] = b[1]; v[3] = b[2]; v[4] = b[3]; v[5] = c[0];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int cx = c0 + i;
                out[2 * i] = cx == 0 ? v[1 + i] : (3 * v[1 + i] + v[i] + 1) >> 2;
                out[2 * i + 1] = cx == p.dw - 1 ? v[1 + i] : (3 * v[1 + i] + v[2 + i] + 2) >> 2;
            }
            return;
        }
        const int cy = y >> 1;
        const int fy = (y & 1) ? min(cy + 1, p.dh - 1) : max(cy - 1, 0);
        const uint8_t* r0 = P + (int64_t)cy * pitch;
        const uint8_t* r1 = P + (int64_t)fy * pitch;
        int d[4], e[4], f[4];
        load4(r0, c0 - 4, a); load4(r0, c0, b); load4(r0, c0 + 4, c);
        load4(r1, c0 - 4, d); load4(r1, c0, e); load4(r1, c0 + 4, f);
        v[0] = 3 * a[3] + d[3];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[1 + i] = 3 * b[i] + e[i];
        v[5] = 3 * c[0] + f[0];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int cx = c0 + i;
            out[2 * i] = cx == 0 ? (v[1 + i] * 4 + 8) >> 4 : (3 * v[1 + i] + v[i] + 8) >> 4;
            out[2 * i + 1] = cx == p.dw - 1 ? (v[1 + i] * 4 + 7) >> 4 : (3 * v[1 + i] + v[2 + i] + 7) >> 4;
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = upsampled(p, x0 + i, y);      // box replication
}

// One thread = eight horizontally adjacent pixels = 24 bytes of the RGB row.
__global__ void __launch_bounds__(256)
jpeg_color_kernel(JpegImageDesc d) {
    const int groups = (d.width + 7) >> 3;
    const int64_t total = (int64_t)groups * d.height;
    const bool wide = (((uintptr_t)d.dst | (uintptr_t)d.dst_pitch) & 7) == 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / groups), x0 = (int)(i - (int64_t)y * groups) * 8;
        int yy[8], cb[8], cr[8];
        upsampled8(d.comp[0], x0, y, yy);
        uint32_t rgb[6] = {0, 0, 0, 0, 0, 0};
        if (d.ncomp == 3) {
            upsampled8(d.comp[1], x0, y, cb);
            upsampled8(d.comp[2], x0, y, cr);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int r, g, b;
            if (d.ncomp == 1) {
                r = g = b = yy[k];
            } else {
                const int u = cb[k] - 128, v = cr[k] - 128;
                r = clamp255(yy[k] + ((FIX_1_40200 * v + 32768) >> 16));
                g = clamp255(yy[k] + ((-FIX_0_34414 * u + 32768 - FIX_0_71414 * v) >> 16));
                b = clamp255(yy[k] + ((FIX_1_77200 * u + 32768) >> 16));
            }
            const int o = 3 * k;
            rgb[o >> 2] |= (uint32_t)r << (8 * (o & 3));
            rgb[(o + 1) >> 2] |= (uint32_t)g << (8 * ((o + 1) & 3));
            rgb[(o + 2) >> 2] |= (uint32_t)b << (8 * ((o + 2) & 3));
        }
        uint8_t* row = d.dst + (int64_t)y * d.dst_pitch + (int64_t)x0 * 3;
        if (x0 + 8 <= d.width && wide) {
            uint2* w = reinterpret_cast<uint2*>(row);
            w[0] = make_uint2(rgb[0], rgb[1]); w[1] = make_uint2(rgb[2], rgb[3]); w[2] = make_uint2(rgb[4], rgb[5]);
        } else {
            const int n = min(8, d.width - x0) * 3;
            for (int k = 0; k < n; ++k) row[k] = (uint8_t)(rgb[k >> 2] >> (8 * (k & 3)));
        }
    }
}

// OpenCV's ApplyExifOrientation: 2 flip x, 3 flip both, 4 flip y, 5 transpose, 6 transpose + flip x, 7 transpose +
// flip both, 8 transpose + flip y.  One thread = four output pixels of a row (three words).
__global__ void __launch_bounds__(256)
jpeg_orient_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, int H, int W, int orientation, uint8_t* __restrict__ dst,
                   int64_t dst_pitch) {
    const int Ho = orientation >= 5 ? W : H, Wo = orientation >= 5 ? H : W;
    const int quads = (Wo + 3) >> 2;
    const int64_t total = (int64_t)quads * Ho;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int oy = (int)(i / quads), ox0 = (int)(i - (int64_t)oy * quads) * 4;
        uint8_t px[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ox = min(ox0 + k, Wo - 1);
            int sy, sx;
            switch (orientation) {
                case 2: sy = oy; sx = W - 1 - ox; break;
                case 3: sy = H - 1 - oy; sx = W - 1 - ox; break;
                case 4: sy = H - 1 - oy; sx = ox; break;
                case 5: sy = ox; sx = oy; break;
                case 6: sy = H - 1 - ox; sx = oy; break;
                case 7: sy = H - 1 - ox; sx = W - 1 - oy; break;
                default: sy = ox; sx = W - 1 - oy; break;       // 8
            }
            const uint8_t* p = src + (int64_t)sy * src_pitch + (int64_t)sx * 3;
            px[3 * k] = p[0]; px[3 * k + 1] = p[1]; px[3 * k + 2] = p[2];
        }
        uint8_t* row = dst + (int64_t)oy * dst_pitch + (int64_t)ox0 * 3;
        if (ox0 + 4 <= Wo && (((uintptr_t)dst | (uintptr_t)dst_pitch) & 3) == 0) {
            uint32_t* w = reinterpret_cast<uint32_t*>(row);
#pragma unroll
            for (int k = 0; k < 3; ++k)
                w[k] = (uint32_t)px[4 * k] | ((uint32_t)px[4 * k + 1] << 8) | ((uint32_t)px[4 * k + 2] << 16) | ((uint32_t)px[4 * k + 3] << 24);
        } else {
            const int n = min(4, Wo - ox0) * 3;
            for (int k = 0; k < n; ++k) row[k] = px[k];
        }
    }
}

}  // namespace

cudaError_t launch_jpeg_orient(const uint8_t* d_src, int64_t src_pitch, int H, int W, int orientation, uint8_t* d_dst,
                               int64_t dst_pitch, cudaStream_t stream) {
    const int Ho = orientation >= 5 ? W : H, Wo = orientation >= 5 ? H : W;
    int64_t g = ((int64_t)((Wo + 3) >> 2) * Ho + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    jpeg_orient_kernel<<<(unsigned)g, 256, 0, stream>>>(d_src, src_pitch, H, W, orientation, d_dst, dst_pitch);
    return cudaGetLastError();
}

cudaError_t launch_jpeg_decode(const JpegImageDesc& d, cudaStream_t stream) {
    int64_t max_blocks = 0;
    for (int c = 0; c < d.ncomp; ++c) max_blocks = max_blocks > (int64_t)d.comp[c].blocks_w * d.comp[c].blocks_h ? max_blocks : (int64_t)d.comp[c].blocks_w * d.comp[c].blocks_h;
    int64_t gx = (max_blocks + 127) / 128;
    if (gx > 148 * 16) gx = 148 * 16;
    if (gx < 1) gx = 1;
    jpeg_idct_kernel<<<dim3((unsigned)gx, d.ncomp), 128, 0, stream>>>(d);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int64_t total = (int64_t)((d.width + 7) >> 3) * d.height;
    int64_t g2 = (total + 255) / 256;
    if (g2 > 148 * 16) g2 = 148 * 16;
    if (g2 < 1) g2 = 1;
    jpeg_color_kernel<<<(unsigned)g2, 256, 0, stream>>>(d);
    return cudaGetLastError();
}

}  // namespace wicca
