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
    for (int l = 1; l <= depth; ++l) {
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
    for (int l = depth; l >= 1; --l) {
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
    return cudaSuccess;
}

}  // namespace wicca
