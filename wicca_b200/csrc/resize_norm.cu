// resize_norm.cu - classifier-input epilogue: icon (uint8 HWC) -> cv2.resize(..., INTER_AREA)
// (wicca/classifying_tools.py:318) -> np.stack (:323) -> preprocess_input + float32 cast
// (:286-287), one thread per output element, bit-exact with OpenCV's three 8-bit code paths
// (restated in oracle/resize_oracle.py):
//   regime 1  integer factors     : integer block sum; 2x2 -> (s+2)>>2, else rint(float(s) * (1/area))
//   regime 2  general area        : fp32 taps, accumulated in OpenCV's order (x taps inside each
//                                   source row, then beta * row over the y taps), rint, saturate
//   regime 3  any axis upscaled   : 2-tap bilinear in 11-bit fixed point ("area mode" taps)
// No FMA contraction anywhere: OpenCV's scalar code multiplies and adds separately.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace wicca {

__device__ __forceinline__ float norm_value(float v, int c, int mode) {
    switch (mode) {
        case 1: return __fsub_rn(__fdiv_rn(v, 127.5f), 1.0f);                                   // "tf"
        case 3: {                                                                               // "torch"
            const float mean = (c == 0) ? 0.485f : (c == 1) ? 0.456f : 0.406f;
            const float sd = (c == 0) ? 0.229f : (c == 1) ? 0.224f : 0.225f;
            return __fdiv_rn(__fsub_rn(__fdiv_rn(v, 255.0f), mean), sd);
        }
        default: return v;                                                                      // identity
    }
}

// uint8 -> float without the quarter-rate I2F: 0x4B000000 | v is the float 2^23 + v, exactly.
__device__ __forceinline__ float u8_to_float(uint32_t v) { return __fsub_rn(__uint_as_float(0x4B000000u | v), 8388608.0f); }

__device__ __forceinline__ uint8_t sat_rint_u8(float x) {
    const float r = rintf(x);                       // round half to even, like cvRound
    return (uint8_t)fminf(fmaxf(r, 0.0f), 255.0f);
}

__device__ __forceinline__ void emit_value(uint8_t v, int c, int64_t i, int norm_mode, float* __restrict__ out,
                                           uint8_t* __restrict__ out_u8) {
    if (out_u8) out_u8[i] = v;
    if (norm_mode == 2) {
        // "caffe": RGB -> BGR then subtract the BGR means: output channel (2 - c) takes this value
        const float mean = (c == 0) ? 123.68f : (c == 1) ? 116.779f : 103.939f;
        out[i - c + (2 - c)] = __fsub_rn((float)v, mean);
    } else {
        out[i] = norm_value((float)v, c, norm_mode);
    }
}

// skip_area != 0: images in the general-area regime are left to resize_area_rows_kernel.
__global__ void resize_norm_kernel(ResizeTables t, int n, int out_h, int out_w, int norm_mode, float* __restrict__ out,
                                   uint8_t* __restrict__ out_u8, int skip_area) {
    const int64_t per = (int64_t)out_h * out_w * 3;
    const int64_t total = per * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int img = (int)(i / per);
        int64_t r = i - (int64_t)img * per;
        const int c = (int)(r % 3);
        r /= 3;
        const int dx = (int)(r % out_w);
        const int dy = (int)(r / out_w);
        const ResizeJob j = t.jobs[img];
        if (skip_area && j.regime == 2) continue;
        const int64_t srow = j.pitch;
        const uint8_t* s = j.src + c;
        uint8_t v;
        if (j.regime == 0) {
            v = s[(int64_t)dy * srow + (int64_t)dx * 3];
        } else if (j.regime == 1) {
            uint32_t sum = 0;
            for (int yy = 0; yy < j.isy; ++yy) {
                const uint8_t* p = s + (int64_t)(dy * j.isy + yy) * srow + (int64_t)(dx * j.isx) * 3;
                for (int xx = 0; xx < j.isx; ++xx) sum += p[xx * 3];
            }
            if (j.isx == 2 && j.isy == 2) v = (uint8_t)((sum + 2) >> 2);
            else v = sat_rint_u8(__fmul_rn((float)sum, 1.0f / (float)(j.isx * j.isy)));
        } else if (j.regime == 2) {
            const AreaDesc ax = t.area[j.xoff + dx];
            const AreaDesc ay = t.area[j.yoff + dy];
            // one source row: buf = 0; buf += src * alpha over the x taps, in order (0 + x == x)
            auto hrow = [&](int sy) -> float {
                const uint8_t* p = s + (int64_t)sy * srow;
                float h = 0.0f;
                if (ax.w_left != 0.0f) h = __fmul_rn(u8_to_float(p[(int64_t)ax.s_left * 3]), ax.w_left);
                const uint8_t* q = p + (int64_t)ax.s_first * 3;
#pragma unroll 4
                for (int kx = 0; kx < ax.n_full; ++kx) h = __fadd_rn(h, __fmul_rn(u8_to_float(q[kx * 3]), ax.w_full));
                if (ax.w_right != 0.0f) h = __fadd_rn(h, __fmul_rn(u8_to_float(p[(int64_t)ax.s_right * 3]), ax.w_right));
                return h;
            };
            // sum = beta * buf for the first y tap, sum += beta * buf for the others
            float acc = 0.0f;
            if (ay.w_left != 0.0f) acc = __fmul_rn(ay.w_left, hrow(ay.s_left));
            for (int ky = 0; ky < ay.n_full; ++ky) acc = __fadd_rn(acc, __fmul_rn(ay.w_full, hrow(ay.s_first + ky)));
            if (ay.w_right != 0.0f) acc = __fadd_rn(acc, __fmul_rn(ay.w_right, hrow(ay.s_right)));
            v = sat_rint_u8(acc);
        } else {
            const LinTap tx = t.lin[j.xoff + dx];
            const LinTap ty = t.lin[j.yoff + dy];
            const uint8_t* p0 = s + (int64_t)ty.i0 * srow;
            const uint8_t* p1 = s + (int64_t)ty.i1 * srow;
            const int r0 = ((int)p0[(int64_t)tx.i0 * 3] * tx.c0 + (int)p0[(int64_t)tx.i1 * 3] * tx.c1) >> 4;
            const int r1 = ((int)p1[(int64_t)tx.i0 * 3] * tx.c0 + (int)p1[(int64_t)tx.i1 * 3] * tx.c1) >> 4;
            int o = (((ty.c0 * r0) >> 16) + ((ty.c1 * r1) >> 16) + 2) >> 2;
            o = o < 0 ? 0 : (o > 255 ? 255 : o);
            v = (uint8_t)o;
        }
        emit_value(v, c, i, norm_mode, out, out_u8);
    }
}

// ------------------------------------------------------------------------------------------
// General-area regime, row-streaming form: one CTA per (output row, image).  The few source
// rows that feed the output row are streamed through a double-buffered shared-memory row with
// 16-byte cp.async copies (every source byte is read from HBM/L2 exactly once, coalesced); each
// thread owns up to kRowsMaxElems (dx, c) elements of the output row, keeps their closed-form x
// taps in registers and accumulates the rows in OpenCV's order.
// ------------------------------------------------------------------------------------------
constexpr int kRowsThreads = 256;
constexpr int kRowsMaxElems = 4;          // covers out_w * 3 <= 1024

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(kRowsThreads)
resize_area_rows_kernel(ResizeTables t, int out_h, int out_w, int norm_mode, float* __restrict__ out,
                        uint8_t* __restrict__ out_u8, int buf_bytes) {
    extern __shared__ __align__(16) uint8_t s_rows[];          // two row buffers of buf_bytes each
    const int img = blockIdx.y, dy = blockIdx.x;
    const ResizeJob j = t.jobs[img];
    if (j.regime != 2) return;
    const AreaDesc ay = t.area[j.yoff + dy];
    const int row_bytes = j.sw * 3;
    const int n_elems = out_w * 3;
    // the y taps of this output row, in OpenCV's order: [left partial] + full rows + [right partial]
    const int has_l = ay.w_left != 0.0f, has_r = ay.w_right != 0.0f;
    const int n_rows = has_l + ay.n_full + has_r;
    auto row_index = [&](int k) { return (has_l && k == 0) ? ay.s_left : (k - has_l < ay.n_full ? ay.s_first + (k - has_l) : ay.s_right); };
    auto row_weight = [&](int k) { return (has_l && k == 0) ? ay.w_left : (k - has_l < ay.n_full ? ay.w_full : ay.w_right); };
    auto fetch = [&](int k) {
        const uint8_t* g = j.src + (int64_t)row_index(k) * j.pitch;
        uint8_t* s = s_rows + (k & 1) * buf_bytes;
        if (((uintptr_t)g & 15) == 0) {
            const int chunks = row_bytes >> 4;
            for (int q = threadIdx.x; q < chunks; q += kRowsThreads) cp_async16(s + q * 16, g + q * 16);
            for (int b = (chunks << 4) + threadIdx.x; b < row_bytes; b += kRowsThreads) s[b] = g[b];
        } else {
            for (int b = threadIdx.x; b < row_bytes; b += kRowsThreads) s[b] = g[b];
        }
        cp_async_commit();
    };
    AreaDesc ax[kRowsMaxElems];
    int ch[kRowsMaxElems];
    float acc[kRowsMaxElems];
#pragma unroll
    for (int m = 0; m < kRowsMaxElems; ++m) {
        const int e = threadIdx.x + m * kRowsThreads;
        const int dx = e < n_elems ? e / 3 : 0;
        ch[m] = e - (e / 3) * 3;
        ax[m] = t.area[j.xoff + dx];
        acc[m] = 0.0f;
    }
    if (n_rows > 0) fetch(0);
    for (int k = 0; k < n_rows; ++k) {
        if (k + 1 < n_rows) { fetch(k + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const uint8_t* s = s_rows + (k & 1) * buf_bytes;
        const float beta = row_weight(k);
#pragma unroll
        for (int m = 0; m < kRowsMaxElems; ++m) {
            if (threadIdx.x + m * kRowsThreads < n_elems) {
                const uint8_t* p = s + ch[m];
                float h = 0.0f;
                if (ax[m].w_left != 0.0f) h = __fmul_rn(u8_to_float(p[ax[m].s_left * 3]), ax[m].w_left);
                const uint8_t* q = p + ax[m].s_first * 3;
                const float wf = ax[m].w_full;
#pragma unroll 4
                for (int kx = 0; kx < ax[m].n_full; ++kx) h = __fadd_rn(h, __fmul_rn(u8_to_float(q[kx * 3]), wf));
                if (ax[m].w_right != 0.0f) h = __fadd_rn(h, __fmul_rn(u8_to_float(p[ax[m].s_right * 3]), ax[m].w_right));
                const float bh = __fmul_rn(beta, h);
                acc[m] = (k == 0) ? bh : __fadd_rn(acc[m], bh);
            }
        }
        __syncthreads();          // the buffer is refilled two iterations later
    }
#pragma unroll
    for (int m = 0; m < kRowsMaxElems; ++m) {
        const int e = threadIdx.x + m * kRowsThreads;
        if (e < n_elems) {
            const int64_t i = ((int64_t)img * out_h + dy) * n_elems + e;
            emit_value(sat_rint_u8(acc[m]), ch[m], i, norm_mode, out, out_u8);
        }
    }
}

// max_src_w: widest source of the batch; n_area / n_other: how many images are / are not in the
// general-area regime (the host knows, it built the tables).
cudaError_t launch_resize_norm(const ResizeTables& t, int n, int out_h, int out_w, int norm_mode, float* d_out,
                               uint8_t* d_out_u8, int max_src_w, int n_area, int n_other, cudaStream_t stream) {
    const int64_t total = (int64_t)n * out_h * out_w * 3;
    if (total <= 0) return cudaSuccess;
    const int buf_bytes = (max_src_w * 3 + 31) / 16 * 16;
    const size_t smem = (size_t)2 * buf_bytes;
    const bool rows_ok = n_area > 0 && out_w * 3 <= kRowsThreads * kRowsMaxElems && smem <= 200 * 1024 && n <= 65535;
    if (rows_ok) {
        static thread_local int configured_dev = -1;
        static thread_local size_t configured_smem = 0;
        int dev = 0;
        cudaGetDevice(&dev);
        if (configured_dev != dev || configured_smem < smem) {
            cudaError_t e = cudaFuncSetAttribute(resize_area_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return e;
            configured_dev = dev; configured_smem = 200 * 1024;
        }
        resize_area_rows_kernel<<<dim3(out_h, n), kRowsThreads, smem, stream>>>(t, out_h, out_w, norm_mode, d_out, d_out_u8, buf_bytes);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (n_other == 0) return cudaSuccess;
    }
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    resize_norm_kernel<<<(int)blocks, 256, 0, stream>>>(t, n, out_h, out_w, norm_mode, d_out, d_out_u8, rows_ok ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace wicca
