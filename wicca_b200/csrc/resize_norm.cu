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
__global__ void resize_norm_kernel(ResizeTables t, int img0, int out_h, int out_w, int norm_mode, float* __restrict__ out,
                                   uint8_t* __restrict__ out_u8, int skip_area) {
    // one image per blockIdx.y; the element index inside an image fits 32 bits (the launcher checks), which keeps
    // the index arithmetic off the 64-bit division path
    const uint32_t per = (uint32_t)out_h * (uint32_t)out_w * 3u;
    const int img = img0 + blockIdx.y;
    const ResizeJob j = t.jobs[img];
    if (skip_area && j.regime == 2) return;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < per; e += gridDim.x * blockDim.x) {
        const uint32_t px = e / 3u;
        const int c = (int)(e - px * 3u);
        const int dy = (int)(px / (uint32_t)out_w);
        const int dx = (int)(px - (uint32_t)dy * (uint32_t)out_w);
        const int64_t i = (int64_t)img * per + e;
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
// General-area regime, row-streaming form: one CTA per (output row, image), one thread per output
// pixel.  The few source rows that feed the output row are streamed through a double-buffered
// shared-memory row (one 1-D bulk copy per row, completion on an mbarrier; every source byte is
// read from HBM/L2 exactly once).  A thread's x taps are a contiguous run of pixels of the row:
// [left partial] + n_full whole pixels + [right partial].  The run is read as 32-bit words,
// realigned with one funnel shift per word and consumed four taps (12 bytes) at a time, per channel
// in OpenCV's order (three independent chains per thread).  The run always starts at pixel
// s_first - 1 and is padded to whole groups with weight-0 taps: src * 0 = +0 and h + 0 = h exactly,
// so a padded tap changes nothing, and only the first and the last group need per-tap weights
// (kept in registers: they do not depend on the row).
//
// src * alpha costs two instructions per byte: PRMT builds the float 2^23 + src, and
// fma(2^23 + src, alpha, -2^23 * alpha) rounds the exact product src * alpha once, i.e. it is
// bit-identical to (float)src * alpha (2^23 * alpha is exact).
// ------------------------------------------------------------------------------------------
constexpr int kRowsMaxThreads = 512;      // one thread per output pixel: out_w <= 512
constexpr int kRowPadFront = 16;          // bytes in front of a staged row (pixel -1 of a weight-0 left tap)
constexpr int kRowPadBack = 32;           // bytes behind it (weight-0 taps of the last group + word overread)

__device__ __forceinline__ uint32_t rs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rs_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rs_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rs_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rs_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rs_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(rs_smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void rs_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     rs_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(rs_smem_u32(bar))
                 : "memory");
}

// (float)byte_k(word) * w, exactly: see the note above.  nw = -2^23 * w.
template <int K>
__device__ __forceinline__ float byte_times(uint32_t word, float w, float nw) {
    return __fmaf_rn(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7440 | K)), w, nw);
}

// Four taps = 12 bytes = three realigned words; tap t has weight w[t].  Channel chains stay in tap order.
__device__ __forceinline__ void area_group(const uint32_t*& wp, uint32_t& w0, uint32_t shift, const float (&w)[4],
                                           const float (&nw)[4], float (&h)[3]) {
    const uint32_t w1 = wp[1], w2 = wp[2], w3 = wp[3];
    const uint32_t a0 = __funnelshift_r(w0, w1, shift), a1 = __funnelshift_r(w1, w2, shift), a2 = __funnelshift_r(w2, w3, shift);
    h[0] = __fadd_rn(h[0], byte_times<0>(a0, w[0], nw[0])); h[1] = __fadd_rn(h[1], byte_times<1>(a0, w[0], nw[0]));
    h[2] = __fadd_rn(h[2], byte_times<2>(a0, w[0], nw[0])); h[0] = __fadd_rn(h[0], byte_times<3>(a0, w[1], nw[1]));
    h[1] = __fadd_rn(h[1], byte_times<0>(a1, w[1], nw[1])); h[2] = __fadd_rn(h[2], byte_times<1>(a1, w[1], nw[1]));
    h[0] = __fadd_rn(h[0], byte_times<2>(a1, w[2], nw[2])); h[1] = __fadd_rn(h[1], byte_times<3>(a1, w[2], nw[2]));
    h[2] = __fadd_rn(h[2], byte_times<0>(a2, w[2], nw[2])); h[0] = __fadd_rn(h[0], byte_times<1>(a2, w[3], nw[3]));
    h[1] = __fadd_rn(h[1], byte_times<2>(a2, w[3], nw[3])); h[2] = __fadd_rn(h[2], byte_times<3>(a2, w[3], nw[3]));
    w0 = w3;
    wp += 3;
}

__global__ void __launch_bounds__(kRowsMaxThreads)
resize_area_rows_kernel(ResizeTables t, int out_h, int out_w, int norm_mode, float* __restrict__ out,
                        uint8_t* __restrict__ out_u8, int buf_bytes) {
    extern __shared__ __align__(16) uint8_t s_rows[];          // two row buffers of buf_bytes each
    __shared__ __align__(8) uint64_t s_full[2];
    const int img = blockIdx.y, dy = blockIdx.x;
    const ResizeJob j = t.jobs[img];
    if (j.regime != 2) return;
    const AreaDesc ay = t.area[j.yoff + dy];
    // this CTA's segment of the output row (blockIdx.z) and the source bytes that feed it
    const int x0 = blockIdx.z * blockDim.x;
    if (x0 >= out_w) return;
    const int x_last = min(x0 + (int)blockDim.x, out_w) - 1;
    const AreaDesc a_first = t.area[j.xoff + x0], a_last = t.area[j.xoff + x_last];
    const int row_bytes = j.sw * 3;
    const int lo = max(0, (a_first.s_first - 1) * 3) & ~15;                            // first byte fetched
    const int hi = min(row_bytes, (a_last.s_first + a_last.n_full + 1) * 3);           // one past the last byte needed
    const int seg_bytes = hi - lo;
    const int copy_bytes = (seg_bytes + 15) & ~15;
    // a segment can be bulk-copied when it starts on a 16-byte boundary and its rounded-up length stays inside the pitch
    const bool bulk = (((uintptr_t)j.src | (uintptr_t)j.pitch) & 15) == 0 && lo + copy_bytes <= j.pitch;
    if (threadIdx.x == 0) {
        rs_mbar_init(&s_full[0], 1); rs_mbar_init(&s_full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the y taps of this output row, in OpenCV's order: [left partial] + full rows + [right partial]
    const int has_l = ay.w_left != 0.0f, has_r = ay.w_right != 0.0f;
    const int n_rows = has_l + ay.n_full + has_r;
    auto row_index = [&](int k) { return (has_l && k == 0) ? ay.s_left : (k - has_l < ay.n_full ? ay.s_first + (k - has_l) : ay.s_right); };
    auto row_weight = [&](int k) { return (has_l && k == 0) ? ay.w_left : (k - has_l < ay.n_full ? ay.w_full : ay.w_right); };
    auto fetch = [&](int k) {
        const uint8_t* g = j.src + (int64_t)row_index(k) * j.pitch + lo;
        uint8_t* s = s_rows + (k & 1) * buf_bytes + kRowPadFront;
        if (bulk) {
            if (threadIdx.x == 0) {
                rs_mbar_expect_tx(&s_full[k & 1], (uint32_t)copy_bytes);
                rs_bulk_load(s, g, (uint32_t)copy_bytes, &s_full[k & 1]);
            }
        } else {
            for (int b = threadIdx.x; b < seg_bytes; b += blockDim.x) s[b] = g[b];
        }
    };
    // ---- this thread's x taps (independent of the row)
    const int dx = x0 + threadIdx.x;
    const bool active = dx < out_w;
    const AreaDesc ax = t.area[j.xoff + (active ? dx : 0)];
    const int n_taps = ax.n_full + 2;                          // pixels s_first - 1 .. s_first + n_full
    const int groups = (n_taps + 3) >> 2;
    auto tap_weight = [&](int tp) { return tp == 0 ? ax.w_left : (tp <= ax.n_full ? ax.w_full : (tp == ax.n_full + 1 ? ax.w_right : 0.0f)); };
    float wa[4], nwa[4], wz[4], nwz[4], wm[4], nwm[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        wa[q] = tap_weight(q); nwa[q] = -8388608.0f * wa[q];
        wz[q] = tap_weight(4 * (groups - 1) + q); nwz[q] = -8388608.0f * wz[q];
        wm[q] = ax.w_full; nwm[q] = -8388608.0f * ax.w_full;
    }
    const int b_start = (ax.s_first - 1) * 3 - lo + kRowPadFront;   // >= 13
    const uint32_t shift = (uint32_t)(b_start & 3) * 8;
    float acc[3] = {0.0f, 0.0f, 0.0f};
    __syncthreads();                                           // mbarrier init visible
    fetch(0);
    for (int k = 0; k < n_rows; ++k) {
        if (k + 1 < n_rows) fetch(k + 1);
        if (bulk) rs_mbar_wait(&s_full[k & 1], (uint32_t)(k >> 1) & 1u);
        else __syncthreads();
        if (active) {
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(s_rows + (k & 1) * buf_bytes) + (b_start >> 2);
            uint32_t w0 = wp[0];
            float h[3] = {0.0f, 0.0f, 0.0f};
            area_group(wp, w0, shift, wa, nwa, h);
            for (int g = 1; g < groups - 1; ++g) area_group(wp, w0, shift, wm, nwm, h);
            if (groups > 1) area_group(wp, w0, shift, wz, nwz, h);
            const float beta = row_weight(k);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float bh = __fmul_rn(beta, h[c]);
                acc[c] = (k == 0) ? bh : __fadd_rn(acc[c], bh);
            }
        }
        __syncthreads();          // the buffer is refilled two iterations later
    }
    if (active) {
        const int64_t i = (((int64_t)img * out_h + dy) * out_w + dx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) emit_value(sat_rint_u8(acc[c]), c, i + c, norm_mode, out, out_u8);
    }
}

// max_src_w: widest source of the batch; n_area / n_other: how many images are / are not in the
// general-area regime (the host knows, it built the tables).
cudaError_t launch_resize_norm(const ResizeTables& t, int n, int out_h, int out_w, int norm_mode, float* d_out,
                               uint8_t* d_out_u8, int max_src_w, int n_area, int n_other, cudaStream_t stream) {
    const int64_t total = (int64_t)n * out_h * out_w * 3;
    if (total <= 0) return cudaSuccess;
    // one CTA per (output row, image, row segment); a segment is as wide as possible while the grid still fills
    // the GPU a few times over (a lone 53 MP image would otherwise run on 224 CTAs, each streaming 0.7 MB)
    int seg_w = (out_w + 31) / 32 * 32;
    while (seg_w > 32 && (int64_t)out_h * (n_area > 0 ? n_area : 1) * ((out_w + seg_w - 1) / seg_w) < 148 * 8) seg_w = (seg_w / 2 + 31) / 32 * 32;
    const int segs = (out_w + seg_w - 1) / seg_w;
    // source span of one segment: seg_w target pixels x scale, + the early pixel, the padded last group, alignment slack
    const int64_t span_px = segs == 1 ? (int64_t)max_src_w + 8 : ((int64_t)seg_w * max_src_w + out_w - 1) / out_w + 10;
    const int buf_bytes = kRowPadFront + (int)((span_px * 3 + 16 + 15) / 16 * 16) + kRowPadBack;
    const size_t smem = (size_t)2 * buf_bytes;
    const bool rows_ok = n_area > 0 && out_w <= kRowsMaxThreads && smem <= 200 * 1024 && n <= 65535 && segs <= 65535;
    if (rows_ok) {
        static thread_local int configured_dev = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (configured_dev != dev) {
            cudaError_t e = cudaFuncSetAttribute(resize_area_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return e;
            configured_dev = dev;
        }
        resize_area_rows_kernel<<<dim3(out_h, n, segs), seg_w, smem, stream>>>(t, out_h, out_w, norm_mode, d_out, d_out_u8, buf_bytes);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (n_other == 0) return cudaSuccess;
    }
    const int64_t per = (int64_t)out_h * out_w * 3;
    if (per > 0x7FFFFFFF) return cudaErrorInvalidValue;                    // 32-bit element index inside an image
    for (int img0 = 0; img0 < n; img0 += 65535) {                          // one image per blockIdx.y
        const int m = n - img0 < 65535 ? n - img0 : 65535;
        int64_t blocks = (per + 255) / 256;
        const int64_t cap = (148 * 32 + m - 1) / m;                        // enough CTAs in all to fill the GPU a few times
        if (blocks > cap) blocks = cap;
        resize_norm_kernel<<<dim3((unsigned)blocks, (unsigned)m), 256, 0, stream>>>(t, img0, out_h, out_w, norm_mode, d_out, d_out_u8, rows_ok ? 1 : 0);
    }
    return cudaGetLastError();
}

}  // namespace wicca
