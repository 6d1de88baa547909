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
// s_first - 1 and ends in a group of four or of two taps, padded with weight-0 taps where needed:
// src * 0 = +0 and h + 0 = h exactly, so a padded tap changes nothing, and only the first and the
// last group need per-tap weights (kept in registers: they do not depend on the row).
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
__device__ __forceinline__ void rs_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rs_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void rs_bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// Blackwell's packed float32 pair instructions (FFMA2 / FADD2): two independent IEEE operations per issue slot.  The
// kernel is issue bound, and the accumulation chains of two SOURCE ROWS of the same output pixel are independent
// (OpenCV finishes a row's x taps before it touches the sum over rows), so a thread walks two rows at once with the
// chains of row A in the low halves and those of row B in the high halves: 1 PRMT + 1/2 FFMA2 + 1/2 FADD2 per byte
// instead of PRMT + FFMA + FADD.  Every lane result is the scalar result bit for bit.
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// (2^23 + byte K of wa, 2^23 + byte K of wb) as a float pair
template <int K>
__device__ __forceinline__ uint64_t byte_pair(uint32_t wa, uint32_t wb) {
    return f2_pack(__uint_as_float(__byte_perm(wa, 0x4B000000u, 0x7440 | K)), __uint_as_float(__byte_perm(wb, 0x4B000000u, 0x7440 | K)));
}
#define WICCA_TAP(K, A, B, T, C) h[C] = f2_add(h[C], f2_fma(byte_pair<K>(A, B), w[T], nw[T]))

// Four taps = 12 bytes = three realigned words of row A and of row B; tap t has weight w[t] (the same in both halves).
// Channel chains stay in tap order.
__device__ __forceinline__ void area_group2(const uint32_t*& pa, const uint32_t*& pb, uint32_t& a0, uint32_t& b0, uint32_t shift,
                                            const uint64_t (&w)[4], const uint64_t (&nw)[4], uint64_t (&h)[3]) {
    const uint32_t a1 = pa[1], a2 = pa[2], a3 = pa[3];
    const uint32_t b1 = pb[1], b2 = pb[2], b3 = pb[3];
    const uint32_t x0 = __funnelshift_r(a0, a1, shift), x1 = __funnelshift_r(a1, a2, shift), x2 = __funnelshift_r(a2, a3, shift);
    const uint32_t y0 = __funnelshift_r(b0, b1, shift), y1 = __funnelshift_r(b1, b2, shift), y2 = __funnelshift_r(b2, b3, shift);
    WICCA_TAP(0, x0, y0, 0, 0); WICCA_TAP(1, x0, y0, 0, 1); WICCA_TAP(2, x0, y0, 0, 2);
    WICCA_TAP(3, x0, y0, 1, 0); WICCA_TAP(0, x1, y1, 1, 1); WICCA_TAP(1, x1, y1, 1, 2);
    WICCA_TAP(2, x1, y1, 2, 0); WICCA_TAP(3, x1, y1, 2, 1); WICCA_TAP(0, x2, y2, 2, 2);
    WICCA_TAP(1, x2, y2, 3, 0); WICCA_TAP(2, x2, y2, 3, 1); WICCA_TAP(3, x2, y2, 3, 2);
    a0 = a3; b0 = b3;
    pa += 3; pb += 3;
}
// The last one or two taps of a run as a half group: two taps = 6 bytes (weights w[0], w[1]).  A run is walked as
// [first group, per-tap weights] + whole groups of the uniform weight + [a whole or a half group, per-tap weights], so at
// most one padded (weight-0) tap is ever walked instead of up to three.
__device__ __forceinline__ void area_half2(const uint32_t* pa, const uint32_t* pb, uint32_t a0, uint32_t b0, uint32_t shift,
                                           const uint64_t (&w)[4], const uint64_t (&nw)[4], uint64_t (&h)[3]) {
    const uint32_t a1 = pa[1], a2 = pa[2];
    const uint32_t b1 = pb[1], b2 = pb[2];
    const uint32_t x0 = __funnelshift_r(a0, a1, shift), x1 = __funnelshift_r(a1, a2, shift);
    const uint32_t y0 = __funnelshift_r(b0, b1, shift), y1 = __funnelshift_r(b1, b2, shift);
    WICCA_TAP(0, x0, y0, 0, 0); WICCA_TAP(1, x0, y0, 0, 1); WICCA_TAP(2, x0, y0, 0, 2);
    WICCA_TAP(3, x0, y0, 1, 0); WICCA_TAP(0, x1, y1, 1, 1); WICCA_TAP(1, x1, y1, 1, 2);
}
#undef WICCA_TAP

// kMaxRegs: the register cap of the instantiation.  Shared memory admits four CTAs per SM on the headline shapes; whether
// the registers do depends on the CTA size (one thread per output pixel), and the kernel wants them when it can have
// them: 64 for up to 256 threads, 48 up to 320, 40 above (331 px: 4 CTAs of 352 threads; with 54 registers only three).
template <int kMaxRegs>
__global__ void __maxnreg__(kMaxRegs)
resize_area_rows_kernel(ResizeTables t, int out_h, int out_w, int norm_mode, float* __restrict__ out,
                        uint8_t* __restrict__ out_u8, int buf_bytes) {
    extern __shared__ __align__(16) uint8_t s_rows[];          // two PAIRS of row buffers, buf_bytes each
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
    // the y taps of this output row, in OpenCV's order: [left partial] + full rows + [right partial].  computeResizeAreaTab
    // lists consecutive source rows (the left partial row is s_first - 1, the right one s_first + n_full), so tap k is
    // row row0 + k; its weight is w_left / w_full / w_right by position and 0 for the pad row of an odd count.
    const int has_l = ay.w_left != 0.0f, has_r = ay.w_right != 0.0f;
    const int n_rows = has_l + ay.n_full + has_r;
    const int n_pairs = (n_rows + 1) >> 1;
    const int k_right = has_r ? n_rows - 1 : -1;
    auto row_weight = [&](int k) {
        float b = k < n_rows ? ay.w_full : 0.0f;
        if (has_l && k == 0) b = ay.w_left;
        if (k == k_right) b = ay.w_right;
        return b;
    };
    const uint8_t* g_rows = j.src + (int64_t)(ay.s_first - has_l) * j.pitch + lo;     // row 2p of pair p: + 2p * pitch
    // pair p = source rows 2p and 2p+1 (the last pair of an odd count has one row) into buffers 2(p&1), 2(p&1)+1
    const uint32_t bar0 = rs_smem_u32(&s_full[0]), rows0 = rs_smem_u32(s_rows);
    auto fetch = [&](int p) {
        const int n_here = (2 * p + 1 < n_rows) ? 2 : 1;
        const uint8_t* g = g_rows + (int64_t)(2 * p) * j.pitch;
        uint8_t* s = s_rows + (2 * (p & 1)) * buf_bytes + kRowPadFront;
        if (bulk) {
            if (threadIdx.x == 0) {
                const uint32_t bar = bar0 + 8u * (p & 1), dst = rows0 + (2 * (p & 1)) * buf_bytes + kRowPadFront;
                rs_mbar_expect_tx(bar, (uint32_t)(copy_bytes * n_here));
                rs_bulk_load(dst, g, (uint32_t)copy_bytes, bar);
                if (n_here == 2) rs_bulk_load(dst + buf_bytes, g + j.pitch, (uint32_t)copy_bytes, bar);
            }
        } else {
            for (int q = 0; q < n_here; ++q)
                for (int b = threadIdx.x; b < seg_bytes; b += blockDim.x) s[q * buf_bytes + b] = g[q * j.pitch + b];
        }
    };
    fetch(0);      // thread 0 initialised the barriers itself: the first rows travel while the x taps are set up
    // ---- this thread's x taps (independent of the row): pixels s_first - 1 .. s_first + n_full.  Lanes beyond the end of
    // the output row walk the last pixel's taps again and store nothing, so the row loop has no per-lane branch.
    const int dx = x0 + threadIdx.x;
    const bool active = dx < out_w;
    const bool warp_active = x0 + (int)(threadIdx.x & ~31u) < out_w;          // a segment's last CTA can hold idle warps
    const AreaDesc ax = t.area[j.xoff + min(dx, out_w - 1)];
    const int n_taps = ax.n_full + 2;
    auto tap_weight = [&](int tp) { return tp == 0 ? ax.w_left : (tp <= ax.n_full ? ax.w_full : (tp == ax.n_full + 1 ? ax.w_right : 0.0f)); };
    // A run of more than four taps: first group (w_left, then the uniform weight), `mid` whole groups of the uniform weight,
    // and the last `tail_halves` half groups (one for an odd count of half groups, two for an even one) with per-tap
    // weights.  A run of at most four taps is a tail of two half groups and nothing else.
    const bool has_first = n_taps > 4;
    const int halves = has_first ? (n_taps - 3) >> 1 : 2;
    const int own_tail = 2 - (halves & 1);
    const int mid = (halves - own_tail) >> 1;
    // ... and a whole group for the entire warp as soon as one lane needs it (the two extra taps of the others have
    // weight 0), so that the lanes of a warp never take the two tail forms one after the other
    const int tail_halves = __any_sync(0xFFFFFFFFu, own_tail == 2) ? 2 : 1;
    const int tail_tap = has_first ? 4 + 4 * mid : 0;
    uint64_t wa[4], nwa[4], wz[4], nwz[4], wm[4], nwm[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float fa = q == 0 ? ax.w_left : ax.w_full, fz = tap_weight(tail_tap + q);
        wa[q] = f2_pack(fa, fa); nwa[q] = f2_pack(-8388608.0f * fa, -8388608.0f * fa);
        wz[q] = f2_pack(fz, fz); nwz[q] = f2_pack(-8388608.0f * fz, -8388608.0f * fz);
        wm[q] = f2_pack(ax.w_full, ax.w_full); nwm[q] = f2_pack(-8388608.0f * ax.w_full, -8388608.0f * ax.w_full);
    }
    const int b_start = (ax.s_first - 1) * 3 - lo + kRowPadFront;   // >= 13
    const uint32_t shift = (uint32_t)(b_start & 3) * 8;
    const uint8_t* s_mine = s_rows + (b_start & ~3);
    const int lane = threadIdx.x & 31;
    float my_beta_a = 0.0f, my_beta_b = 0.0f;
    float acc[3] = {0.0f, 0.0f, 0.0f};                          // 0 + x == x: the first y tap needs no special case
    __syncthreads();                                           // mbarrier init visible
    for (int p = 0; p < n_pairs; ++p) {
        if ((p & 31) == 0) {          // lane l keeps the y weights of pair p + l for the next 32 pairs
            my_beta_a = row_weight(2 * (p + lane)); my_beta_b = row_weight(2 * (p + lane) + 1);
        }
        if (p + 1 < n_pairs) fetch(p + 1);
        if (bulk) rs_mbar_wait(bar0 + 8u * (p & 1), (uint32_t)(p >> 1) & 1u);
        else __syncthreads();
        if (warp_active) {
            const uint32_t* pa = reinterpret_cast<const uint32_t*>(s_mine + (2 * (p & 1)) * buf_bytes);
            const uint32_t* pb = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(pa) + (2 * p + 1 < n_rows ? buf_bytes : 0));
            uint32_t a0 = pa[0], b0 = pb[0];
            uint64_t h[3] = {0ull, 0ull, 0ull};                // (+0.0f, +0.0f)
            if (has_first) area_group2(pa, pb, a0, b0, shift, wa, nwa, h);
            for (int g = 0; g < mid; ++g) area_group2(pa, pb, a0, b0, shift, wm, nwm, h);
            if (tail_halves == 2) area_group2(pa, pb, a0, b0, shift, wz, nwz, h);
            else area_half2(pa, pb, a0, b0, shift, wz, nwz, h);
            // sum over rows: acc += beta_a * h_a, then += beta_b * h_b (the pad row of an odd count re-reads row a with weight 0)
            const float beta_a = __shfl_sync(0xFFFFFFFFu, my_beta_a, p & 31), beta_b = __shfl_sync(0xFFFFFFFFu, my_beta_b, p & 31);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float ha, hb;
                f2_unpack(h[c], ha, hb);
                acc[c] = __fadd_rn(__fadd_rn(acc[c], __fmul_rn(beta_a, ha)), __fmul_rn(beta_b, hb));
            }
        }
        __syncthreads();          // the buffers are refilled two iterations later
    }
    if (active) {
        const int64_t i = (((int64_t)img * out_h + dy) * out_w + dx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) emit_value(sat_rint_u8(acc[c]), c, i + c, norm_mode, out, out_u8);
    }
}

struct ResizeRowsLaunch {
    ResizeTables t; int out_h, out_w, norm_mode; float* out; uint8_t* out_u8; int buf_bytes; dim3 grid; int threads; size_t smem;
    cudaStream_t stream;
};
template <int kMaxRegs>
cudaError_t launch_rows(const ResizeRowsLaunch& a) {
    static thread_local int configured_dev = -1;            // the opt-in shared-memory size is set once per (thread, device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(resize_area_rows_kernel<kMaxRegs>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        configured_dev = dev;
    }
    resize_area_rows_kernel<kMaxRegs><<<a.grid, a.threads, a.smem, a.stream>>>(a.t, a.out_h, a.out_w, a.norm_mode, a.out, a.out_u8, a.buf_bytes);
    return cudaGetLastError();
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
    // source span of one segment: seg_w target pixels x scale, + the early pixel, the padded last group, alignment slack
    auto buf_for = [&](int sw) {
        const int sg = (out_w + sw - 1) / sw;
        const int64_t span_px = sg == 1 ? (int64_t)max_src_w + 8 : ((int64_t)sw * max_src_w + out_w - 1) / out_w + 10;
        return kRowPadFront + (int)((span_px * 3 + 16 + 15) / 16 * 16) + kRowPadBack;
    };
    // ... and no wider than lets four CTAs share an SM's shared memory (two pairs of rows each): the rows of a 53 MP
    // source image are 25 KB, and whole-row CTAs ran two to an SM (0.83 ms for 30 images; 0.79 ms in two segments)
    while (seg_w > 64 && (size_t)4 * buf_for(seg_w) > 64 * 1024) seg_w = (seg_w / 2 + 31) / 32 * 32;
    const int segs = (out_w + seg_w - 1) / seg_w;
    const int buf_bytes = buf_for(seg_w);
    const size_t smem = (size_t)4 * buf_bytes;                // two pairs of source rows in flight
    const bool rows_ok = n_area > 0 && out_w <= kRowsMaxThreads && smem <= 200 * 1024 && n <= 65535 && segs <= 65535;
    if (rows_ok) {
        const ResizeRowsLaunch a{t, out_h, out_w, norm_mode, d_out, d_out_u8, buf_bytes, dim3(out_h, n, segs), seg_w, smem, stream};
        cudaError_t e = seg_w <= 256 ? launch_rows<64>(a) : seg_w <= 320 ? launch_rows<48>(a) : launch_rows<40>(a);
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
