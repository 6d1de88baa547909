// capi_fir.cu - C entry point for the longer orthogonal wavelets (row N4), host image in / host icon out,
// same conventions as wicca_haar_icon_u8.
#include <string.h>

#include "host_common.h"
#include "kernels.h"

using namespace wicca;

extern "C" int wicca_wavelet_icon_u8(const uint8_t* src, int H, int W, int C, int64_t src_row_stride, int depth,
                                     int border_type, double border_const, const float* taps, int n_taps, uint8_t* dst,
                                     int device, wicca_timing* t) {
    if (t) memset(t, 0, sizeof(*t));
    int rc = validate_icon_args(src, H, W, C, &depth, 1, border_type);
    if (rc) return rc;
    if (!dst) return fail(WICCA_EINVAL, "dst is NULL");
    if (!taps || n_taps < 2 || n_taps > 16 || (n_taps & 1)) return fail(WICCA_EINVAL, "need an even number of taps between 2 and 16");
    if (depth > 16) return fail(WICCA_EDEPTH, "transform depth above 16");
    const int64_t rowb = (int64_t)W * C;
    if (src_row_stride == 0) src_row_stride = rowb;
    if (src_row_stride < rowb) return fail(WICCA_EINVAL, "src_row_stride < W*C");
    if (depth <= 0) {                                   // the reference returns the image itself for depth 0
        for (int y = 0; y < H; ++y) memcpy(dst + (size_t)y * rowb, src + (size_t)y * src_row_stride, (size_t)rowb);
        return 0;
    }
    rc = check_device(device);
    if (rc) return rc;
    CtxLease lease;
    rc = acquire_ctx(device, &lease.c);
    if (rc) return rc;
    Ctx& c = *lease.c;
    FirTaps ft;
    memset(&ft, 0, sizeof ft);
    ft.n = n_taps; ft.c = n_taps / 2 - 1;
    for (int k = 0; k < n_taps; ++k) ft.g[k] = taps[k];
    const int64_t ratio = (int64_t)1 << depth;
    const int64_t Hp = (H + ratio - 1) / ratio * ratio, Wp = (W + ratio - 1) / ratio * ratio;
    const int ih = (int)(Hp >> depth), iw = (int)(Wp >> depth);
    const int64_t pitch = wicca_pitch_bytes(W, C);
    WICCA_CUDA(c.d_src.reserve((size_t)pitch * H + 256));
    WICCA_CUDA(c.d_f32a.reserve((size_t)(Hp * (Wp / 2) * C) * sizeof(float)));
    WICCA_CUDA(c.d_f32b.reserve((size_t)((Hp / 2) * (Wp / 2) * C) * sizeof(float)));
    WICCA_CUDA(c.d_icons.reserve((size_t)ih * iw * C + 256));
    WICCA_CUDA(cudaEventRecord(c.ev[0], c.stream));
    rc = upload_image_async(c, src, H, rowb, src_row_stride, pitch);
    if (rc) { cudaStreamSynchronize(c.stream); c.pending.clear(); return rc; }
    WICCA_CUDA(cudaEventRecord(c.ev[1], c.stream));
    cudaError_t e = launch_wavelet_fir((const uint8_t*)c.d_src.p, pitch, H, W, C, depth, border_type & ~16, saturate_u8(border_const),
                                       ft, (uint8_t*)c.d_icons.p, (int64_t)iw * C, (float*)c.d_f32a.p, (float*)c.d_f32b.p, c.stream);
    if (e != cudaSuccess) { cudaStreamSynchronize(c.stream); return cuda_fail(e, "wavelet FIR kernels"); }
    WICCA_CUDA(cudaEventRecord(c.ev[2], c.stream));
    const size_t icon_bytes = (size_t)ih * iw * C;
    uint8_t* target = dst;
    const bool direct = is_pinned_host(dst);
    if (!direct) {
        WICCA_CUDA(c.h_bounce.reserve(icon_bytes));
        target = (uint8_t*)c.h_bounce.p;
    }
    WICCA_CUDA(cudaMemcpyAsync(target, c.d_icons.p, icon_bytes, cudaMemcpyDeviceToHost, c.stream));
    WICCA_CUDA(cudaEventRecord(c.ev[3], c.stream));
    WICCA_CUDA(cudaStreamSynchronize(c.stream));
    if (!direct) memcpy(dst, target, icon_bytes);
    if (t) {
        cudaEventElapsedTime(&t->h2d_ms, c.ev[0], c.ev[1]);
        cudaEventElapsedTime(&t->kernel_ms, c.ev[1], c.ev[2]);
        cudaEventElapsedTime(&t->d2h_ms, c.ev[2], c.ev[3]);
        cudaEventElapsedTime(&t->total_ms, c.ev[0], c.ev[3]);
    }
    return 0;
}
