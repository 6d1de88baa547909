// capi_subband.cu - C ABI of the full sub-band forward / inverse transform (extension A4).
#include <string.h>

#include "host_common.h"
#include "kernels.h"

using namespace wicca;

namespace {
int check_subband_args(int H, int W, int C, int depth) {
    if (H <= 0 || W <= 0 || C <= 0) return fail(WICCA_EINVAL, "empty image (H=%d W=%d C=%d)", H, W, C);
    if (depth < 1 || depth > WICCA_MAX_DEPTH) return fail(WICCA_EDEPTH, "depth %d outside 1..%d", depth, WICCA_MAX_DEPTH);
    return 0;
}
size_t work_floats(int Hp, int Wp, int C) {
    const size_t n = (size_t)Hp * Wp * C;
    return n / 4 + n / 16 + 64;
}
}  // namespace

extern "C" {

int wicca_haar_forward_dev(const uint8_t* d_src, int H, int W, int C, int64_t src_pitch, int depth, int border_type,
                           double border_const, float* d_coeffs, float* d_work, int device, void* stream) {
    int rc = check_subband_args(H, W, C, depth);
    if (rc) return rc;
    if (!d_src || !d_coeffs || !d_work) return fail(WICCA_EINVAL, "null device pointer");
    if (src_pitch < (int64_t)W * C) return fail(WICCA_EINVAL, "src_pitch < W*C");
    const int Hp = icon_dim(H, depth) << depth, Wp = icon_dim(W, depth) << depth;
    if ((Hp != H || Wp != W) && !border_valid(border_type)) return fail(WICCA_EBORDER, "unsupported border type %d", border_type);
    rc = check_device(device);
    if (rc) return rc;
    WICCA_CUDA(cudaSetDevice(device));
    cudaError_t e = launch_forward(d_src, src_pitch, H, W, C, Hp, Wp, depth, border_base(border_type),
                                   saturate_u8(border_const), d_coeffs, d_work, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "forward sub-band kernels");
    return 0;
}

int wicca_haar_inverse_dev(const float* d_coeffs, int Hp, int Wp, int C, int depth, float* d_image, float* d_work,
                           int device, void* stream) {
    int rc = check_subband_args(Hp, Wp, C, depth);
    if (rc) return rc;
    if (!d_coeffs || !d_image || !d_work) return fail(WICCA_EINVAL, "null device pointer");
    if ((Hp % (1 << depth)) || (Wp % (1 << depth))) return fail(WICCA_EINVAL, "coefficient plane %dx%d is not a multiple of 2^%d", Hp, Wp, depth);
    rc = check_device(device);
    if (rc) return rc;
    WICCA_CUDA(cudaSetDevice(device));
    cudaError_t e = launch_inverse(d_coeffs, Hp, Wp, C, depth, d_image, d_work, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "inverse sub-band kernels");
    return 0;
}

int wicca_haar_forward_f32(const uint8_t* src, int H, int W, int C, int64_t src_row_stride, int depth, int border_type,
                           double border_const, float* coeffs, int device, wicca_timing* t) {
    int rc = check_subband_args(H, W, C, depth);
    if (rc) return rc;
    if (!src || !coeffs) return fail(WICCA_EINVAL, "null pointer");
    const int64_t rowb = (int64_t)W * C;
    if (src_row_stride == 0) src_row_stride = rowb;
    if (src_row_stride < rowb) return fail(WICCA_EINVAL, "row stride < W*C");
    const int Hp = icon_dim(H, depth) << depth, Wp = icon_dim(W, depth) << depth;
    if ((Hp != H || Wp != W) && !border_valid(border_type)) return fail(WICCA_EBORDER, "unsupported border type %d", border_type);
    if (t) memset(t, 0, sizeof(*t));
    CtxLease L;
    rc = acquire_ctx(device, &L.c);
    if (rc) return rc;
    Ctx& c = *L.c;
    const int64_t pitch = wicca_pitch_bytes(W, C);
    const size_t plane_bytes = (size_t)Hp * Wp * C * sizeof(float);
    WICCA_CUDA(c.d_src.reserve((size_t)pitch * H + 256));
    WICCA_CUDA(c.d_f32a.reserve(plane_bytes));
    WICCA_CUDA(c.d_f32b.reserve(work_floats(Hp, Wp, C) * sizeof(float)));
    WICCA_CUDA(cudaEventRecord(c.ev[0], c.stream));
    rc = upload_image_async(c, src, H, rowb, src_row_stride, pitch);
    if (rc) { cudaStreamSynchronize(c.stream); return rc; }
    WICCA_CUDA(cudaEventRecord(c.ev[1], c.stream));
    cudaError_t e = launch_forward((const uint8_t*)c.d_src.p, pitch, H, W, C, Hp, Wp, depth, border_base(border_type),
                                   saturate_u8(border_const), (float*)c.d_f32a.p, (float*)c.d_f32b.p, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "forward sub-band kernels");
    WICCA_CUDA(cudaEventRecord(c.ev[2], c.stream));
    WICCA_CUDA(cudaMemcpyAsync(coeffs, c.d_f32a.p, plane_bytes, cudaMemcpyDeviceToHost, c.stream));
    WICCA_CUDA(cudaEventRecord(c.ev[3], c.stream));
    WICCA_CUDA(cudaStreamSynchronize(c.stream));
    if (t) {
        cudaEventElapsedTime(&t->h2d_ms, c.ev[0], c.ev[1]);
        cudaEventElapsedTime(&t->kernel_ms, c.ev[1], c.ev[2]);
        cudaEventElapsedTime(&t->d2h_ms, c.ev[2], c.ev[3]);
        cudaEventElapsedTime(&t->total_ms, c.ev[0], c.ev[3]);
    }
    return 0;
}

int wicca_haar_inverse_f32(const float* coeffs, int Hp, int Wp, int C, int depth, float* image, int device,
                           wicca_timing* t) {
    int rc = check_subband_args(Hp, Wp, C, depth);
    if (rc) return rc;
    if (!coeffs || !image) return fail(WICCA_EINVAL, "null pointer");
    if ((Hp % (1 << depth)) || (Wp % (1 << depth))) return fail(WICCA_EINVAL, "coefficient plane %dx%d is not a multiple of 2^%d", Hp, Wp, depth);
    if (t) memset(t, 0, sizeof(*t));
    CtxLease L;
    rc = acquire_ctx(device, &L.c);
    if (rc) return rc;
    Ctx& c = *L.c;
    const size_t plane_bytes = (size_t)Hp * Wp * C * sizeof(float);
    WICCA_CUDA(c.d_f32a.reserve(plane_bytes));
    WICCA_CUDA(c.d_f32b.reserve(work_floats(Hp, Wp, C) * sizeof(float)));
    WICCA_CUDA(c.d_misc.reserve(plane_bytes));
    WICCA_CUDA(cudaEventRecord(c.ev[0], c.stream));
    WICCA_CUDA(cudaMemcpyAsync(c.d_f32a.p, coeffs, plane_bytes, cudaMemcpyHostToDevice, c.stream));
    WICCA_CUDA(cudaEventRecord(c.ev[1], c.stream));
    cudaError_t e = launch_inverse((const float*)c.d_f32a.p, Hp, Wp, C, depth, (float*)c.d_misc.p, (float*)c.d_f32b.p, c.stream);
    if (e != cudaSuccess) return cuda_fail(e, "inverse sub-band kernels");
    WICCA_CUDA(cudaEventRecord(c.ev[2], c.stream));
    WICCA_CUDA(cudaMemcpyAsync(image, c.d_misc.p, plane_bytes, cudaMemcpyDeviceToHost, c.stream));
    WICCA_CUDA(cudaEventRecord(c.ev[3], c.stream));
    WICCA_CUDA(cudaStreamSynchronize(c.stream));
    if (t) {
        cudaEventElapsedTime(&t->h2d_ms, c.ev[0], c.ev[1]);
        cudaEventElapsedTime(&t->kernel_ms, c.ev[1], c.ev[2]);
        cudaEventElapsedTime(&t->d2h_ms, c.ev[2], c.ev[3]);
        cudaEventElapsedTime(&t->total_ms, c.ev[0], c.ev[3]);
    }
    return 0;
}

}  // extern "C"
