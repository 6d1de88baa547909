// kernels.h - launch entry points of the .cu translation units (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "icon_types.h"

namespace wicca {

// haar_icon.cu
// stream_hint: 1 = mark the input L2::evict_first (use when no level <= 2 is written)
cudaError_t launch_icon_tma(const IconImage* d_imgs, const uint8_t* const* d_strips, int n_images, int total_items,
                            int border_type, int border_const, int sm_count, int variant, int stream_hint,
                            cudaStream_t stream);
cudaError_t launch_edge_strips(const IconImage* d_imgs, uint8_t* const* d_strips, int n_images, int max_rows,
                               int border_type, int border_const, cudaStream_t stream);
cudaError_t launch_icon_generic(const GenericIconArgs& a, cudaStream_t stream);
// haar_rows.cu
int rows_kernel_groups(int C, int depth);            // 0: haar_icon_rows_kernel cannot take this (C, depth)
cudaError_t launch_icon_rows(const GenericIconArgs& a, cudaStream_t stream);
cudaError_t launch_icon_tail_fill(const TailArgs& a, cudaStream_t stream);   // once per image, before the tails
cudaError_t launch_icon_tail(const TailArgs& a, cudaStream_t stream);
cudaError_t launch_level_f32(const float* in, float* out, uint8_t* out_u8, int out_h, int out_w, int C,
                             cudaStream_t stream);

// haar_subband.cu   (d_work: scratch of >= Hp*Wp*C*5/16 floats)
cudaError_t launch_forward(const uint8_t* d_src, int64_t pitch, int H, int W, int C, int Hp, int Wp, int depth,
                           int border_type, int border_const, float* d_coeffs, float* d_work, cudaStream_t stream);
cudaError_t launch_inverse(const float* d_coeffs, int Hp, int Wp, int C, int depth, float* d_image, float* d_work,
                           cudaStream_t stream);

// resize_norm.cu
struct ResizeJob {           // one icon of a batch
    const uint8_t* src;      // device, (sh, sw, 3) uint8, rows `pitch` bytes apart
    int64_t pitch;
    int sh, sw;
    int regime;              // 0 = same size (copy), 1 = integer-factor area, 2 = general area, 3 = bilinear "area mode"
    int isx, isy;            // regime 1: integer scale factors
    int xoff, yoff;          // regime 2: offsets into the AreaDesc array (out_w / out_h entries)
                             // regime 3: offsets into the bilinear tap array (out_w / out_h entries)
};
// computeResizeAreaTab for one destination index, in closed form: OpenCV's tap list is always
// [left partial pixel] + n_full whole pixels of equal weight + [right partial pixel], in that order.
// A missing partial tap has weight 0 (it is skipped, which is exact: x + 0 == x).
struct AreaDesc { int s_left, s_first, n_full, s_right; float w_left, w_full, w_right; int pad; };
struct LinTap { int i0, i1, c0, c1; };
struct ResizeTables {
    const ResizeJob* jobs;
    const AreaDesc* area;
    const LinTap* lin;
};
cudaError_t launch_resize_norm(const ResizeTables& t, int n, int out_h, int out_w, int norm_mode, float* d_out,
                               uint8_t* d_out_u8, int max_src_w, int n_area, int n_other, cudaStream_t stream);

// ---- copy_rows.cu: device-to-device copy between row pitches (the device side of the flat host-link copies)
cudaError_t launch_copy_rows(void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t row_bytes, int rows,
                             cudaStream_t stream);

// ---- other orthogonal wavelets (wavelet_fir.cu): LL pyramid with a longer low-pass filter, periodic per level
struct FirTaps { float g[16]; int n, c; };     // taps (sum 1), their count (even, <= 16), centre offset n/2 - 1
cudaError_t launch_wavelet_fir(const uint8_t* d_src, int64_t pitch, int H, int W, int C, int depth, int border_type,
                               int border_const, const FirTaps& taps, uint8_t* d_icon, int64_t icon_pitch, float* d_rows,
                               float* d_ll, cudaStream_t stream);

// ---- JPEG ingest (jpeg_kernels.cu): dense quantised coefficients -> pitched RGB image, libjpeg-turbo arithmetic
struct JpegPlaneDesc {
    const int16_t* coefs;     // (blocks_h, blocks_w, 64) natural order, quantised
    uint8_t* plane;           // (blocks_h * 8, plane_pitch) samples
    int blocks_w, blocks_h, plane_pitch;
    int dw, dh;               // real samples of the component
    int hf, vf;               // expansion factors to the image grid
    int mode;                 // 0 none, 1 h2v1 fancy, 2 h2v2 fancy, 3 h1v2 fancy, 4 box
    uint16_t qt[64];          // natural order
};
struct JpegImageDesc {
    JpegPlaneDesc comp[3];
    int ncomp, width, height;
    uint8_t* dst; int64_t dst_pitch;     // RGB, HWC
};
cudaError_t launch_jpeg_decode(const JpegImageDesc& d, cudaStream_t stream);
// EXIF orientation 2..8 as cv2.imread applies it (flip / transpose + flip): src (H, W, 3) -> dst, which is (W, H, 3)
// for orientations >= 5.
cudaError_t launch_jpeg_orient(const uint8_t* d_src, int64_t src_pitch, int H, int W, int orientation, uint8_t* d_dst,
                               int64_t dst_pitch, cudaStream_t stream);

}  // namespace wicca
