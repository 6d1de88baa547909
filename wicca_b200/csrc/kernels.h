// kernels.h - launch entry points of the .cu translation units (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "icon_types.h"

namespace wicca {

// haar_icon.cu
cudaError_t launch_icon_tma(const IconImage* d_imgs, const uint8_t* const* d_strips, int n_images, int total_items,
                            int border_type, int border_const, int sm_count, int variant, cudaStream_t stream);
cudaError_t launch_edge_strips(const IconImage* d_imgs, uint8_t* const* d_strips, int n_images, int max_rows,
                               int border_type, int border_const, cudaStream_t stream);
cudaError_t launch_icon_generic(const GenericIconArgs& a, cudaStream_t stream);
cudaError_t launch_level_f32(const float* in, float* out, uint8_t* out_u8, int out_h, int out_w, int C,
                             cudaStream_t stream);

// haar_subband.cu
cudaError_t launch_forward_level1_u8(const uint8_t* src, int64_t pitch, int H, int W, int C, int Hp, int Wp,
                                     int border_type, int border_const, float* coeffs, cudaStream_t stream);
cudaError_t launch_forward_level_f32(const float* ll_in, int64_t in_row_elems, float* coeffs, int64_t co_row_elems,
                                     int out_h, int out_w, int C, cudaStream_t stream);
cudaError_t launch_inverse_level_f32(const float* coeffs, int64_t co_row_elems, const float* ll_in,
                                     int64_t ll_row_elems, float* out, int64_t out_row_elems, int h, int w, int C,
                                     cudaStream_t stream);

// resize_norm.cu
struct ResizeTap { int dst; int src; float w; };
cudaError_t launch_resize_area_generic(const uint8_t* const* d_icons, const int* d_hs, const int* d_ws, int n,
                                       const ResizeTap* d_xtabs, const int* d_xtab_off, const ResizeTap* d_ytabs,
                                       const int* d_ytab_off, int out_h, int out_w, int norm_mode, float* d_out,
                                       uint8_t* d_out_u8, float* d_hbuf, const int64_t* d_hbuf_off,
                                       cudaStream_t stream);

}  // namespace wicca
