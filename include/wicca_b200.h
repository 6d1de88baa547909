/*
 * wicca_b200.h - C ABI of libwicca_b200.so, the B200 (sm_100a) implementation of
 * the WICCA HaarCoder hot path.
 *
 * The reference (Todmount/wicca) is pure Python and has no FFI; the interface
 * each entry point replaces is the Python call it is bound under (see
 * INTEGRATION.md for the ctypes stub a reference maintainer would add):
 *
 *   wicca_haar_icon_u8            HaarCoder.get_small_copy        wicca/wavelet_coder.py:50-67
 *                                   incl. validate_image          wicca/validation.py:80-101
 *                                   and   get_padded_copy         wicca/data_loader.py:66-117
 *   wicca_haar_icons_multi_u8     the per-depth loop that calls it wicca/classifying_tools.py:546-551, :317
 *   wicca_batch_icons_u8          the per-image loop               wicca/classifying_tools.py:312-321
 *   wicca_icon_resize_norm_f32    cv2.resize(icon) + np.stack      wicca/classifying_tools.py:318, :323
 *                                   + preprocess_input / cast      wicca/classifying_tools.py:286-287
 *   wicca_haar_forward_f32 / wicca_haar_inverse_f32
 *                                 extension: all sub-bands of the
 *                                 transform whose LL is wavelet_coder.py:61-65
 *
 * Conventions
 *   - plain C types only; images are uint8, HWC (channels interleaved), row
 *     stride given in bytes; icons are uint8 HWC.
 *   - return 0 on success, a negative WICCA_E* code for argument errors, a
 *     positive value = cudaError_t for CUDA failures.  wicca_last_error()
 *     returns a thread-local message for the last non-zero return.
 *   - every entry point is re-entrant: concurrent calls from different host
 *     threads (the reference calls the coder from a ThreadPoolExecutor,
 *     classifying_tools.py:414-418) use separate streams and scratch buffers.
 *   - there is no CPU fallback: without a usable CUDA device every compute
 *     entry point fails with a cudaError_t.
 */
#ifndef WICCA_B200_H
#define WICCA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WICCA_API __attribute__((visibility("default")))

/* argument errors (negative) */
#define WICCA_EINVAL      (-1)  /* null pointer / non-positive size / bad stride            */
#define WICCA_ECHANNELS   (-2)  /* channel count unsupported for this call                   */
#define WICCA_EBORDER     (-3)  /* border type not one of cv2.BORDER_{CONSTANT,REPLICATE,REFLECT,WRAP,REFLECT_101} */
#define WICCA_EDEPTH      (-4)  /* depth out of the supported range for this call            */
#define WICCA_EDEVICE     (-5)  /* device ordinal out of range                               */
#define WICCA_EALIGN      (-6)  /* device pointer / pitch alignment not met (device-pointer entry points) */
#define WICCA_ENOMEM      (-7)  /* host allocation failed                                    */
#define WICCA_ESTATE      (-8)  /* plan / handle used incorrectly                            */
#define WICCA_EUNSUPPORTED (-9) /* a valid file outside the decoded subset (see wicca_jpeg_probe) */

/* OpenCV border codes accepted by cv2.copyMakeBorder (data_loader.py:116); bit 16 (BORDER_ISOLATED) is ignored */
#define WICCA_BORDER_CONSTANT    0
#define WICCA_BORDER_REPLICATE   1
#define WICCA_BORDER_REFLECT     2
#define WICCA_BORDER_WRAP        3
#define WICCA_BORDER_REFLECT_101 4

/* preprocess_input modes (keras imagenet_utils; classifying_tools.py:286) */
#define WICCA_NORM_IDENTITY 0   /* cast only            (EfficientNet)                     */
#define WICCA_NORM_TF       1   /* x/127.5 - 1          (MobileNetV2, NASNet, Inception*, Xception) */
#define WICCA_NORM_CAFFE    2   /* RGB->BGR, - mean     (VGG, ResNet)                      */
#define WICCA_NORM_TORCH    3   /* x/255, (x-mean)/std  (DenseNet)                         */

#define WICCA_MAX_FUSED_DEPTH 6   /* levels produced by the one-pass tiled kernel            */
#define WICCA_MAX_DEPTH       16  /* deepest transform accepted (levels > 8 are done in fp32 like the reference) */

typedef struct wicca_timing {
    float h2d_ms;      /* host->device copies (CUDA events)                 */
    float kernel_ms;   /* kernels only                                      */
    float d2h_ms;      /* device->host copies                               */
    float total_ms;    /* first enqueue to last completion, on the device   */
} wicca_timing;

/* ---- library / device ------------------------------------------------- */
WICCA_API const char* wicca_version(void);
WICCA_API const char* wicca_last_error(void);
WICCA_API int         wicca_device_count(void);           /* 0 when no CUDA device is usable */
WICCA_API int         wicca_shutdown(void);               /* frees every cached stream / buffer */

/* Row pitch (bytes, multiple of 128) the library uses for device-resident images of width W. */
WICCA_API int64_t wicca_pitch_bytes(int W, int C);
/* Icon extent at a depth: ceil(H / 2^depth), ceil(W / 2^depth) (depth <= 0: H, W). */
WICCA_API int     wicca_icon_dim(int n, int depth);

/* Page-locked host memory for inputs/outputs that should be DMA'd without a bounce copy. */
WICCA_API int wicca_host_alloc(void** ptr, size_t bytes);
/* Same, with the pages placed on the NUMA node of `device` (its PCIe root's local CPUs, from sysfs),
 * so the DMA does not cross the socket interconnect: by first touch when the process may run there, else by
 * mmap + mbind + cudaHostRegister (memory policy needs no CPU on the node).  device < 0: no placement.
 * WICCA_HOST_ALLOC=cuda disables the placement. */
WICCA_API int wicca_host_alloc_near(void** ptr, size_t bytes, int device);
WICCA_API int wicca_host_free(void* ptr);
/* Page-lock memory the caller owns (e.g. a POSIX shared-memory segment that several single-GPU processes gather
 * their icons into - wicca_b200/sharding.py), so results are DMA'd straight into it.  Undo with _unregister. */
WICCA_API int wicca_host_register(void* ptr, size_t bytes);
WICCA_API int wicca_host_unregister(void* ptr);

/* ---- HaarCoder.get_small_copy  (host buffers in, host buffers out) ----- */
/* dst must hold wicca_icon_dim(H,depth) * wicca_icon_dim(W,depth) * C bytes (tight HWC).
 * depth <= 0 copies the image (reference behaviour).  t may be NULL. */
WICCA_API int wicca_haar_icon_u8(const uint8_t* src, int H, int W, int C, int64_t src_row_stride,
                                 int depth, int border_type, double border_const,
                                 uint8_t* dst, int device, wicca_timing* t);

/* Several depths from ONE upload and (for depths 1..6, C == 3) ONE pass over the image. */
WICCA_API int wicca_haar_icons_multi_u8(const uint8_t* src, int H, int W, int C, int64_t src_row_stride,
                                        const int* depths, int n_depths,
                                        int border_type, double border_const,
                                        uint8_t* const* dsts, int device, wicca_timing* t);

/* ---- device-resident variants (benchmark / tensor bridge) -------------- */
/* d_src: device pointer, rows src_pitch bytes apart.  d_dsts[i]: device pointer, rows
 * dst_pitches[i] bytes apart.  The one-pass kernel is used when C == 3, all depths are in
 * 1..6, d_src is 16-byte aligned and src_pitch is a multiple of 16, each dst is 16-byte
 * aligned with a pitch that is a multiple of 16; otherwise the general kernel runs.  On the one-pass
 * path the padding bytes of a destination row between w*C and the next 16-byte boundary (always inside
 * the row's pitch) may be overwritten - TMA store clips at 16-byte granularity; nothing else is touched.
 * stream is a cudaStream_t (NULL = legacy default stream); the call only enqueues. */
WICCA_API int wicca_haar_icons_multi_dev(const uint8_t* d_src, int H, int W, int C, int64_t src_pitch,
                                         const int* depths, int n_depths,
                                         int border_type, double border_const,
                                         uint8_t* const* d_dsts, const int64_t* dst_pitches,
                                         int device, void* stream);

/* ---- batch plans: many device-resident images, one launch -------------- */
typedef struct wicca_plan wicca_plan;
/* All images share C, depths and border.  Icons are allocated by the plan. */
WICCA_API int wicca_plan_create(int device, int n_images, const uint8_t* const* d_srcs,
                                const int* Hs, const int* Ws, const int64_t* src_pitches, int C,
                                const int* depths, int n_depths, int border_type, double border_const,
                                wicca_plan** plan);
WICCA_API int wicca_plan_launch(wicca_plan* plan, void* stream);
/* Device pointer / extents of one icon produced by the plan. */
WICCA_API int wicca_plan_icon(const wicca_plan* plan, int image, int depth_index,
                              uint8_t** d_icon, int* h, int* w, int64_t* pitch);
/* Copy one icon to a tight host buffer (synchronous). */
WICCA_API int wicca_plan_read_icon(const wicca_plan* plan, int image, int depth_index, uint8_t* dst);
/* Fused epilogue on the icons the plan just produced (they never leave the device): for every image,
 * cv2.resize(icon[depth_index], (out_w, out_h), INTER_AREA) -> preprocess_input(norm_mode) -> float32.
 * d_dst: device (n_images, out_h, out_w, 3) float32; d_dst_u8 (nullable): the uint8 batch cv2 would
 * return.  Enqueue after wicca_plan_launch on the same stream.  (classifying_tools.py:318,323,286-287) */
WICCA_API int wicca_plan_resize_norm(wicca_plan* plan, int depth_index, int out_h, int out_w, int norm_mode,
                                     float* d_dst, uint8_t* d_dst_u8, void* stream);
/* Number of kernel launches one wicca_plan_launch issues, and the algorithmic bytes it moves. */
WICCA_API int wicca_plan_info(const wicca_plan* plan, int* launches, int64_t* bytes_read, int64_t* bytes_written);
WICCA_API int wicca_plan_destroy(wicca_plan* plan);

/* ---- sharded host batch: images are independent units ------------------ */
/* srcs[i]: host image i, (Hs[i], Ws[i], C) uint8 with row stride strides[i] (0 = tight).
 * dsts[i * n_depths + k]: tight host icon of image i at depths[k].
 * Image i runs on devices[i % n_devices] (devices == NULL: ordinals 0..n_devices-1); each
 * device has its own worker thread, streams and double-buffered upload slots, so H2D of
 * image i+1 overlaps the kernel of image i.  No inter-GPU traffic.  t (nullable) receives the
 * SUM over images of the per-stage device times. */
WICCA_API int wicca_batch_icons_u8(const uint8_t* const* srcs, const int* Hs, const int* Ws,
                                   const int64_t* strides, int n_images, int C,
                                   const int* depths, int n_depths,
                                   int border_type, double border_const,
                                   uint8_t* const* dsts,
                                   const int* devices, int n_devices, wicca_timing* t);

/* ---- extension: full sub-band transform -------- */
/* Coefficient layout ("pyramid in place"): one float32 HWC plane of the padded size
 * (Hp, Wp, C), Hp = ceil(H/2^depth)*2^depth.  After level l the LL_l block occupies
 * [0,Hp/2^l) x [0,Wp/2^l); HL_l (high-pass along x) sits to its right, LH_l (high-pass along y)
 * below it, HH_l diagonally - the classic Mallat arrangement. */
WICCA_API int wicca_haar_forward_f32(const uint8_t* src, int H, int W, int C, int64_t src_row_stride,
                                     int depth, int border_type, double border_const,
                                     float* coeffs /* host, Hp*Wp*C */, int device, wicca_timing* t);
WICCA_API int wicca_haar_inverse_f32(const float* coeffs /* host, Hp*Wp*C */, int Hp, int Wp, int C,
                                     int depth, float* image /* host, Hp*Wp*C */, int device, wicca_timing* t);
/* Device-resident versions: d_coeffs/d_image are tight (Hp, Wp, C) float32 device planes;
 * d_work is a scratch plane of the same size (ping-pong for levels >= 2). */
WICCA_API int wicca_haar_forward_dev(const uint8_t* d_src, int H, int W, int C, int64_t src_pitch,
                                     int depth, int border_type, double border_const,
                                     float* d_coeffs, float* d_work, int device, void* stream);
WICCA_API int wicca_haar_inverse_dev(const float* d_coeffs, int Hp, int Wp, int C, int depth,
                                     float* d_image, float* d_work, int device, void* stream);

/* ---- epilogue: icon -> INTER_AREA resize -> preprocess_input ----------- */
/* icons[i]: host uint8 HWC icon i (hs[i], ws[i], 3), tight.  dst: host float32 (n, out_h, out_w, 3).
 * dst_u8 (nullable): the intermediate uint8 batch exactly as cv2.resize would return it. */
WICCA_API int wicca_icon_resize_norm_f32(const uint8_t* const* icons, const int* hs, const int* ws, int n,
                                         int out_h, int out_w, int norm_mode,
                                         float* dst, uint8_t* dst_u8, int device, wicca_timing* t);

/* Classifier-ready batches straight from host images: for every image i, icon = get_small_copy(image, depth),
 * dst_icons[i] = preprocess(cv2.resize(icon, (out_w, out_h), INTER_AREA)) and, when dst_images is not NULL,
 * dst_images[i] = preprocess(cv2.resize(image, (out_w, out_h), INTER_AREA)) - the body of
 * ClassifierProcessor._get_img_batch (classifying_tools.py:312-323) plus preprocess_input and the float32
 * cast (:286-287).  Both outputs are host float32 (n_images, out_h, out_w, 3); the icon never leaves the
 * GPU.  3-channel images only; image i runs on devices[i % n_devices] with double-buffered uploads. */
WICCA_API int wicca_batch_classifier_inputs_f32(const uint8_t* const* srcs, const int* Hs, const int* Ws,
                                                const int64_t* strides, int n_images, int depth,
                                                int border_type, double border_const,
                                                int out_h, int out_w, int norm_mode,
                                                float* dst_icons, float* dst_images,
                                                const int* devices, int n_devices, wicca_timing* t);

/* One upload, every (depth, target) batch - row N3 of the hot-path table: the reference re-runs
 * _get_img_batch once per (classifier, depth) pair (loops at classifying_tools.py:546-551 and :339-346), i.e.
 * it reloads, re-transforms and re-resizes every image for every classifier input size and preprocessing
 * family.  Here image i is uploaded once, all `depths` are produced by one pass of the fused icon kernel, and
 * for every target t (out_h, out_w, norm_mode):
 *   dst_icons[t * n_depths + k][i] = preprocess_t(cv2.resize(get_small_copy(image_i, depths[k]), target_t, INTER_AREA))
 *   dst_images[t][i]               = preprocess_t(cv2.resize(image_i, target_t, INTER_AREA))      (dst_images may be NULL)
 * Every destination is a host float32 (n_images, out_h_t, out_w_t, 3) batch.  depths must be >= 1. */
typedef struct wicca_target { int out_h, out_w, norm_mode; } wicca_target;
WICCA_API int wicca_batch_classifier_inputs_multi_f32(const uint8_t* const* srcs, const int* Hs, const int* Ws,
                                                      const int64_t* strides, int n_images,
                                                      const int* depths, int n_depths,
                                                      int border_type, double border_const,
                                                      const wicca_target* targets, int n_targets,
                                                      float* const* dst_icons, float* const* dst_images,
                                                      const int* devices, int n_devices, wicca_timing* t);

/* Device-resident variant for sources that already live in HBM - icons, or the full-size source images of
 * the reference's other branch, cv2.resize(image, shape, interpolation) (classifying_tools.py:315).
 * d_srcs[i]: device uint8 (hs[i], ws[i], 3), rows pitches[i] bytes apart.  Enqueues on `stream`.  The tap tables of a
 * (sources, target) combination are kept on the device (keyed by pointers and geometry, not by content), so repeating the
 * call on a resident batch is the kernel launch alone; the first call of a combination uploads them synchronously. */
WICCA_API int wicca_resize_norm_dev(const uint8_t* const* d_srcs, const int* hs, const int* ws, const int64_t* pitches,
                                    int n, int out_h, int out_w, int norm_mode, float* d_dst, uint8_t* d_dst_u8,
                                    int device, void* stream);

/* ---- other orthogonal wavelets behind WaveletCoder (row N4) ---------------
 * The reference implements only Haar; its README (README.md:25, :222) lists Daubechies / Coiflet coders as the
 * roadmap for the same abstract interface (wicca/wavelet_coder.py:26-38).  This is that get_small_copy for any
 * orthogonal low-pass filter: pad bottom/right to a multiple of 2^depth as get_padded_copy does, then `depth`
 * levels of separable low-pass filtering with periodic wrap-around and decimation by 2, float32, clip and truncate
 * to uint8.  taps: n_taps (even, 2..16) float32 coefficients that sum to 1 (dec_lo / sqrt 2); tap n of output
 * sample k multiplies input sample 2k + n - (n_taps/2 - 1).  With taps {0.5, 0.5} the result equals
 * wicca_haar_icon_u8.  There is no reference implementation to be at parity with for the longer filters: the
 * definition is restated in oracle/fir_oracle.py. */
WICCA_API int wicca_wavelet_icon_u8(const uint8_t* src, int H, int W, int C, int64_t src_row_stride, int depth,
                                    int border_type, double border_const, const float* taps, int n_taps,
                                    uint8_t* dst, int device, wicca_timing* t);

/* ---- JPEG ingest (row N2 of the hot-path table) -------------------------
 * Replaces `cv2.imread(file_path)` + `cv2.cvtColor(image, cv2.COLOR_BGR2RGB)` in load_image
 * (wicca/data_loader.py:53-58) for baseline JPEG files.  The host parses the markers and strips the byte stuffing;
 * Huffman decoding, dequantisation, the inverse DCT, chroma upsampling and YCbCr->RGB run on the GPU with
 * libjpeg-turbo's default arithmetic (islow IDCT, fancy upsampling), so the result is bit-identical to cv2's.
 * Decoded subset: SOF0/SOF1 single-scan Huffman files, 8-bit, 1 (grey, returned as 3 equal channels like
 * IMREAD_COLOR) or 3 (YCbCr) components, integral sampling ratios, restart intervals; the EXIF orientation
 * is applied as cv2.imread applies it (H and W below are those of the oriented image).  Progressive (SOF2) and
 * multi-scan files are not decoded: WICCA_EUNSUPPORTED, from wicca_jpeg_probe already.
 * Everything else returns WICCA_EUNSUPPORTED with the reason in wicca_last_error() - nothing is ever decoded
 * approximately and there is no CPU fallback: route such files through cv2.imread as before. */
WICCA_API int wicca_jpeg_probe(const uint8_t* data, size_t len, int* H, int* W, int* n_components,
                               int* h_max, int* v_max);
/* Number of int16 coefficients of the dense coefficient array (negative = error code): per component, blocks in
 * raster order, 64 quantised coefficients each in natural order. */
WICCA_API int64_t wicca_jpeg_coeff_count(const uint8_t* data, size_t len);
/* Those coefficients from the GPU Huffman decoder (the first stage of every entry point below).  passes (nullable):
 * re-synchronisation passes it took.  For tests and measurements. */
WICCA_API int wicca_jpeg_decode_coeffs_gpu(const uint8_t* data, size_t len, int16_t* dst, int64_t dst_count, int device,
                                           int* passes);
/* JPEG bytes -> host RGB image (H, W, 3), rows dst_stride bytes apart (0 = tight).  host_decode_ms (nullable):
 * wall time of the Huffman stage; t: device stages. */
WICCA_API int wicca_jpeg_decode_u8(const uint8_t* data, size_t len, uint8_t* dst, int64_t dst_stride, int device,
                                   wicca_timing* t, float* host_decode_ms);
/* JPEG bytes -> device RGB image, rows d_pitch bytes apart (use wicca_pitch_bytes for the fused icon path).
 * Work is enqueued on `stream` and the call returns after it has completed (its staging buffers are pooled). */
WICCA_API int wicca_jpeg_decode_dev(const uint8_t* data, size_t len, uint8_t* d_dst, int64_t d_pitch, int device,
                                    void* stream);
/* JPEG bytes -> icons: load_image + get_small_copy for every depth (>= 1) without the RGB image ever existing on
 * the host.  dsts[k]: host (icon_dim(H, d_k), icon_dim(W, d_k), 3) uint8, tight. */
WICCA_API int wicca_jpeg_icons_multi_u8(const uint8_t* data, size_t len, const int* depths, int n_depths,
                                        int border_type, double border_const, uint8_t* const* dsts, int device,
                                        wicca_timing* t, float* host_decode_ms);
/* Files -> every classifier-ready batch: wicca_batch_classifier_inputs_multi_f32 with load_image folded in, i.e. the
 * whole of ClassifierProcessor._get_img_batch (classifying_tools.py:312-323, which starts from file paths) plus
 * preprocess_input for every classifier input and depth, with nothing but the JPEG bytes crossing PCIe. */
WICCA_API int wicca_batch_classifier_inputs_multi_from_jpeg(const uint8_t* const* datas, const size_t* lens, int n_images,
                                                            const int* depths, int n_depths,
                                                            int border_type, double border_const,
                                                            const wicca_target* targets, int n_targets,
                                                            float* const* dst_icons, float* const* dst_images,
                                                            const int* devices, int n_devices, wicca_timing* t);
/* The same for n files: dsts[i * n_depths + k]; file i runs on devices[i % n_devices]; n_threads host threads
 * (0 = one per core, at most 32) each take the next file - Huffman decoding is the bottleneck and scales with cores. */
WICCA_API int wicca_batch_icons_from_jpeg(const uint8_t* const* datas, const size_t* lens, int n_images,
                                          const int* depths, int n_depths, int border_type, double border_const,
                                          uint8_t* const* dsts, const int* devices, int n_devices, int n_threads,
                                          float* host_decode_ms);

#ifdef __cplusplus
}
#endif
#endif /* WICCA_B200_H */
