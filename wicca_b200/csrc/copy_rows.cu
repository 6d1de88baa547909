// copy_rows.cu - device-to-device copy between row pitches (tight rows <-> the 128-byte pitched layout of section 3 of
// DESIGN.md).  Everything large crosses the host link as ONE FLAT copy (host_common.h: upload_image_async /
// download_rows_async); this kernel does the re-pitching on the device side of it, at HBM speed - a
// cudaMemcpy2DAsync(DeviceToDevice) of the same 53 MB of icon rows took 0.5 ms, which a one-image call feels.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace wicca {

// One thread per (row, 4-byte chunk of the row).  A chunk moves as one 32-bit word when its source / destination
// address is word aligned (always so on the pitched side; on the tight side when row_bytes is a multiple of 4), else
// byte by byte; the last chunk of a row may be partial.
__global__ void __launch_bounds__(256)
copy_rows_kernel(uint8_t* __restrict__ dst, int64_t dpitch, const uint8_t* __restrict__ src, int64_t spitch, uint32_t row_bytes,
                 uint32_t rows, uint32_t chunks) {
    const uint64_t total = (uint64_t)rows * chunks;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(i / chunks), c = (uint32_t)(i - (uint64_t)r * chunks);
        const uint8_t* s = src + (int64_t)r * spitch + 4 * (int64_t)c;
        uint8_t* d = dst + (int64_t)r * dpitch + 4 * (int64_t)c;
        const uint32_t n = min(4u, row_bytes - 4 * c);
        uint32_t w;
        if (n == 4 && ((uintptr_t)s & 3) == 0) {
            w = *reinterpret_cast<const uint32_t*>(s);
        } else {
            w = 0;
            for (uint32_t k = 0; k < n; ++k) w |= (uint32_t)s[k] << (8 * k);
        }
        if (n == 4 && ((uintptr_t)d & 3) == 0) {
            *reinterpret_cast<uint32_t*>(d) = w;
        } else {
            for (uint32_t k = 0; k < n; ++k) d[k] = (uint8_t)(w >> (8 * k));
        }
    }
}

cudaError_t launch_copy_rows(void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t row_bytes, int rows,
                             cudaStream_t stream) {
    if (row_bytes <= 0 || rows <= 0) return cudaSuccess;
    if (row_bytes > 0x7FFFFFF0ll) return cudaErrorInvalidValue;
    const uint32_t chunks = (uint32_t)((row_bytes + 3) / 4);
    const uint64_t total = (uint64_t)rows * chunks;
    uint64_t blocks = (total + 255) / 256;
    if (blocks > 148ull * 64) blocks = 148ull * 64;            // grid-stride beyond 64 CTAs' worth per SM
    copy_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>((uint8_t*)dst, dpitch, (const uint8_t*)src, spitch, (uint32_t)row_bytes,
                                                           (uint32_t)rows, chunks);
    return cudaGetLastError();
}

}  // namespace wicca
