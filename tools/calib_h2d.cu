// calib_h2d.cu - how fast does this box move one 6393 x 8284 x 3 image from page-locked memory to the GPU:
// as a pitched 2-D copy (what the ingest path does: rows of 24,852 bytes into a 24,960-byte pitch), as one
// contiguous 1-D copy, and as a 2-D copy whose source rows are padded to the pitch as well.
// A fourth line cycles through 12 distinct page-locked images (1.9 GB) instead of re-sending one: what a batch does.
// nvcc -O2 -o gpurun_out/calib_h2d tools/calib_h2d.cu && gpurun_out/calib_h2d
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main() {
    const size_t H = 6393, W = 8284, rowb = W * 3, pitch = (rowb + 127) / 128 * 128;
    const int n = 30;
    unsigned char *h = nullptr, *d = nullptr;
    CK(cudaHostAlloc((void**)&h, pitch * H, cudaHostAllocDefault));
    memset(h, 1, pitch * H);
    CK(cudaMalloc((void**)&d, pitch * H + 256));
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int kDistinct = 12;
    unsigned char* hs[kDistinct];
    for (int k = 0; k < kDistinct; ++k) { CK(cudaHostAlloc((void**)&hs[k], rowb * H, cudaHostAllocDefault)); memset(hs[k], k, rowb * H); }
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0, s));
            for (int i = 0; i < n; ++i) {
                if (mode == 0) CK(cudaMemcpy2DAsync(d, pitch, h, rowb, rowb, H, cudaMemcpyHostToDevice, s));
                else if (mode == 1) CK(cudaMemcpyAsync(d, h, rowb * H, cudaMemcpyHostToDevice, s));
                else if (mode == 2) CK(cudaMemcpy2DAsync(d, pitch, h, pitch, rowb, H, cudaMemcpyHostToDevice, s));
                else CK(cudaMemcpy2DAsync(d, pitch, hs[i % kDistinct], rowb, rowb, H, cudaMemcpyHostToDevice, s));
            }
            CK(cudaEventRecord(e1, s));
            CK(cudaStreamSynchronize(s));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep) printf("%s: %.3f ms per image, %.1f GB/s\n", mode == 0 ? "2-D, dense source rows -> pitched" : mode == 1 ? "1-D contiguous" : mode == 2 ? "2-D, pitched source -> pitched" : "2-D, 12 distinct source images in turn", ms / n, rowb * H * n / ms / 1e6);
        }
    }
    return 0;
}
