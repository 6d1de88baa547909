// calib_read.cu - developer calibration: how fast can one B200 stream a pitched image out of HBM
// (a) with plain 128-bit loads, (b) with TMA box loads of different shapes / ring depths,
// without any arithmetic.  Gives the ceiling the icon kernel is compared against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_build/calib_read tools/calib_read.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}

__global__ void read_ldg(const uint4* __restrict__ p, size_t n16, uint32_t* sink) {
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n16; i += 8 * stride) {
        uint4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(p + i + k * stride));
#pragma unroll
        for (int k = 0; k < 8; ++k) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    }
    for (; i < n16; i += stride) { uint4 v = p[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) sink[0] = acc;
}

// persistent TMA reader: warp 0 lane 0 produces, warp 1 consumes (touches one word per lane, releases)
__global__ void __launch_bounds__(64, 1)
read_tma(const __grid_constant__ CUtensorMap tmap, int boxes_x, int boxes_y, int box_w_elems, int box_h, int stage_bytes, int stages, int hint, uint32_t* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = (uint64_t*)(smem + (size_t)stages * stage_bytes);
    uint64_t* empty = full + stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int total = boxes_x * boxes_y;
    if (threadIdx.x == 0) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        int k = 0;
        for (int g = blockIdx.x; g < total; g += gridDim.x, ++k) {
            const int s = k % stages; const uint32_t ph = (k / stages) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            const int by = g / boxes_x, bx = g - by * boxes_x;
            mbar_expect(&full[s], stage_bytes);
            if (hint)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                             ::"r"(smem_u32(smem + (size_t)s * stage_bytes)), "l"(&tmap), "r"(bx * box_w_elems), "r"(by * box_h), "r"(smem_u32(&full[s])), "l"(pol) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(smem + (size_t)s * stage_bytes)), "l"(&tmap), "r"(bx * box_w_elems), "r"(by * box_h), "r"(smem_u32(&full[s])) : "memory");
        }
    } else if (threadIdx.x >= 32) {
        uint32_t acc = 0; int k = 0;
        for (int g = blockIdx.x; g < total; g += gridDim.x, ++k) {
            const int s = k % stages; const uint32_t ph = (k / stages) & 1;
            mbar_wait(&full[s], ph);
            acc ^= *(const uint32_t*)(smem + (size_t)s * stage_bytes + (threadIdx.x - 32) * 4);
            __syncwarp();
            if (threadIdx.x == 32) mbar_arrive(&empty[s]);
        }
        if (acc == 0x12345678u) sink[1] = acc;
    }
}

// read 384x64 boxes and write a 192x32 box (the level-1 icon's share) per item with TMA store, no arithmetic
__global__ void __launch_bounds__(64, 1)
read_write_tma(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap omap, int boxes_x, int boxes_y, int stages, int contiguous, uint32_t* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int stage_bytes = 24576;
    uint64_t* full = (uint64_t*)(smem + (size_t)stages * stage_bytes);
    uint64_t* empty = full + stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int total = boxes_x * boxes_y;
    const int per = (total + gridDim.x - 1) / gridDim.x;
    const int g0 = contiguous ? blockIdx.x * per : blockIdx.x;
    const int g1 = contiguous ? min(total, g0 + per) : total;
    const int gs = contiguous ? 1 : gridDim.x;
    if (threadIdx.x == 0) {
        int k = 0;
        for (int g = g0; g < g1; g += gs, ++k) {
            const int s = k % stages; const uint32_t ph = (k / stages) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            const int by = g / boxes_x, bx = g - by * boxes_x;
            mbar_expect(&full[s], stage_bytes);
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(smem + (size_t)s * stage_bytes)), "l"(&tmap), "r"(bx * 48), "r"(by * 64), "r"(smem_u32(&full[s])) : "memory");
        }
    } else if (threadIdx.x == 32) {
        int k = 0;
        for (int g = g0; g < g1; g += gs, ++k) {
            const int s = k % stages; const uint32_t ph = (k / stages) & 1;
            mbar_wait(&full[s], ph);
            const int by = g / boxes_x, bx = g - by * boxes_x;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&omap), "r"(bx * 192), "r"(by * 32), "r"(smem_u32(smem + (size_t)s * stage_bytes)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_arrive(&empty[s]);
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int64_t pitch = 24960; const int64_t rows = 6393LL * 30;
    const size_t bytes = (size_t)pitch * rows;
    uint8_t* d; CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 1, bytes));
    uint32_t* sink; CK(cudaMalloc(&sink, 64));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](auto fn, const char* name) {
        for (int i = 0; i < 2; ++i) fn();
        CK(cudaDeviceSynchronize());
        std::vector<float> ts;
        for (int i = 0; i < 5; ++i) { CK(cudaEventRecord(e0)); fn(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ts.push_back(ms); }
        CK(cudaGetLastError());
        std::sort(ts.begin(), ts.end());
        printf("%-44s med %.4f ms  %.1f GB/s   best %.1f GB/s\n", name, ts[2], bytes / ts[2] / 1e6, bytes / ts[0] / 1e6);
        fflush(stdout);
    };
    for (int mult : {2, 4, 8, 16}) {
        char nm[64]; snprintf(nm, sizeof nm, "ldg.128 grid=148x%d x256thr", mult);
        timeit([&] { read_ldg<<<148 * mult, 256>>>((const uint4*)d, bytes / 16, sink); }, nm);
    }
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fp;
    struct Shape { int w_bytes, h; };
    CK(cudaFuncSetAttribute(read_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    for (Shape sh : {Shape{384, 64}, Shape{768, 32}, Shape{1536, 16}, Shape{1920, 8}, Shape{384, 32}, Shape{768, 64}}) {
        for (int stages : {2, 4, 8}) {
            for (int hint : {0, 1}) {
                const int esz = 8;   // UINT64 elements so boxes up to 2048 B wide are legal
                CUtensorMap tm;
                cuuint64_t gdim[2] = {(cuuint64_t)(pitch / esz), (cuuint64_t)rows};
                cuuint64_t gstr[1] = {(cuuint64_t)pitch};
                cuuint32_t box[2] = {(cuuint32_t)(sh.w_bytes / esz), (cuuint32_t)sh.h};
                cuuint32_t es[2] = {1, 1};
                CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { printf("encode failed %d for %dx%d\n", (int)r, sh.w_bytes, sh.h); continue; }
                const int stage_bytes = sh.w_bytes * sh.h;
                if ((size_t)stage_bytes * stages + 256 > 220 * 1024) continue;
                const int bx = (int)((pitch + sh.w_bytes - 1) / sh.w_bytes), by = (int)((rows + sh.h - 1) / sh.h);
                char nm[96]; snprintf(nm, sizeof nm, "tma box %4dB x %2d rows, %d stages, hint=%d", sh.w_bytes, sh.h, stages, hint);
                const size_t smem = (size_t)stage_bytes * stages + 2 * stages * 8;
                timeit([&] { read_tma<<<148, 64, smem>>>(tm, bx, by, sh.w_bytes / esz, sh.h, stage_bytes, stages, hint, sink); }, nm);
            }
        }
    }
    {   // read + quarter-size write, strided vs contiguous item assignment
        uint8_t* o; const int64_t opitch = 12544; const int64_t orows = rows / 2;
        CK(cudaMalloc(&o, (size_t)opitch * orows));
        CUtensorMap tm, om;
        cuuint64_t gdim[2] = {(cuuint64_t)(pitch / 8), (cuuint64_t)rows};
        cuuint64_t gstr[1] = {(cuuint64_t)pitch};
        cuuint32_t box[2] = {48, 64};
        cuuint32_t es[2] = {1, 1};
        enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t odim[2] = {(cuuint64_t)12426, (cuuint64_t)orows};
        cuuint64_t ostr[1] = {(cuuint64_t)opitch};
        cuuint32_t obox[2] = {192, 32};
        CUresult r = enc(&om, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, o, odim, ostr, obox, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) printf("out encode failed %d\n", (int)r);
        CK(cudaFuncSetAttribute(read_write_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        for (int contiguous : {0, 1})
            for (int stages : {4, 8}) {
                char nm[96]; snprintf(nm, sizeof nm, "tma read 384x64 + store 192x32, %d st, contig=%d", stages, contiguous);
                const size_t smem = (size_t)24576 * stages + 2 * stages * 8;
                timeit([&] { read_write_tma<<<148, 64, smem>>>(tm, om, 65, (int)((rows + 63) / 64), stages, contiguous, sink); }, nm);
            }
        printf("(GB/s above counts READ bytes only; add 25%% for the stores)\n");
    }
    // two CTAs per SM, smaller rings
    for (int stages : {2, 4}) {
        CUtensorMap tm;
        cuuint64_t gdim[2] = {(cuuint64_t)(pitch / 8), (cuuint64_t)rows};
        cuuint64_t gstr[1] = {(cuuint64_t)pitch};
        cuuint32_t box[2] = {48, 64};
        cuuint32_t es[2] = {1, 1};
        enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        char nm[96]; snprintf(nm, sizeof nm, "tma 384x64, %d stages, 2 CTA/SM (grid 296)", stages);
        const size_t smem = (size_t)24576 * stages + 2 * stages * 8;
        timeit([&] { read_tma<<<296, 64, smem>>>(tm, 65, (int)((rows + 63) / 64), 48, 64, 24576, stages, 1, sink); }, nm);
    }
    return 0;
}
