// jpeg_gpu.h - descriptors shared by the host code and the GPU Huffman decoder (jpeg_huffman.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace wicca {

constexpr uint32_t kSubBits = 2048;          // bits per sub-sequence (one thread each); 1024 decodes one file 10 % faster, 2048 a batch 11 % faster

struct JpegGpuTables {                        // 0..3 DC tables, 4..7 AC tables (by Huffman table id)
    uint16_t look[8][1024];
    int32_t limit[8][8];                      // [t][l - 10], l = 10..16: first 16-bit window value beyond the codes of <= l bits
    int32_t valoffset[8][17];
    uint8_t symbols[8][256];
};
struct JpegGpuSlot {                          // one block position inside an MCU
    int64_t coef_offset;                      // of its component
    int h, v, bx, by, blocks_w;
    int dc_table, ac_table;
};
struct JpegGpuComp {
    int64_t coef_offset, n_blocks;
    int h, v, blocks_w;
};
struct JpegGpuScan {
    const uint32_t* words;                    // the scan, byte stuffing removed, zero padded
    uint32_t total_bits, n_sub;
    int blocks_per_mcu, mcux, ncomp;
    int64_t total_blocks, total_coefs;
    JpegGpuSlot slot[10];
    JpegGpuComp comp[3];
    int16_t* coefs;                           // dense quantised coefficients (output)
    uint64_t* start_used;                     // [n_sub] state each sub-sequence was last decoded from
    uint32_t* count;                          // [n_sub] blocks completed
    uint32_t* base;                           // [n_sub] blocks before the sub-sequence
    int* changed;
    const JpegGpuTables* tables;              // device copy
    // restart intervals: bit offsets at which intervals 1, 2, ... start, ascending, closed by 0xFFFFFFFF
    const uint32_t* bounds;
    uint32_t n_bounds;                        // without the sentinel; 0 for files without restart markers
    int64_t blocks_per_interval;              // restart_interval * blocks_per_mcu
    int32_t* dc_prefix;                       // [max component blocks] running sums for the segmented DC scan
};

size_t jpeg_gpu_chunk_sum_capacity(const JpegGpuScan& sc);
cudaError_t launch_jpeg_huffman(const JpegGpuScan& sc, uint64_t* d_exit_a, uint64_t* d_exit_b, int64_t* d_chunk_sums,
                                int* h_changed, int max_passes, int* passes_out, cudaStream_t stream);

}  // namespace wicca
