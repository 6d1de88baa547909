// jpeg_huffman.cu - Huffman decoding of a JPEG scan on the GPU (row N2).  A Huffman bit stream has no
// random-access points, but decoders that start at a wrong position re-synchronise with the true symbol sequence
// after a few symbols.  The scan (byte stuffing already removed) is cut into sub-sequences of kSubBits bits:
//
//   1. every thread decodes its sub-sequence from a guessed state (bit = start, first block of an MCU, DC
//      expected) and records where - and in which state - it leaves the sub-sequence;
//   2. repeat: thread i restarts from the exit state of thread i-1 whenever that differs from the state it used
//      before.  Sub-sequence 0 starts from the true state, so after pass j at least sub-sequences 0..j are right;
//      in practice the wrong starts have already synchronised and a handful of passes reach the fixed point,
//      which is the true decode by induction;
//   3. blocks completed per sub-sequence -> exclusive scan = the block each sub-sequence starts in;
//   4. decode once more from the true states, writing coefficients (DC as differences) into the dense array;
//   5. prefix-sum the DC differences per component in decode order.
//
// The decoder state is (bit position, block slot inside the MCU, next coefficient index); the step function is
// deterministic for any input, so garbage decoded from a wrong start is harmless.  Restart markers (removed by the
// host together with the byte stuffing) become hard synchronisation points: every decoder continues from an interval
// boundary in the state intervals start in, and the DC prefix sums restart there.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "jpeg_gpu.h"

namespace wicca {

namespace {

__device__ const uint8_t d_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct State { uint32_t pos; int slot, k; };

__device__ __forceinline__ uint64_t pack(const State& s) { return ((uint64_t)s.pos << 16) | ((uint64_t)s.slot << 8) | (uint64_t)s.k; }
__device__ __forceinline__ State unpack(uint64_t v) { State s; s.pos = (uint32_t)(v >> 16); s.slot = (int)((v >> 8) & 255); s.k = (int)(v & 255); return s; }

// The stream from bit `pos` on, MSB first, in a 64-bit register: at least 32 valid bits after every advance().
// The buffer behind `words` is padded with zero words.
struct BitWindow {
    const uint32_t* __restrict__ words;
    uint64_t buf;
    uint32_t next;                   // index of the next word to append
    int avail;
    __device__ __forceinline__ static uint32_t be(uint32_t w) { return __byte_perm(w, 0, 0x0123); }
    __device__ __forceinline__ void start(const uint32_t* w, uint32_t pos) {
        words = w;
        const uint32_t i = pos >> 5, sh = pos & 31;
        buf = (((uint64_t)be(w[i]) << 32) | (uint64_t)be(w[i + 1])) << sh;
        avail = 64 - (int)sh;
        next = i + 2;
    }
    __device__ __forceinline__ uint32_t top32() const { return (uint32_t)(buf >> 32); }
    __device__ __forceinline__ void advance(int n) {
        buf <<= n;
        avail -= n;
        if (avail < 32) { buf |= (uint64_t)be(words[next++]) << (32 - avail); avail += 32; }
    }
};

struct SharedTables {
    uint16_t look[8][1024];          // 0..3 DC, 4..7 AC: (length << 8) | symbol for codes of <= 10 bits
    uint8_t symbols[8][256];
    int32_t limit[8][8];             // [t][l - 10]: first 16-bit window value that is NOT a code of <= l bits (l = 10..16)
    int32_t valoffset[8][17];
    JpegGpuSlot slot[10];
    uint8_t zigzag[64];
};

// One Huffman symbol from the 32-bit window: its code length and the symbol.  Codes longer than the 10 lookahead
// bits are rare, but with 32 lanes in different places one of them meets one every other iteration, so that path is
// branch-free too: the length is 10 + the number of length classes the 16-bit window lies beyond.  Bit patterns that
// are no code at all decode as (16 bits, symbol 0): deterministic, which is all the fixed point needs.
__device__ __forceinline__ void symbol(const SharedTables& T, int table, uint32_t win, int& len, int& sym) {
    const uint32_t e = T.look[table][win >> 22];
    if (e) { len = (int)(e >> 8); sym = (int)(e & 255); return; }
    const int w16 = (int)(win >> 16);
    len = 11;
#pragma unroll
    for (int l = 1; l <= 5; ++l) len += (w16 >= T.limit[table][l]) ? 1 : 0;        // limit[1..5]: lengths 11..15
    const bool valid = w16 < T.limit[table][6];
    const int code = w16 >> (16 - len);
    sym = valid ? T.symbols[table][(code + T.valoffset[table][len]) & 255] : 0;
    len = valid ? len : 16;
}

__device__ __forceinline__ int extend(uint32_t win, int len, int s) {          // the s bits after the code, sign-extended (T.81 F.12)
    const int v = (int)((win << len) >> (32 - s));
    return v + (((v - (1 << (s - 1))) >> 31) & (1 - (1 << s)));
}

// Decode every symbol that starts in [st.pos, limit).  kWrite: also store coefficients, starting in block
// `block` (global decode order), never past block_end.  Returns the number of blocks completed.
template <bool kWrite>
__device__ __forceinline__ uint32_t decode_span(const JpegGpuScan& sc, const SharedTables& T, State& st, uint32_t limit,
                                                int64_t block, int64_t block_end) {
    uint32_t done = 0;
    int16_t* blk = nullptr;
    int mx = 0, my = 0;
    auto block_ptr = [&](int slot) -> int16_t* {
        const JpegGpuSlot& q = T.slot[slot];
        return sc.coefs + q.coef_offset + ((int64_t)(my * q.v + q.by) * q.blocks_w + (mx * q.h + q.bx)) * 64;
    };
    if (kWrite) {
        if (block >= block_end) return 0;
        const int64_t mcu = block / sc.blocks_per_mcu;
        my = (int)(mcu / sc.mcux); mx = (int)(mcu - (int64_t)my * sc.mcux);
        blk = block_ptr(st.slot);
    }
    BitWindow bw;
    bw.start(sc.words, st.pos);
    int dc_tab = T.slot[st.slot].dc_table, ac_tab = 4 + T.slot[st.slot].ac_table;
    // restart intervals: the first boundary after the start position (binary search), 0xFFFFFFFF when there is none
    uint32_t bi = 0, bound = 0xFFFFFFFFu;
    if (sc.n_bounds) {
        uint32_t lo = 0, hi = sc.n_bounds;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sc.bounds[mid] <= st.pos) lo = mid + 1; else hi = mid; }
        bi = lo; bound = sc.bounds[bi];
    }
    while (st.pos < limit) {
        if (sc.n_bounds) {
            // A correct decoder reaches an interval boundary exactly, or stops in front of it at an MCU boundary with
            // fewer than 8 one-bits of padding left.  Anything that runs across the boundary was never synchronised:
            // both cases continue from the boundary in the state every interval starts in, which also makes the
            // boundaries hard synchronisation points.
            bool jump = st.pos >= bound;
            const uint32_t gap = bound - st.pos;                       // 1..7 in the padding case
            if (!jump && st.slot == 0 && st.k == 0 && gap < 8) jump = (bw.top32() >> (32 - gap)) == (1u << gap) - 1u;
            if (jump) {
                st.pos = bound; st.slot = 0; st.k = 0;
                dc_tab = T.slot[0].dc_table; ac_tab = 4 + T.slot[0].ac_table;
                bw.start(sc.words, st.pos);
                if (kWrite) {
                    block = (int64_t)(bi + 1) * sc.blocks_per_interval;
                    if (block >= block_end) break;
                    const int64_t mcu = block / sc.blocks_per_mcu;
                    my = (int)(mcu / sc.mcux); mx = (int)(mcu - (int64_t)my * sc.mcux);
                    blk = block_ptr(0);
                }
                bound = sc.bounds[++bi];
                continue;
            }
        }
        // one symbol per iteration, DC and AC through the same instructions: lanes of a warp sit in different
        // blocks and at different coefficients, so every branch here would serialise them
        const uint32_t win = bw.top32();
        const bool dc = st.k == 0;
        int len, sym;
        symbol(T, dc ? dc_tab : ac_tab, win, len, sym);
        const int s = sym & 15;
        const int r = dc ? 0 : (sym >> 4);
        const int idx = st.k + r;                                  // the coefficient a value belongs to (0 for DC)
        if (kWrite && s && idx <= 63) blk[T.zigzag[idx]] = (int16_t)extend(win, len, s);
        st.pos += (uint32_t)(len + s);
        bw.advance(len + s);
        st.k = s ? idx + 1 : (dc ? 1 : (r == 15 ? st.k + 16 : 64));   // value / empty DC / ZRL / end of block
        if (st.k >= 64) {
            st.k = 0;
            ++done;
            if (++st.slot == sc.blocks_per_mcu) {
                st.slot = 0;
                if (kWrite && ++mx == sc.mcux) { mx = 0; ++my; }
            }
            dc_tab = T.slot[st.slot].dc_table; ac_tab = 4 + T.slot[st.slot].ac_table;
            if (kWrite) {
                if (++block >= block_end) break;
                blk = block_ptr(st.slot);
            }
        }
    }
    return done;
}

__device__ __forceinline__ void load_tables(SharedTables& T, const JpegGpuScan& sc) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(sc.tables->look);
    uint32_t* dst = reinterpret_cast<uint32_t*>(T.look);
    for (int i = threadIdx.x; i < 8 * 1024 / 2; i += blockDim.x) dst[i] = src[i];
    {
        const uint32_t* ssrc = reinterpret_cast<const uint32_t*>(sc.tables->symbols);
        uint32_t* sdst = reinterpret_cast<uint32_t*>(T.symbols);
        for (int i = threadIdx.x; i < 8 * 256 / 4; i += blockDim.x) sdst[i] = ssrc[i];
        for (int i = threadIdx.x; i < 8 * 17; i += blockDim.x) T.valoffset[i / 17][i % 17] = sc.tables->valoffset[i / 17][i % 17];
        for (int i = threadIdx.x; i < 8 * 8; i += blockDim.x) T.limit[i >> 3][i & 7] = sc.tables->limit[i >> 3][i & 7];
    }
    if (threadIdx.x < 10) T.slot[threadIdx.x] = sc.slot[threadIdx.x];
    if (threadIdx.x < 64) T.zigzag[threadIdx.x] = d_zigzag[threadIdx.x];
    __syncthreads();
}

// Pass 1 (first != 0) and the re-synchronisation passes.  In a re-synchronisation pass most thread blocks have
// nothing to do: they find that out before touching the tables.
__global__ void __launch_bounds__(128)
huff_sync_kernel(JpegGpuScan sc, const uint64_t* __restrict__ exit_in, uint64_t* __restrict__ exit_out, int first) {
    __shared__ SharedTables T;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool inside = i < sc.n_sub;
    State st;
    st.pos = i * kSubBits; st.slot = 0; st.k = 0;
    bool work = inside && first;
    if (inside && !first) {
        if (i > 0) {
            const uint64_t s = exit_in[i - 1];
            if (s != sc.start_used[i]) { st = unpack(s); work = true; }
        }
        if (!work) exit_out[i] = exit_in[i];                  // unchanged (sub-sequence 0 never changes)
    }
    if (!__syncthreads_or(work)) return;
    load_tables(T, sc);
    if (!work) return;
    if (!first) atomicAdd(sc.changed, 1);
    const uint32_t limit = min(i * kSubBits + kSubBits, sc.total_bits);
    sc.start_used[i] = pack(st);
    sc.count[i] = decode_span<false>(sc, T, st, limit, 0, 0);
    exit_out[i] = pack(st);
}

__global__ void __launch_bounds__(128)
huff_write_kernel(JpegGpuScan sc) {
    __shared__ SharedTables T;
    load_tables(T, sc);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sc.n_sub) return;
    const uint32_t limit = min(i * kSubBits + kSubBits, sc.total_bits);
    State st = unpack(sc.start_used[i]);
    decode_span<true>(sc, T, st, limit, (int64_t)sc.base[i], sc.total_blocks);
}

// ---- exclusive scan of uint32 (block counts), three small kernels -------------------------------------------
constexpr int kScanThreads = 256, kScanItems = 4, kScanChunk = kScanThreads * kScanItems;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t* s_warp, int64_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int64_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int64_t w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int64_t y = __shfl_up_sync(0xFFFFFFFFu, w, o); if (lane >= o) w += y; }
        if (lane < (int)(blockDim.x >> 5)) s_warp[lane] = w;
    }
    __syncthreads();
    total = s_warp[(blockDim.x >> 5) - 1];
    const int64_t before = warp ? s_warp[warp - 1] : 0;
    __syncthreads();
    return before + x - v;
}

// value(j): element j of the sequence being scanned.  kind 0: block counts; kind 1: DC differences of a component.
__device__ __forceinline__ int64_t dc_address(const JpegGpuScan& sc, int comp, int64_t j) {
    const JpegGpuComp& q = sc.comp[comp];
    const int per = q.h * q.v;
    const int64_t mcu = j / per;
    const int w = (int)(j - mcu * per);
    const int by = w / q.h, bx = w - by * q.h;
    const int my = (int)(mcu / sc.mcux), mx = (int)(mcu - (int64_t)my * sc.mcux);
    return q.coef_offset + ((int64_t)(my * q.v + by) * q.blocks_w + (mx * q.h + bx)) * 64;
}

template <int kKind>
__device__ __forceinline__ int64_t scan_value(const JpegGpuScan& sc, int comp, int64_t j, int64_t n) {
    if (j >= n) return 0;
    if (kKind == 0) return (int64_t)sc.count[j];
    return (int64_t)sc.coefs[dc_address(sc, comp, j)];          // kinds 1 and 2: the DC difference of block j
}

template <int kKind>
__global__ void __launch_bounds__(kScanThreads)
scan_chunk_sums_kernel(JpegGpuScan sc, int comp, int64_t n, int64_t* __restrict__ chunk_sums) {
    __shared__ int64_t s_warp[32];
    const int64_t j0 = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
    int64_t v = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) v += scan_value<kKind>(sc, comp, j0 + q, n);
    int64_t total;
    block_exclusive_scan(v, s_warp, total);
    if (threadIdx.x == 0) chunk_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
scan_of_sums_kernel(int64_t* __restrict__ chunk_sums, int n_chunks) {
    __shared__ int64_t s_warp[32];
    int64_t carry = 0;
    for (int base = 0; base < n_chunks; base += 1024) {
        const int i = base + threadIdx.x;
        const int64_t v = i < n_chunks ? chunk_sums[i] : 0;
        int64_t total;
        const int64_t ex = block_exclusive_scan(v, s_warp, total);
        if (i < n_chunks) chunk_sums[i] = carry + ex;
        carry += total;
    }
}

template <int kKind>
__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(JpegGpuScan sc, int comp, int64_t n, const int64_t* __restrict__ chunk_sums) {
    __shared__ int64_t s_warp[32];
    const int64_t j0 = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
    int64_t item[kScanItems], v = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) { item[q] = scan_value<kKind>(sc, comp, j0 + q, n); v += item[q]; }
    int64_t total;
    int64_t run = chunk_sums[blockIdx.x] + block_exclusive_scan(v, s_warp, total);
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
        if (j0 + q < n) {
            if (kKind == 0) sc.base[j0 + q] = (uint32_t)run;                                   // exclusive: blocks before
            else if (kKind == 1) sc.coefs[dc_address(sc, comp, j0 + q)] = (int16_t)(run + item[q]);   // inclusive: the DC value
            else sc.dc_prefix[j0 + q] = (int32_t)(run + item[q]);                              // running sum over the whole scan
        }
        run += item[q];
    }
}

// Restart intervals reset the DC predictor: DC(j) = prefix(j) - prefix(last block of the previous interval).
__global__ void __launch_bounds__(256)
dc_segment_kernel(JpegGpuScan sc, int comp, int64_t n, int64_t seg_blocks) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int64_t start = (j / seg_blocks) * seg_blocks;
    const int32_t before = start ? sc.dc_prefix[start - 1] : 0;
    sc.coefs[dc_address(sc, comp, j)] = (int16_t)(sc.dc_prefix[j] - before);
}

template <int kKind>
cudaError_t run_scan(const JpegGpuScan& sc, int comp, int64_t n, int64_t* d_chunk_sums, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int n_chunks = (int)((n + kScanChunk - 1) / kScanChunk);
    scan_chunk_sums_kernel<kKind><<<n_chunks, kScanThreads, 0, stream>>>(sc, comp, n, d_chunk_sums);
    scan_of_sums_kernel<<<1, 1024, 0, stream>>>(d_chunk_sums, n_chunks);
    scan_apply_kernel<kKind><<<n_chunks, kScanThreads, 0, stream>>>(sc, comp, n, d_chunk_sums);
    return cudaGetLastError();
}

}  // namespace

size_t jpeg_gpu_chunk_sum_capacity(const JpegGpuScan& sc) {
    int64_t n = (int64_t)sc.n_sub;
    for (int c = 0; c < sc.ncomp; ++c) n = n > sc.comp[c].n_blocks ? n : sc.comp[c].n_blocks;
    return (size_t)((n + kScanChunk - 1) / kScanChunk + 1) * sizeof(int64_t);
}

// h_changed: page-locked int the fixed-point loop polls.  Returns cudaErrorNotReady when max_passes were not enough.
cudaError_t launch_jpeg_huffman(const JpegGpuScan& sc, uint64_t* d_exit_a, uint64_t* d_exit_b, int64_t* d_chunk_sums,
                                int* h_changed, int max_passes, int* passes_out, cudaStream_t stream) {
    const int threads = 128;
    const int grid = (int)((sc.n_sub + threads - 1) / threads);
    cudaError_t e = cudaMemsetAsync(sc.coefs, 0, (size_t)sc.total_coefs * sizeof(int16_t), stream);
    if (e != cudaSuccess) return e;
    huff_sync_kernel<<<grid, threads, 0, stream>>>(sc, d_exit_b, d_exit_a, 1);
    uint64_t* cur = d_exit_a;
    uint64_t* nxt = d_exit_b;
    int passes = 0;
    bool converged = sc.n_sub <= 1;
    while (!converged && passes < max_passes) {
        // a few passes per round trip: the flag is cleared, the passes run, the flag comes back
        e = cudaMemsetAsync(sc.changed, 0, sizeof(int), stream);
        if (e != cudaSuccess) return e;
        // typical files settle in under ten passes; files almost without end-of-block codes (quality 100) need a
        // couple of hundred, so the bursts between two looks at the flag grow
        const int burst = passes == 0 ? 2 : (passes < 8 ? 4 : (passes < 32 ? 8 : 32));
        for (int b = 0; b < burst; ++b, ++passes) {
            // only the last pass of a burst decides: an earlier change is fine as long as the last one finds none
            if (b == burst - 1) { e = cudaMemsetAsync(sc.changed, 0, sizeof(int), stream); if (e != cudaSuccess) return e; }
            huff_sync_kernel<<<grid, threads, 0, stream>>>(sc, cur, nxt, 0);
            uint64_t* t = cur; cur = nxt; nxt = t;
        }
        e = cudaMemcpyAsync(h_changed, sc.changed, sizeof(int), cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
        converged = (*h_changed == 0);
        if (getenv("WICCA_JPEG_DEBUG")) fprintf(stderr, "[jpeg huffman] n_sub %u pass %d: %d sub-sequences re-decoded in the last pass\n", sc.n_sub, passes, *h_changed);
    }
    if (passes_out) *passes_out = passes;
    if (!converged) return cudaErrorNotReady;
    e = run_scan<0>(sc, 0, (int64_t)sc.n_sub, d_chunk_sums, stream);
    if (e != cudaSuccess) return e;
    huff_write_kernel<<<grid, threads, 0, stream>>>(sc);
    for (int c = 0; c < sc.ncomp; ++c) {
        const int64_t n = sc.comp[c].n_blocks;
        if (sc.n_bounds == 0) {
            e = run_scan<1>(sc, c, n, d_chunk_sums, stream);
        } else {
            e = run_scan<2>(sc, c, n, d_chunk_sums, stream);
            if (e != cudaSuccess) return e;
            const int64_t seg_blocks = sc.blocks_per_interval / sc.blocks_per_mcu * sc.comp[c].h * sc.comp[c].v;
            dc_segment_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(sc, c, n, seg_blocks);
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

}  // namespace wicca
