// haar_icon.cu - Haar LL "icon" kernels for sm_100a.
//
// Replaces the body of HaarCoder.get_small_copy (wicca/wavelet_coder.py:56-67) including the
// padding of get_padded_copy (wicca/data_loader.py:107-117), which is never materialised.
//
//   haar_icon_tma2_kernel  one HBM pass over a pitched RGB image produces the icons of ANY
//                          subset of depths 1..6.  Persistent CTAs; a producer lane streams
//                          128 px x 64 row tiles (24 KB) into a shared-memory ring with TMA
//                          (cp.async.bulk.tensor + mbarrier); two consumer warps per tile reduce
//                          it: levels 1-3 in registers (packed 16-bit lanes) and out through
//                          TMA store, levels 4-5 with warp shuffles, level 6 through a mailbox.
//                          HBM-bound: 3 B/px read, <= 1 B/px written.
//   edge_strip_kernel      tiny pre-pass, only for BORDER_REFLECT / REFLECT_101 / WRAP: materialises
//                          the border-extended right strip (<= 78 px per row).  REPLICATE and
//                          CONSTANT borders are patched from the shared-memory tile instead.
//   haar_icon_generic_kernel  any C, any alignment, depth <= 8: one thread per output element.
//   haar_level_f32_kernel  one further level in fp32, for depths > 8 (the reference's fp32
//                          arithmetic stops being exact there, so it is replayed literally).
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>

#include "haar_math.cuh"
#include "icon_types.h"
#include "kernels.h"

namespace wicca {

// ------------------------------------------------------------------------------------------
// PTX helpers (mbarrier + TMA)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// Tensor maps are read through the tensormap proxy; they were written by the host (cudaMemcpy),
// so an acquire fence at system scope is required before the first use from global memory.
__device__ __forceinline__ void fence_tensormap_acquire(const CUtensorMap* tmap) {
    asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar,
                                            uint64_t cache_policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(cache_policy)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* smem_src, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(x),
                 "r"(y), "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_nohint(void* smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// Work item of a CTA's k-th turn.  run_len consecutive turns take horizontally adjacent items, so
// the narrow output rows of neighbouring items reach L2 together and merge into full lines.
__device__ __forceinline__ int item_index(int k, int run_len) {
    const int u = k / run_len;
    return (u * (int)gridDim.x + (int)blockIdx.x) * run_len + (k - u * run_len);
}

// ------------------------------------------------------------------------------------------
// The one-pass kernel.  dev_mode exists only in -DWICCA_DEV builds (tools/bench_variants.py: 1 = consumers skip the
// arithmetic, 2 = arithmetic but no level 1..3 output); the release library compiles it to the constant 0, so no
// environment variable or argument can make the product kernel skip work.
// Two consumer warps per stage: warp pair p owns stage p; the upper warp reduces rows 0..31 of the
// 128 x 64 item, the lower warp rows 32..63 (8 rows per lane), so a stage is held half as long and
// twice as many warps hide each other's latencies.  Levels 1..5 are complete inside a warp; the
// two 64 x 32 partial sums of level 6 meet through a small shared-memory mailbox.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <int kStages>
__global__ void __launch_bounds__(32 * (1 + 2 * kStages), 1)
haar_icon_tma2_kernel(const IconImage* __restrict__ imgs, const uint8_t* const* __restrict__ strips, int n_images,
                      int total_items, int border_type, int border_const, int dev_mode, int run_len, int stream_hint) {
    static_assert(kStages <= 15, "one named barrier per warp pair");
#ifdef WICCA_DEV
    const int debug = dev_mode;
#else
    constexpr int debug = 0;
    (void)dev_mode;
#endif
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* stages = smem_raw;
    uint8_t* out_tiles = smem_raw + (size_t)kStages * kStageBytes;
    uint32_t* mailbox = reinterpret_cast<uint32_t*>(out_tiles + (size_t)2 * kStages * kHalfStageBytes);   // [pair][2][8]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(mailbox + kStages * 16);
    uint64_t* empty_bar = full_bar + kStages;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 2);       // both warps of the pair release the stage
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            int img = -1, base = 0, next_base = 0, items_x = 1;
            const CUtensorMap* tmap = nullptr;
            int k = 0;
            for (;; ++k) {
                const int g = item_index(k, run_len);
                if (g >= total_items) break;
                if (g >= next_base) {
                    do {
                        ++img;
                        next_base = (img + 1 < n_images) ? imgs[img + 1].item_base : 0x7FFFFFFF;
                    } while (g >= next_base);
                    base = imgs[img].item_base;
                    items_x = imgs[img].items_x;
                    tmap = &imgs[img].tmap;
                    fence_tensormap_acquire(tmap);
                }
                const int local = g - base;
                const int iy = local / items_x;
                const int ix = local - iy * items_x;
                const int s = k % kStages;
                const uint32_t ph = (uint32_t)(k / kStages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
                // L2::evict_first on the input only pays off when (almost) nothing is written: with a
                // sizeable output stream it keeps dirty lines in L2 too long and costs ~13 % (measured).
                if (stream_hint) tma_load_2d(stages + (size_t)s * kStageBytes, tmap, ix * (kStageRowBytes / 4), iy * kItemH,
                                             &full_bar[s], pol);
                else tma_load_2d_nohint(stages + (size_t)s * kStageBytes, tmap, ix * (kStageRowBytes / 4), iy * kItemH,
                                        &full_bar[s]);
            }
        }
        return;
    }

    const int cw = warp - 1;
    const int pair = cw >> 1;          // == stage index
    const int half = cw & 1;           // 0: rows 0..31 of the item, 1: rows 32..63
    const int cx = lane & 7;
    const int ry = lane >> 3;          // 8-row group inside the half
    const uint32_t fill = (uint32_t)border_const * 0x01010101u;
    uint8_t* tile = out_tiles + (size_t)cw * kHalfStageBytes;
    uint32_t* box = mailbox + pair * 16;
    int img = -1, base = 0, next_base = 0;
    ImageGeom geo;
    IconSink sk;
    const uint8_t* strip = nullptr;
    const CUtensorMap* hmap = nullptr;
    unsigned mask = 0;
    uint32_t sink_word = 0;

    for (int k = pair;; k += kStages) {
        const int g = item_index(k, run_len);
        if (g >= total_items) break;
        const uint32_t ph = (uint32_t)(k / kStages) & 1u;
        if (g >= next_base) {
            do {
                ++img;
                next_base = (img + 1 < n_images) ? imgs[img + 1].item_base : 0x7FFFFFFF;
            } while (g >= next_base);
            const IconImage& im = imgs[img];
            base = im.item_base;
            geo = make_geom(im);
            sk = make_sink(im);
            strip = strips[img];
            hmap = im.hmap;
            mask = 0;
#pragma unroll
            for (int l = 0; l < 3; ++l) mask |= (sk.icon[l] != nullptr ? 1u : 0u) << l;
            if (debug == 2) {
                mask = 0;
#pragma unroll
                for (int l = 0; l < 3; ++l) sk.icon[l] = nullptr;
            }
            if (lane == 0) {
#pragma unroll
                for (int l = 0; l < 3; ++l)
                    if ((mask >> l) & 1u) fence_tensormap_acquire(&hmap[l]);
            }
        }
        const int local = g - base;
        const int iy = local / geo.items_x;
        const int ix = local - iy * geo.items_x;
        const ChunkSrc cs = make_chunk_src(geo, strip, stages + (size_t)pair * kStageBytes, ix, iy, cx,
                                           half * 32 + ry * 8, 8, border_type, fill);

        if (lane == 0) bulk_wait_read0();     // previous TMA stores have finished reading the tile
        __syncwarp();

        mbar_wait(&full_bar[pair], ph);

        uint32_t acc[3];
        if (debug == 1) {
            sink_word ^= *reinterpret_cast<const uint32_t*>(cs.smem);
            acc[0] = acc[1] = acc[2] = 0u;
        } else {
            const StagedEmit em = staged_emit_half(tile, cx, ry, mask);
            reduce_lane(cs, em, acc);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[pair]);

        if (mask != 0u && debug != 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (mask & 1u) tma_store_2d(&hmap[0], tile + kHalf1Off, ix * kOut1Row, iy * 32 + half * 16);
                if (mask & 2u) tma_store_2d(&hmap[1], tile + kHalf2Off, ix * kOut2Row, iy * 16 + half * 8);
                if (mask & 4u) tma_store_2d(&hmap[2], tile + kHalf3Off, ix * kOut3Row, iy * 8 + half * 4);
                bulk_commit();
            }
        }

        // level 4: two 8-row groups; level 5: two chunk columns x two 16-row groups; level 6 partial:
        // two 32-px blocks (this warp's 32 rows), completed through the pair's mailbox
        uint32_t s4[3], s5[3], s6[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t a = acc[c] + __shfl_xor_sync(0xFFFFFFFFu, acc[c], 8);
            s4[c] = a;
            uint32_t b = a + __shfl_xor_sync(0xFFFFFFFFu, a, 1);
            b += __shfl_xor_sync(0xFFFFFFFFu, b, 16);
            s5[c] = b;
            s6[c] = b + __shfl_xor_sync(0xFFFFFFFFu, b, 2);
        }
        if ((sk.icon[5] != nullptr || sk.sum6 != nullptr) && debug != 1) {
            uint32_t* slot = box + (k / kStages & 1) * 8;        // double-buffered per item parity
            if (half == 1 && (cx & 3) == 0 && ry == 0) {
                uint32_t* q = slot + (cx >> 2) * 3;
                q[0] = s6[0]; q[1] = s6[1]; q[2] = s6[2];
            }
            pair_barrier(1 + pair);
            if (half == 0 && (cx & 3) == 0 && ry == 0) {
                const uint32_t* q = slot + (cx >> 2) * 3;
                s6[0] += q[0]; s6[1] += q[1]; s6[2] += q[2];
            }
        }
        if (debug != 1) emit_tail_half(sk, cs.x0, cs.y0, cx, ry, half == 0, s4, s5, s6);
    }
    if (lane == 0) bulk_wait0();
    if (debug == 1 && sink_word == 0x9E3779B9u && sk.icon[5] != nullptr) sk.icon[5][0] = (uint8_t)sink_word;
}

// ------------------------------------------------------------------------------------------
// Right-edge strip: strip[y][ (x - Wa)*3 + c ] = padded_image[y][x][c] for x in [Wa, Wa + strip_px)
// ------------------------------------------------------------------------------------------
__global__ void edge_strip_kernel(const IconImage* __restrict__ imgs, uint8_t* const* __restrict__ strips,
                                  int border_type, int border_const) {
    const IconImage& im = imgs[blockIdx.y];
    uint8_t* strip = strips[blockIdx.y];
    if (strip == nullptr) return;
    const int npx = im.Wp_max - (im.W & ~(kChunkPx - 1));   // <= 78
    if (npx <= 0) return;
    const int words = (npx * 3 + 3) >> 2;                   // <= 59
    const int total = im.H * words;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / words;
        strip_word(im, strip, y, i - y * words, border_type, border_const);
    }
}

// ------------------------------------------------------------------------------------------
// General kernel: thread per output element, exact integer block sum (depth <= 8)
// ------------------------------------------------------------------------------------------
__global__ void haar_icon_generic_kernel(GenericIconArgs a) {
    const int r = 1 << a.depth;
    const int64_t n = (int64_t)a.out_h * a.out_w * a.C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % a.C);
        const int64_t t = i / a.C;
        const int ox = (int)(t % a.out_w);
        const int oy = (int)(t / a.out_w);
        uint32_t sum = 0;
        for (int dy = 0; dy < r; ++dy) {
            const int ym = border_index(oy * r + dy, a.H, a.border_type);
            if (ym < 0) { sum += (uint32_t)a.border_const * (uint32_t)r; continue; }
            const uint8_t* row = a.src + (int64_t)ym * a.pitch + c;
            const int xb = ox * r;
            if (xb + r <= a.W) {
                for (int dx = 0; dx < r; ++dx) sum += row[(int64_t)(xb + dx) * a.C];
            } else {
                for (int dx = 0; dx < r; ++dx) {
                    const int xm = border_index(xb + dx, a.W, a.border_type);
                    sum += (xm < 0) ? (uint32_t)a.border_const : (uint32_t)row[(int64_t)xm * a.C];
                }
            }
        }
        if (a.dst_u8 != nullptr) {
            a.dst_u8[(int64_t)oy * a.dst_pitch + (int64_t)ox * a.C + c] = (uint8_t)(sum >> (2 * a.depth));
        } else {
            // exact: sum < 2^24 for depth <= 8, and the scale is a power of two
            a.dst_f32[i] = __uint2float_rn(sum) * (1.0f / (float)(1u << (2 * a.depth)));
        }
    }
}

// One more level in fp32, literally wavelet_coder.py:62-65: (even_row + odd_row), then
// (even_col + odd_col) * 0.25.  in: (2*out_h, 2*out_w, C) tight fp32.
__global__ void haar_level_f32_kernel(const float* __restrict__ in, float* __restrict__ out, uint8_t* out_u8,
                                      int out_h, int out_w, int C) {
    const int64_t n = (int64_t)out_h * out_w * C;
    const int64_t in_row = (int64_t)2 * out_w * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t t = i / C;
        const int ox = (int)(t % out_w);
        const int oy = (int)(t / out_w);
        const float* p = in + (int64_t)(2 * oy) * in_row + (int64_t)(2 * ox) * C + c;
        const float s0 = __fadd_rn(p[0], p[in_row]);          // sums[:, even]
        const float s1 = __fadd_rn(p[C], p[in_row + C]);      // sums[:, odd]
        const float v = __fmul_rn(__fadd_rn(s0, s1), 0.25f);
        if (out_u8 != nullptr) {
            const float cl = fminf(fmaxf(v, 0.0f), 255.0f);   // np.clip, then astype(uint8) truncates
            out_u8[i] = (uint8_t)cl;
        } else {
            out[i] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Launchers
// ------------------------------------------------------------------------------------------
template <int S>
static cudaError_t launch_tma2_variant(const IconImage* d_imgs, const uint8_t* const* d_strips, int n_images,
                                       int total_items, int border_type, int border_const, int grid, int debug,
                                       int run_len, int stream_hint, cudaStream_t stream) {
    const size_t smem = (size_t)S * kStageBytes + (size_t)2 * S * kHalfStageBytes + (size_t)S * 16 * sizeof(uint32_t) +
                        2 * S * sizeof(uint64_t);
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(haar_icon_tma2_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured_dev = dev;
    }
    haar_icon_tma2_kernel<S><<<grid, 32 * (1 + 2 * S), smem, stream>>>(d_imgs, d_strips, n_images, total_items, border_type,
                                                                      border_const, debug, run_len, stream_hint);
    return cudaGetLastError();
}

// variant 0 (default): 6 stages, runs of up to 16 horizontally adjacent items per CTA (shorter when
// the launch is small, to keep the CTAs balanced).  Developer knobs (WICCA_ICON_VARIANT, -DWICCA_DEV builds only):
// variant % 100 in {20,21,22,23} = 7/6/5/4 stages, + 100 * debug mode, + 1000 * (1 + log2(run length)).
cudaError_t launch_icon_tma(const IconImage* d_imgs, const uint8_t* const* d_strips, int n_images, int total_items,
                            int border_type, int border_const, int sm_count, int variant, int stream_hint,
                            cudaStream_t stream) {
    if (total_items <= 0) return cudaSuccess;
    const int grid = total_items < sm_count ? total_items : sm_count;
    int run_len;
    const int run_code = (variant / 1000) % 10;
    if (run_code > 0) {
        run_len = 1 << (run_code - 1);
    } else {
        run_len = 16;
        while (run_len > 1 && (int64_t)8 * run_len * grid > total_items) run_len >>= 1;
    }
    if ((int64_t)run_len * grid > total_items) run_len = 1;
#ifdef WICCA_DEV
    const int debug = (variant / 100) % 10;
    switch (variant % 100) {
        case 20: return launch_tma2_variant<7>(d_imgs, d_strips, n_images, total_items, border_type, border_const, grid, debug, run_len, stream_hint, stream);
        case 22: return launch_tma2_variant<5>(d_imgs, d_strips, n_images, total_items, border_type, border_const, grid, debug, run_len, stream_hint, stream);
        case 23: return launch_tma2_variant<4>(d_imgs, d_strips, n_images, total_items, border_type, border_const, grid, debug, run_len, stream_hint, stream);
        default: return launch_tma2_variant<6>(d_imgs, d_strips, n_images, total_items, border_type, border_const, grid, debug, run_len, stream_hint, stream);
    }
#else
    return launch_tma2_variant<6>(d_imgs, d_strips, n_images, total_items, border_type, border_const, grid, 0, run_len, stream_hint, stream);
#endif
}

cudaError_t launch_edge_strips(const IconImage* d_imgs, uint8_t* const* d_strips, int n_images, int max_rows,
                               int border_type, int border_const, cudaStream_t stream) {
    if (n_images <= 0) return cudaSuccess;
    int bx = (max_rows * 60 + 255) / 256;
    if (bx < 1) bx = 1;
    if (bx > 1024) bx = 1024;
    edge_strip_kernel<<<dim3(bx, n_images), 256, 0, stream>>>(d_imgs, d_strips, border_type, border_const);
    return cudaGetLastError();
}

cudaError_t launch_icon_generic(const GenericIconArgs& a, cudaStream_t stream) {
    const int64_t n = (int64_t)a.out_h * a.out_w * a.C;
    if (n <= 0) return cudaSuccess;
    // coalesced row-streaming kernel (haar_rows.cu) whenever one output pixel's row segment fits a tile;
    // the scalar kernel below remains for absurd channel counts only
    if (rows_kernel_groups(a.C, a.depth) > 0 && (((int64_t)a.out_w << a.depth) + 8) * a.C < 0x7FFFFFFF) return launch_icon_rows(a, stream);
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    haar_icon_generic_kernel<<<(int)blocks, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_level_f32(const float* in, float* out, uint8_t* out_u8, int out_h, int out_w, int C,
                             cudaStream_t stream) {
    const int64_t n = (int64_t)out_h * out_w * C;
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    haar_level_f32_kernel<<<(int)blocks, 256, 0, stream>>>(in, out, out_u8, out_h, out_w, C);
    return cudaGetLastError();
}

}  // namespace wicca
