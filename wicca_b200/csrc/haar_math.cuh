// haar_math.cuh - register-level arithmetic of the one-pass Haar icon kernel.
//
// Everything here is __host__ __device__ so the exact code the GPU runs can be
// replayed lane by lane on the CPU (tests/cpu_emul/emul_icon.cpp) and compared
// with the oracle without a GPU.
//
// Model.  One thread owns a "chunk": 16 RGB pixels = 48 bytes = 12 words of one
// image row, and walks 8 rows of it two at a time.  The reference computes, per
// level, (even_row + odd_row) then (even_col + odd_col) * 0.25 in fp32
// (wicca/wavelet_coder.py:61-65) and truncates once at the end (:67).  Those
// fp32 values are exact, so level-d output = (sum of the 2^d x 2^d block) >> 2d.
// We therefore carry exact integer block sums up the pyramid:
//   s1 (2x2 sums, <= 1020), s2 (4x4, <= 4080), s3 (8x8, <= 16320) live in packed
//   16-bit lanes (two values per register); s4..s6 in 32-bit registers.
#pragma once
#include <stdint.h>
#include <string.h>

#include "icon_types.h"

#if defined(__CUDACC__)
#define WHD __host__ __device__ __forceinline__
#else
#define WHD inline
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
#endif

namespace wicca {

// PRMT: result byte i = byte sel.nibble[i] of the 8-byte pool {a (0..3), b (4..7)}.
WHD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    uint64_t pool = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        uint32_t n = (sel >> (4 * i)) & 0x7;
        r |= (uint32_t)((pool >> (8 * n)) & 0xFF) << (8 * i);
    }
    return r;
#endif
}

// (x & m) | (y & ~m): one LOP3 on the device.
WHD uint32_t bitsel(uint32_t m, uint32_t x, uint32_t y) { return (x & m) | (y & ~m); }

// ---------------------------------------------------------------------------
// Level 1 of one row pair.
// In : A[12], B[12] - the 48 bytes of rows 2i and 2i+1 of the chunk.
// Out: T[c*4 + j] = ( s1(j, c) , s1(j+4, c) ) packed lo/hi, for c in 0..2, j in 0..3,
//      where s1(q, c) = sum over the 2 rows of pixel 2q and 2q+1, channel c.
// ---------------------------------------------------------------------------
WHD void level1_rowpair(const uint32_t (&A)[12], const uint32_t (&B)[12], uint32_t (&T)[12]) {
    // vertical add, widening bytes to 16-bit lanes:
    //   E[k] = (byte0, byte2) of word k, O[k] = (byte1, byte3) of word k
    uint32_t E[12], O[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        O[k] = prmt(A[k], 0u, 0x4341u) + prmt(B[k], 0u, 0x4341u);
        // A = E_A + (O_A << 8) exactly, so the even-byte sums fall out of one wrap-around
        // subtract (an IMAD on the FMA pipe) instead of two maskings on the ALU pipe.
        E[k] = A[k] + B[k] - (O[k] << 8);
    }
    // G[b] = ( V(b) , V(b+24) ) for byte position b in 0..23 of the chunk row, V = vertical sum
    uint32_t G[24];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        G[4 * k + 0] = prmt(E[k], E[k + 6], 0x5410u);
        G[4 * k + 2] = prmt(E[k], E[k + 6], 0x7632u);
        G[4 * k + 1] = prmt(O[k], O[k + 6], 0x5410u);
        G[4 * k + 3] = prmt(O[k], O[k + 6], 0x7632u);
    }
    // horizontal pair: byte b of pixel 2q pairs with byte b+3 of pixel 2q+1
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) T[c * 4 + j] = G[6 * j + c] + G[6 * j + c + 3];
}

// pack two lane-registers into 4 bytes: ( x.lo>>S, y.lo>>S, x.hi>>S, y.hi>>S ), values < 2^(8+S)
template <int S>
WHD uint32_t pack2(uint32_t x, uint32_t y) {
    return bitsel(0x00FF00FFu, x >> S, y << (8 - S));
}

// Level-1 icon bytes of one row pair: 8 pixels x 3 channels = 24 bytes = W[6].
WHD void icon1_words(const uint32_t (&T)[12], uint32_t (&W)[6]) {
    // output byte 3q+c = s1(q,c) >> 2 ; T index = c*4 + (q & 3), lane = q >> 2
    // word n (q in 0..3) takes entries L[4n..4n+3] of (q,c) = (0,0)(0,1)(0,2)(1,0)(1,1)(1,2)(2,0)...
#pragma unroll
    for (int n = 0; n < 3; ++n) {
        const int b0 = 4 * n, b1 = 4 * n + 1, b2 = 4 * n + 2, b3 = 4 * n + 3;
        uint32_t P1 = pack2<2>(T[(b0 % 3) * 4 + b0 / 3], T[(b1 % 3) * 4 + b1 / 3]);
        uint32_t P2 = pack2<2>(T[(b2 % 3) * 4 + b2 / 3], T[(b3 % 3) * 4 + b3 / 3]);
        W[n]     = prmt(P1, P2, 0x5410u);   // lo lanes: pixels 0..3
        W[n + 3] = prmt(P1, P2, 0x7632u);   // hi lanes: pixels 4..7
    }
}

// Level-2 partial of one row pair: U[c*2+h] = T[c][2h] + T[c][2h+1]
//   U[c*2+0] = ( s2row(0,c), s2row(2,c) ), U[c*2+1] = ( s2row(1,c), s2row(3,c) )
WHD void level2_accumulate(const uint32_t (&T)[12], uint32_t (&acc2)[6]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        acc2[c * 2 + 0] += T[c * 4 + 0] + T[c * 4 + 1];
        acc2[c * 2 + 1] += T[c * 4 + 2] + T[c * 4 + 3];
    }
}

// Level-2 icon bytes (4 pixels x 3 channels = 12 bytes) from complete 4x4 sums.
WHD void icon2_words(const uint32_t (&acc2)[6], uint32_t (&W)[3]) {
    uint32_t PA = pack2<4>(acc2[0], acc2[2]);   // s2(0,0) s2(0,1) | s2(2,0) s2(2,1)
    uint32_t PB = pack2<4>(acc2[4], acc2[1]);   // s2(0,2) s2(1,0) | s2(2,2) s2(3,0)
    uint32_t PC = pack2<4>(acc2[3], acc2[5]);   // s2(1,1) s2(1,2) | s2(3,1) s2(3,2)
    W[0] = prmt(PA, PB, 0x5410u);
    W[1] = prmt(PC, PA, 0x7610u);
    W[2] = prmt(PB, PC, 0x7632u);
}

// Level 3: acc3[c] += ( s3row(0,c), s3row(1,c) ) from complete level-2 sums.
WHD void level3_accumulate(const uint32_t (&acc2)[6], uint32_t (&acc3)[3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) acc3[c] += acc2[c * 2 + 0] + acc2[c * 2 + 1];
}

// Level-3 icon bytes: 2 pixels x 3 channels = 6 bytes, returned as three 16-bit values.
WHD void icon3_halves(const uint32_t (&acc3)[3], uint32_t (&Hh)[3]) {
    const uint32_t p0c0 = (acc3[0] & 0xFFFFu) >> 6, p1c0 = acc3[0] >> 22;
    const uint32_t p0c1 = (acc3[1] & 0xFFFFu) >> 6, p1c1 = acc3[1] >> 22;
    const uint32_t p0c2 = (acc3[2] & 0xFFFFu) >> 6, p1c2 = acc3[2] >> 22;
    Hh[0] = p0c0 | (p0c1 << 8);
    Hh[1] = p0c2 | (p1c0 << 8);
    Hh[2] = p1c1 | (p1c2 << 8);
}

// Level 4: acc4[c] += lo + hi of the complete level-3 sums (32-bit from here on).
WHD void level4_accumulate(const uint32_t (&acc3)[3], uint32_t (&acc4)[3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) acc4[c] += (acc3[c] & 0xFFFFu) + (acc3[c] >> 16);
}

// ---------------------------------------------------------------------------
// Output side.  IconSink = where the (up to six) icons of the current image live.
// ---------------------------------------------------------------------------
struct IconSink {
    uint8_t* icon[6];      // index = depth-1; nullptr = level not requested
    int64_t pitch[6];
    int h[6], w[6];        // icon extents: ceil(H / 2^d), ceil(W / 2^d)
    uint32_t* sum6;        // nullptr, or the plane of exact level-6 block sums (depths > 6), sum6_stride blocks per row
    int sum6_h, sum6_w, sum6_stride;
};

// ---------------------------------------------------------------------------
// Emitter for levels 1..3: dense per-warp tiles in shared memory (a warp owns half an item,
// 128 x 32 px: L1 16 x 192 B, L2 8 x 96 B, L3 4 x 48 B) that one lane then hands to TMA store,
// which clips at the icon edge in hardware.  r = chunk row (0..7) at which the group starts.
// ---------------------------------------------------------------------------
struct StagedEmit {
    uint8_t* t1; uint8_t* t2; uint8_t* t3;   // this lane's first row inside the warp's L1 / L2 / L3 tile, at its column
    unsigned mask;                           // bit l set = level l+1 requested
    WHD bool want(int l) const { return (mask >> l) & 1u; }
    WHD void icon1(int r, const uint32_t (&W)[6]) const {
        uint2* q = reinterpret_cast<uint2*>(t1 + (r >> 1) * kOut1Row);
        uint2 v0, v1, v2;
        v0.x = W[0]; v0.y = W[1]; v1.x = W[2]; v1.y = W[3]; v2.x = W[4]; v2.y = W[5];
        q[0] = v0; q[1] = v1; q[2] = v2;
    }
    WHD void icon2(int r, const uint32_t (&W)[3]) const {
        uint32_t* q = reinterpret_cast<uint32_t*>(t2 + (r >> 2) * kOut2Row);
        q[0] = W[0]; q[1] = W[1]; q[2] = W[2];
    }
    WHD void icon3(int r, const uint32_t (&Hh)[3]) const {
        uint16_t* q = reinterpret_cast<uint16_t*>(t3 + (r >> 3) * kOut3Row);
        q[0] = (uint16_t)Hh[0]; q[1] = (uint16_t)Hh[1]; q[2] = (uint16_t)Hh[2];
    }
};
// Lane (cx, ry) of a warp that owns half an item, 128 x 32 (8 rows per lane, tiles 16/8/4 rows).
WHD StagedEmit staged_emit_half(uint8_t* tile, int cx, int ry, unsigned mask) {
    StagedEmit e;
    e.t1 = tile + kHalf1Off + (ry * 4) * kOut1Row + cx * 24;
    e.t2 = tile + kHalf2Off + (ry * 2) * kOut2Row + cx * 12;
    e.t3 = tile + kHalf3Off + ry * kOut3Row + cx * 6;
    e.mask = mask;
    return e;
}

// Reduce one chunk (16 px x 8 rows): emits levels 1..3 through `em` and returns the three
// 16 x 8 channel sums in acc.  `load(r, A)` fills A[12] with the 48 bytes of chunk row r
// (0..7), border-extended.
template <class Loader>
WHD void reduce_chunk(Loader& load, const StagedEmit& em, uint32_t (&acc)[3]) {
    const bool want1 = em.want(0), want2 = em.want(1), want3 = em.want(2);
    uint32_t acc3[3] = {0u, 0u, 0u};
#pragma unroll
    for (int q4 = 0; q4 < 2; ++q4) {                 // two 4-row groups
        uint32_t acc2[6] = {0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
        for (int p2 = 0; p2 < 2; ++p2) {             // two row pairs
            const int r = q4 * 4 + p2 * 2;
            uint32_t A[12], B[12], T[12];
            load(r, A);
            load(r + 1, B);
            level1_rowpair(A, B, T);
            if (want1) {
                uint32_t Wd[6];
                icon1_words(T, Wd);
                em.icon1(r, Wd);
            }
            level2_accumulate(T, acc2);
        }
        if (want2) {
            uint32_t Wd[3];
            icon2_words(acc2, Wd);
            em.icon2(q4 * 4, Wd);
        }
        level3_accumulate(acc2, acc3);
    }
    if (want3) {
        uint32_t Hh[3];
        icon3_halves(acc3, Hh);
        em.icon3(0, Hh);
    }
    acc[0] = acc[1] = acc[2] = 0u;
    level4_accumulate(acc3, acc);
}

// Levels 4..6 when a warp owns half an item (128 x 32; lane = cx + 8*ry owns rows ry*8..ry*8+7).
//   s4 = 16x16 sums (lanes ry even), s5 = 32x32 sums (cx even, ry == 0),
//   s6 = 64x64 sums, valid on the UPPER warp of the pair for lanes cx in {0, 4}, ry == 0.
WHD void emit_tail_half(const IconSink& sk, int x0, int y0, int cx, int ry, bool upper, const uint32_t (&s4)[3],
                        const uint32_t (&s5)[3], const uint32_t (&s6)[3]) {
    if (sk.icon[3] != nullptr && (ry & 1) == 0) {
        const int oy = y0 >> 4, ox = x0 >> 4;
        if (oy < sk.h[3] && ox < sk.w[3]) {
            uint8_t* p = sk.icon[3] + (int64_t)oy * sk.pitch[3] + (int64_t)ox * 3;
            p[0] = (uint8_t)(s4[0] >> 8); p[1] = (uint8_t)(s4[1] >> 8); p[2] = (uint8_t)(s4[2] >> 8);
        }
    }
    if (sk.icon[4] != nullptr && (cx & 1) == 0 && ry == 0) {
        const int oy = y0 >> 5, ox = x0 >> 5;
        if (oy < sk.h[4] && ox < sk.w[4]) {
            uint8_t* p = sk.icon[4] + (int64_t)oy * sk.pitch[4] + (int64_t)ox * 3;
            p[0] = (uint8_t)(s5[0] >> 10); p[1] = (uint8_t)(s5[1] >> 10); p[2] = (uint8_t)(s5[2] >> 10);
        }
    }
    if (sk.icon[5] != nullptr && upper && (cx & 3) == 0 && ry == 0) {
        const int oy = y0 >> 6, ox = x0 >> 6;
        if (oy < sk.h[5] && ox < sk.w[5]) {
            uint8_t* p = sk.icon[5] + (int64_t)oy * sk.pitch[5] + (int64_t)ox * 3;
            p[0] = (uint8_t)(s6[0] >> 12); p[1] = (uint8_t)(s6[1] >> 12); p[2] = (uint8_t)(s6[2] >> 12);
        }
    }
    if (sk.sum6 != nullptr && upper && (cx & 3) == 0 && ry == 0) {
        const int oy = y0 >> 6, ox = x0 >> 6;
        if (oy < sk.sum6_h && ox < sk.sum6_w) {
            uint32_t* p = sk.sum6 + ((int64_t)oy * sk.sum6_stride + ox) * 3;
            p[0] = s6[0]; p[1] = s6[1]; p[2] = s6[2];
        }
    }
}

// ---------------------------------------------------------------------------
// cv::borderInterpolate for p >= 0 (padding is bottom/right only,
// wicca/data_loader.py:107-117).  Returns -1 for BORDER_CONSTANT.
// ---------------------------------------------------------------------------
WHD int border_index(int p, int n, int border_type) {
    if (p < n) return p;
    switch (border_type) {
        case 1: return n - 1;                               // REPLICATE
        case 2:                                             // REFLECT      fedcba|abcdefgh|hgfedcb
        case 4: {                                           // REFLECT_101  gfedcb|abcdefgh|gfedcba
            if (n == 1) return 0;
            const int delta = (border_type == 4) ? 1 : 0;
            do {
                if (p < 0) p = -p - 1 + delta;
                else       p = n - 1 - (p - n) - delta;
            } while ((unsigned)p >= (unsigned)n);
            return p;
        }
        case 3: return p % n;                               // WRAP
        default: return -1;                                 // CONSTANT
    }
}

// ---------------------------------------------------------------------------
// Input side: where the 48 bytes of a chunk row come from.
//   kSmem  - interior chunk: rows < H from the TMA-filled shared-memory stage, border rows
//            (y >= H) from the mapped image row in global memory;
//            (y >= H) from the stage too for REPLICATE (row H-1 is always in the same band);
//   kEdge  - chunk that touches or lies right of column W, border REPLICATE or CONSTANT: valid
//            bytes from the stage, the rest patched with pixel W-1 (same stage) / the constant;
//   kStrip - same position, border REFLECT / WRAP / REFLECT_101: every row from the pre-built
//            border-extended right strip (edge_strip_kernel);
//   kDead  - chunk beyond the padded extent of the deepest requested level: zeros.
// ---------------------------------------------------------------------------
enum ChunkMode : int { kDead = 0, kSmem = 1, kStrip = 2, kEdge = 3 };

struct ChunkSrc {
    int mode;
    const uint8_t* smem;      // chunk origin inside the stage (row stride kStageRowBytes)
    const uint8_t* gcol;      // kSmem: image + x0*3 ; kStrip: strip + (x0 - Wa)*3   (row 0)
    int64_t gpitch;           // row pitch of gcol's plane
    int x0, y0, rows, H, Hp_max;     // rows = rows owned by the lane (8)
    int nvalid;               // kEdge: pixels of the chunk that are inside the image (0..15)
    int rep_off;              // kEdge: byte offset of pixel W-1 relative to the chunk's row start (may be < 0)
    int border_type;
    uint32_t fill;            // border constant replicated in 4 bytes
};

WHD void words_from(const uint4& a, const uint4& b, const uint4& c, uint32_t (&A)[12]) {
    A[0] = a.x; A[1] = a.y; A[2] = a.z; A[3] = a.w;
    A[4] = b.x; A[5] = b.y; A[6] = b.z; A[7] = b.w;
    A[8] = c.x; A[9] = c.y; A[10] = c.z; A[11] = c.w;
}

// 48 bytes from the shared-memory stage (16-byte aligned)
WHD void load48_stage(const uint8_t* p, uint32_t (&A)[12]) {
#if defined(__CUDA_ARCH__)
    const uint4* q = reinterpret_cast<const uint4*>(p);
    words_from(q[0], q[1], q[2], A);
#else
    memcpy(A, p, 48);
#endif
}

// 48 bytes from global memory (16-byte aligned), read-only path
WHD void load48_global(const uint8_t* p, uint32_t (&A)[12]) {
#if defined(__CUDA_ARCH__)
    uint4 a, b, c;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p));
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p + 16));
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(p + 32));
    words_from(a, b, c, A);
#else
    memcpy(A, p, 48);
#endif
}

// The per-image fields the consumers need (kept in registers across work items).
struct ImageGeom {
    const uint8_t* src;
    int64_t pitch;
    int H, W, Hp_max, Wp_max, items_x;
};
WHD ImageGeom make_geom(const IconImage& im) {
    ImageGeom g;
    g.src = im.src; g.pitch = im.pitch; g.H = im.H; g.W = im.W;
    g.Hp_max = im.Hp_max; g.Wp_max = im.Wp_max; g.items_x = im.items_x;
    return g;
}

// Geometry of lane (cx, ry) of work item (ix, iy) of image `im`.
// row0 = first row of the lane inside the item, rows = how many rows it owns.
WHD ChunkSrc make_chunk_src(const ImageGeom& im, const uint8_t* strip, const uint8_t* stage, int ix, int iy, int cx,
                            int row0, int rows, int border_type, uint32_t fill) {
    ChunkSrc cs;
    cs.x0 = ix * kItemW + cx * kChunkPx;
    cs.y0 = iy * kItemH + row0;
    cs.rows = rows;
    cs.smem = stage + row0 * kStageRowBytes + cx * (kChunkPx * 3);
    cs.H = im.H; cs.Hp_max = im.Hp_max;
    cs.border_type = border_type; cs.fill = fill;
    cs.nvalid = 0; cs.rep_off = 0;
    const int Wa = im.W & ~(kChunkPx - 1);      // first pixel of the chunk that straddles W
    if (cs.x0 >= im.Wp_max || cs.y0 >= im.Hp_max) {
        cs.mode = kDead; cs.gcol = nullptr; cs.gpitch = 0;
    } else if (cs.x0 + kChunkPx <= im.W) {
        cs.mode = kSmem; cs.gcol = im.src + (int64_t)cs.x0 * 3; cs.gpitch = im.pitch;
    } else if (border_type == 1 || border_type == 0) {
        cs.mode = kEdge; cs.gcol = nullptr; cs.gpitch = 0;
        cs.nvalid = im.W > cs.x0 ? im.W - cs.x0 : 0;
        cs.rep_off = (im.W - 1 - cs.x0) * 3;
    } else {
        cs.mode = kStrip; cs.gcol = strip + (int64_t)(cs.x0 - Wa) * 3; cs.gpitch = kStripPitch;
    }
    return cs;
}

WHD bool chunk_is_interior(const ChunkSrc& cs) { return cs.mode == kSmem && cs.y0 + cs.rows <= cs.H; }

// General row loader (any mode, any row).
WHD void load_chunk_row(const ChunkSrc& cs, int r, uint32_t (&A)[12]) {
    const int y = cs.y0 + r;
    if (cs.mode == kSmem && y < cs.H) { load48_stage(cs.smem + r * kStageRowBytes, A); return; }
    if (cs.mode == kDead || y >= cs.Hp_max) {
#pragma unroll
        for (int k = 0; k < 12; ++k) A[k] = 0u;
        return;
    }
    if (cs.border_type == 0 && y >= cs.H) {              // CONSTANT: rows below the image
#pragma unroll
        for (int k = 0; k < 12; ++k) A[k] = cs.fill;
        return;
    }
    // REPLICATE: rows below the image repeat row H-1, which lies in the same 64-row band (stage)
    const int rr = (cs.border_type == 1 && y >= cs.H) ? r - (y - (cs.H - 1)) : r;
    if (cs.mode == kSmem && cs.border_type == 1) { load48_stage(cs.smem + rr * kStageRowBytes, A); return; }
    if (cs.mode == kEdge) {
        const uint8_t* p = cs.smem + rr * kStageRowBytes;
        if (cs.nvalid > 0) load48_stage(p, A);
        else {
#pragma unroll
            for (int k = 0; k < 12; ++k) A[k] = 0u;
        }
        uint32_t R[3];
        if (cs.border_type == 1) {
            const uint32_t c0 = p[cs.rep_off], c1 = p[cs.rep_off + 1], c2 = p[cs.rep_off + 2];
            R[0] = c0 | (c1 << 8) | (c2 << 16) | (c0 << 24);
            R[1] = c1 | (c2 << 8) | (c0 << 16) | (c1 << 24);
            R[2] = c2 | (c0 << 8) | (c1 << 16) | (c2 << 24);
        } else {
            R[0] = R[1] = R[2] = cs.fill;
        }
        const int vb = cs.nvalid * 3;
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const int nb = vb - 4 * k;
            const uint32_t mask = nb >= 4 ? 0xFFFFFFFFu : (nb <= 0 ? 0u : ((1u << (8 * nb)) - 1u));
            A[k] = bitsel(mask, A[k], R[k % 3]);
        }
        return;
    }
    const int ym = border_index(y, cs.H, cs.border_type);
    if (ym < 0) {
#pragma unroll
        for (int k = 0; k < 12; ++k) A[k] = cs.fill;
        return;
    }
    load48_global(cs.gcol + (int64_t)ym * cs.gpitch, A);
}

// One lane's share of a work item: its 16 px x 8 rows, levels 1..3 emitted, 16 x 8 sums returned.
WHD void reduce_lane(const ChunkSrc& cs, const StagedEmit& em, uint32_t (&acc)[3]) {
    if (chunk_is_interior(cs)) {
        const uint8_t* sm = cs.smem;
        auto loader = [sm](int r, uint32_t (&A)[12]) { load48_stage(sm + r * kStageRowBytes, A); };
        reduce_chunk(loader, em, acc);
    } else {
        auto loader = [&cs](int r, uint32_t (&A)[12]) { load_chunk_row(cs, r, A); };
        reduce_chunk(loader, em, acc);
    }
}

WHD IconSink make_sink(const IconImage& im) {
    IconSink sk;
#pragma unroll
    for (int l = 0; l < kMaxFused; ++l) {
        sk.icon[l] = im.icon[l]; sk.pitch[l] = im.icon_pitch[l]; sk.h[l] = im.icon_h[l]; sk.w[l] = im.icon_w[l];
    }
    sk.sum6 = im.sum6; sk.sum6_h = im.sum6_h; sk.sum6_w = im.sum6_w; sk.sum6_stride = im.sum6_stride;
    return sk;
}

// Right-edge strip: strip[y][(x - Wa)*3 + c] = padded_image[y][x][c] for x in [Wa, Wp_max), Wa = W & ~15.
// One call writes the 4 bytes of word j (0..kStripPitch/4-1) of strip row y.
WHD void strip_word(const IconImage& im, uint8_t* strip, int y, int j, int border_type, int border_const) {
    const int Wa = im.W & ~(kChunkPx - 1);
    const int npx = im.Wp_max - Wa;
    const uint8_t* row = im.src + (int64_t)y * im.pitch;
    uint32_t w = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int b = 4 * j + t;
        const int px = b / 3, c = b - 3 * px;
        uint32_t v = 0;
        if (px < npx) {
            const int xm = border_index(Wa + px, im.W, border_type);
            v = (xm < 0) ? (uint32_t)border_const : (uint32_t)row[(int64_t)xm * 3 + c];
        }
        w |= v << (8 * t);
    }
    *reinterpret_cast<uint32_t*>(strip + (int64_t)y * kStripPitch + 4 * j) = w;
}


}  // namespace wicca
