// emul_icon.cpp - CPU replay of the one-pass icon kernel (test infrastructure).
//
// Compiles wicca_b200/csrc/haar_math.cuh as plain C++ and walks every work item / lane exactly
// as haar_icon_tma2_kernel does: a 24 KB "stage" is filled the way the TMA box load would fill it
// (zero beyond the tensor extent, pad bytes of the pitched image included); for each of the two
// warps of the pair, each of the 32 lanes runs make_chunk_src / reduce_lane into the warp's
// output tile, the tile is copied to the icons with TMA store's clipping rule, the level 4/5/6
// shuffles and the mailbox add are replayed with the same xor pattern, and emit_tail_half stores
// the results.  What is NOT covered: mbarrier/TMA mechanics.
//
// Build:  g++ -O2 -shared -fPIC -I/usr/local/cuda/include -Iwicca_b200/csrc tests/cpu_emul/emul_icon.cpp
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "haar_math.cuh"

using namespace wicca;

extern "C" {

// src: tight (H, W, 3) uint8.  mask bit (d-1) requests depth d.  outs[d-1]: tight icon buffers
// (may be NULL when not requested).  Returns 0, or a positive code when a guard byte around an
// icon was overwritten (out-of-bounds store).
// sum6_out (nullable): receives the (ceil(H/64), ceil(W/64), 3) plane of exact level-6 block sums that feeds depths > 6.
int emul_fused_icons(const uint8_t* src, int H, int W, int border_type, int border_const, unsigned mask,
                     uint8_t** outs, uint32_t* sum6_out) {
    const int64_t pitch = ((int64_t)W * 3 + 127) / 128 * 128;
    std::vector<uint8_t> img((size_t)pitch * H + 64, 0xA5);      // pad bytes are garbage on purpose
    for (int y = 0; y < H; ++y) memcpy(&img[(size_t)y * pitch], src + (size_t)y * W * 3, (size_t)W * 3);

    IconImage im;
    memset(&im, 0, sizeof im);
    icon_image_geometry(&im, img.data(), H, W, pitch, 0);
    std::vector<std::vector<uint8_t>> icons(kMaxFused);
    const size_t guard = 256;
    for (int d = 1; d <= kMaxFused; ++d) {
        if (!(mask & (1u << (d - 1)))) continue;
        const int r = 1 << d;
        const int h = (H + r - 1) >> d, w = (W + r - 1) >> d;
        const int64_t ip = ((int64_t)w * 3 + 127) / 128 * 128;
        icons[d - 1].assign((size_t)ip * h + 2 * guard, 0xEE);
        icon_image_add_level(&im, d, icons[d - 1].data() + guard, ip);
    }
    if (sum6_out) icon_image_add_sum6(&im, sum6_out, 0);
    const int Wa = W & ~(kChunkPx - 1);
    const int npx = im.Wp_max - Wa;
    std::vector<uint8_t> strip;
    if (npx > 0 && border_type >= 2) {      // REPLICATE / CONSTANT never touch the strip
        strip.assign((size_t)H * kStripPitch, 0x5A);
        for (int y = 0; y < H; ++y)
            for (int j = 0; j < (npx * 3 + 3) / 4; ++j) strip_word(im, strip.data(), y, j, border_type, border_const);
    }
    const uint32_t fill = (uint32_t)border_const * 0x01010101u;
    const IconSink sk = make_sink(im);
    const ImageGeom geo = make_geom(im);
    std::vector<uint8_t> stage(kStageBytes);
    for (int iy = 0; iy < im.items_y; ++iy)
        for (int ix = 0; ix < im.items_x; ++ix) {
            // TMA box: (kStageRowBytes/4) uint32 x kItemH rows at element (ix*96, iy*64) of the
            // (pitch/4 x H) tensor; out-of-range elements are zero-filled.
            for (int r = 0; r < kItemH; ++r)
                for (int b = 0; b < kStageRowBytes; ++b) {
                    const int y = iy * kItemH + r;
                    const int64_t xb = (int64_t)ix * kStageRowBytes + b;
                    stage[(size_t)r * kStageRowBytes + b] = (y < H && xb < pitch) ? img[(size_t)y * pitch + xb] : 0;
                }

            // two warps per item: half 0 = rows 0..31, half 1 = rows 32..63; 8 rows per lane
            uint32_t s6half[2][32][3];
            ChunkSrc hcs[2][32];
            uint32_t hs4[2][32][3], hs5[2][32][3];
            for (int half = 0; half < 2; ++half) {
                std::vector<uint8_t> tile(kHalfStageBytes, 0xCD);
                uint32_t acc[32][3], a1[32][3], b1[32][3];
                for (int lane = 0; lane < 32; ++lane) {
                    const int cx = lane & 7, ry = lane >> 3;
                    hcs[half][lane] = make_chunk_src(geo, strip.empty() ? nullptr : strip.data(), stage.data(), ix, iy,
                                                     cx, half * 32 + ry * 8, 8, border_type, fill);
                    const StagedEmit em = staged_emit_half(tile.data(), cx, ry, mask & 7u);
                    reduce_lane(hcs[half][lane], em, acc[lane]);
                }
                const int box_w[3] = {kOut1Row, kOut2Row, kOut3Row}, box_h[3] = {16, 8, 4};
                const int off[3] = {kHalf1Off, kHalf2Off, kHalf3Off};
                for (int l = 0; l < 3; ++l) {
                    if (!(mask & (1u << l))) continue;
                    const int64_t wb = (int64_t)im.icon_w[l] * 3;
                    for (int r = 0; r < box_h[l]; ++r)
                        for (int b = 0; b < box_w[l]; ++b) {
                            const int64_t gx = (int64_t)ix * box_w[l] + b;
                            const int gy = iy * (2 * box_h[l]) + half * box_h[l] + r;
                            if (gx < wb && gy < im.icon_h[l])
                                im.icon[l][(int64_t)gy * im.icon_pitch[l] + gx] = tile[off[l] + r * box_w[l] + b];
                        }
                }
                for (int c = 0; c < 3; ++c) {
                    for (int l = 0; l < 32; ++l) hs4[half][l][c] = acc[l][c] + acc[l ^ 8][c];
                    for (int l = 0; l < 32; ++l) a1[l][c] = hs4[half][l][c] + hs4[half][l ^ 1][c];
                    for (int l = 0; l < 32; ++l) hs5[half][l][c] = a1[l][c] + a1[l ^ 16][c];
                    for (int l = 0; l < 32; ++l) b1[l][c] = hs5[half][l][c] + hs5[half][l ^ 2][c];
                    for (int l = 0; l < 32; ++l) s6half[half][l][c] = b1[l][c];
                }
            }
            for (int half = 0; half < 2; ++half)
                for (int lane = 0; lane < 32; ++lane) {
                    uint32_t s6v[3];
                    for (int c = 0; c < 3; ++c) s6v[c] = s6half[0][lane][c] + s6half[1][lane][c];
                    emit_tail_half(sk, hcs[half][lane].x0, hcs[half][lane].y0, lane & 7, lane >> 3, half == 0,
                                   hs4[half][lane], hs5[half][lane], s6v);
                }
        }
    int bad = 0;
    for (int d = 1; d <= kMaxFused; ++d) {
        if (!(mask & (1u << (d - 1)))) continue;
        const int h = im.icon_h[d - 1], w = im.icon_w[d - 1];
        const int64_t ip = im.icon_pitch[d - 1];
        const std::vector<uint8_t>& buf = icons[d - 1];
        for (size_t i = 0; i < guard; ++i)
            if (buf[i] != 0xEE || buf[buf.size() - 1 - i] != 0xEE) bad = 100 + d;
        for (int y = 0; y < h; ++y) {
            for (int64_t b = (int64_t)w * 3; b < ip; ++b)
                if (buf[guard + (size_t)y * ip + b] != 0xEE) bad = 200 + d;     // wrote into the row padding
            if (outs[d - 1]) memcpy(outs[d - 1] + (size_t)y * w * 3, &buf[guard + (size_t)y * ip], (size_t)w * 3);
        }
    }
    return bad;
}

}  // extern "C"
