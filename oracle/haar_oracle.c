/*
 * haar_oracle.c - plain C restatement of the reference HaarCoder hot path (CPU oracle).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ as a fast checker at the
 * full BASELINE sizes.  Not linked into, loaded by, or called from the product.
 *
 * Restates (paths relative to /root/reference):
 *   oracle_border_index   cv::borderInterpolate as used by cv2.copyMakeBorder at
 *                         wicca/data_loader.py:116-117 (bottom/right padding only, :107-110)
 *   oracle_haar_icon_u8   HaarCoder.get_small_copy, wicca/wavelet_coder.py:56-67: pad to a
 *                         multiple of 2^depth, float32, depth x { (even_row + odd_row),
 *                         (even_col + odd_col) * 0.25 }, clip, truncate to uint8
 * Pinned against the live reference through tests/golden/haar_icon_golden.npz
 * (tests/test_oracle_c.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int oracle_border_index(int p, int n, int border_type) {
    border_type &= ~16;                                  /* BORDER_ISOLATED is masked off */
    if (p >= 0 && p < n) return p;
    switch (border_type) {
        case 1: return p < 0 ? 0 : n - 1;                /* REPLICATE */
        case 2:                                          /* REFLECT */
        case 4: {                                        /* REFLECT_101 */
            const int delta = border_type == 4;
            if (n == 1) return 0;
            do {
                if (p < 0) p = -p - 1 + delta;
                else p = n - 1 - (p - n) - delta;
            } while (p < 0 || p >= n);
            return p;
        }
        case 3:                                          /* WRAP */
            if (p < 0) p -= ((p - n + 1) / n) * n;
            if (p >= n) p %= n;
            return p;
        default: return -1;                              /* CONSTANT */
    }
}

/* src: (H, W, C) uint8 with row stride `stride` bytes; dst: tight (ceil(H/2^d), ceil(W/2^d), C).
 * border_const is already saturated to 0..255.  Returns 0, or -1 when out of memory. */
int oracle_haar_icon_u8(const uint8_t* src, int H, int W, int C, int64_t stride, int depth, int border_type,
                        int border_const, uint8_t* dst) {
    if (depth <= 0) {                                    /* zero levels: astype/clip/astype == copy */
        for (int y = 0; y < H; ++y) memcpy(dst + (size_t)y * W * C, src + (size_t)y * stride, (size_t)W * C);
        return 0;
    }
    const int64_t r = (int64_t)1 << depth;
    int64_t h = (H + r - 1) / r * r, w = (W + r - 1) / r * r;        /* data_loader.py:107-110 */
    float* cur = (float*)malloc((size_t)h * w * C * sizeof(float));  /* .astype(np.float32), wavelet_coder.py:59 */
    if (!cur) return -1;
    int* xmap = (int*)malloc((size_t)w * sizeof(int));
    if (!xmap) { free(cur); return -1; }
    for (int64_t x = 0; x < w; ++x) xmap[x] = oracle_border_index((int)x, W, border_type);
    for (int64_t y = 0; y < h; ++y) {
        const int ym = oracle_border_index((int)y, H, border_type);
        float* row = cur + (size_t)y * w * C;
        for (int64_t x = 0; x < w; ++x) {
            const int xm = xmap[x];
            for (int c = 0; c < C; ++c)
                row[x * C + c] = (ym < 0 || xm < 0) ? (float)border_const : (float)src[(size_t)ym * stride + (size_t)xm * C + c];
        }
    }
    free(xmap);
    for (int l = 0; l < depth; ++l) {                                /* wavelet_coder.py:61-65 */
        const int64_t nh = h / 2, nw = w / 2;
        float* nxt = (float*)malloc((size_t)nh * nw * C * sizeof(float));
        if (!nxt) { free(cur); return -1; }
        for (int64_t y = 0; y < nh; ++y) {
            const float* r0 = cur + (size_t)(2 * y) * w * C;
            const float* r1 = r0 + (size_t)w * C;
            float* o = nxt + (size_t)y * nw * C;
            for (int64_t x = 0; x < nw; ++x)
                for (int c = 0; c < C; ++c) {
                    const volatile float s0 = r0[(2 * x) * C + c] + r1[(2 * x) * C + c];          /* sums[:, even] */
                    const volatile float s1 = r0[(2 * x + 1) * C + c] + r1[(2 * x + 1) * C + c];  /* sums[:, odd]  */
                    const volatile float t = s0 + s1;
                    o[x * C + c] = t * 0.25f;
                }
        }
        free(cur);
        cur = nxt; h = nh; w = nw;
    }
    for (int64_t i = 0; i < h * w * C; ++i) {                        /* np.clip(...).astype(np.uint8), :67 */
        float v = cur[i];
        v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
        dst[i] = (uint8_t)v;                                         /* truncation toward zero */
    }
    free(cur);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * oracle_haar_icons_multi_u8 - icons at several depths (each 1..8) of one image in ONE pass, through
 * the integer identity  icon_d = (sum of the 2^d x 2^d block of the border-extended image) >> 2d
 * (exact because every float32 intermediate of wavelet_coder.py:61-65 is
 * k / 4^level with k < 2^24).  The border-extended images of different depths nest (the border rule
 * is a pure function of the padded index, data_loader.py:107-117), so one pyramid of exact uint32
 * block sums serves every depth.  A fast checker for whole batches (bench.py checks all 180 icons of
 * an end-to-end step with it); pinned against oracle_haar_icon_u8 and the reference goldens in
 * tests/test_oracle_c.py.  dsts[k]: tight (ceil(H/2^depths[k]), ceil(W/2^depths[k]), C).
 * Returns 0, -1 out of memory, -2 bad depth. */
int oracle_haar_icons_multi_u8(const uint8_t* src, int H, int W, int C, int64_t stride, const int* depths, int n_depths,
                               int border_type, int border_const, uint8_t* const* dsts) {
    int dmax = 0;
    for (int k = 0; k < n_depths; ++k) {
        if (depths[k] < 1 || depths[k] > 8) return -2;
        if (depths[k] > dmax) dmax = depths[k];
    }
    if (n_depths <= 0) return 0;
    const int64_t r = (int64_t)1 << dmax;
    const int64_t hp = (H + r - 1) / r * r, wp = (W + r - 1) / r * r;
    int* xmap = (int*)malloc((size_t)wp * sizeof(int));
    uint32_t* cur = (uint32_t*)malloc((size_t)(hp / 2) * (wp / 2) * C * sizeof(uint32_t));
    if (!xmap || !cur) { free(xmap); free(cur); return -1; }
    for (int64_t x = 0; x < wp; ++x) xmap[x] = oracle_border_index((int)x, W, border_type);
    /* level 1 from the border-extended pixels */
    for (int64_t y = 0; y < hp / 2; ++y) {
        uint32_t* o = cur + (size_t)y * (wp / 2) * C;
        memset(o, 0, (size_t)(wp / 2) * C * sizeof(uint32_t));
        for (int dy = 0; dy < 2; ++dy) {
            const int ym = oracle_border_index((int)(2 * y + dy), H, border_type);
            const uint8_t* row = ym < 0 ? NULL : src + (size_t)ym * stride;
            for (int64_t x = 0; x < wp; ++x) {
                const int xm = xmap[x];
                for (int c = 0; c < C; ++c)
                    o[(x >> 1) * C + c] += (row && xm >= 0) ? row[(size_t)xm * C + c] : (uint32_t)border_const;
            }
        }
    }
    int64_t h = hp / 2, w = wp / 2;
    for (int l = 1; l <= dmax; ++l) {
        for (int k = 0; k < n_depths; ++k) {
            if (depths[k] != l) continue;
            const int64_t oh = (H + ((int64_t)1 << l) - 1) >> l, ow = (W + ((int64_t)1 << l) - 1) >> l;
            for (int64_t y = 0; y < oh; ++y)
                for (int64_t i = 0; i < ow * C; ++i)
                    dsts[k][(size_t)y * ow * C + i] = (uint8_t)(cur[(size_t)y * w * C + i] >> (2 * l));
        }
        if (l == dmax) break;
        const int64_t nh = h / 2, nw = w / 2;               /* in place: the write index never passes the reads */
        for (int64_t y = 0; y < nh; ++y)
            for (int64_t x = 0; x < nw; ++x)
                for (int c = 0; c < C; ++c) {
                    const uint32_t* a = cur + ((size_t)(2 * y) * w + 2 * x) * C + c;
                    const uint32_t s = a[0] + a[C] + a[(size_t)w * C] + a[(size_t)w * C + C];
                    cur[((size_t)y * nw + x) * C + c] = s;
                }
        h = nh; w = nw;
    }
    free(xmap); free(cur);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Extension (no reference implementation): full sub-band analysis / synthesis,
 * the C twin of haar_oracle.haar_forward / haar_inverse.  Per level, with the 2x2 block  a b / c d :
 *   rs = a + c, rs' = b + d (row-pair sums, wavelet_coder.py:62-63), rd = a - c, rd' = b - d
 *   LL = (rs + rs') * 0.25 (wavelet_coder.py:64-65)   HL = (rs - rs') * 0.25
 *   LH = (rd + rd') * 0.25                            HH = (rd - rd') * 0.25
 * every operation rounded separately in float32.  plane: float32 (Hp, Wp, C), Mallat arrangement
 * (include/wicca_b200.h: LL_l top-left, HL_l right of it, LH_l below, HH_l diagonal). */
int oracle_haar_forward_f32(const uint8_t* src, int H, int W, int C, int64_t stride, int depth, int border_type,
                            int border_const, float* plane) {
    if (depth < 1) return -2;
    const int64_t r = (int64_t)1 << depth;
    const int64_t hp = (H + r - 1) / r * r, wp = (W + r - 1) / r * r;
    int* xmap = (int*)malloc((size_t)wp * sizeof(int));
    if (!xmap) return -1;
    for (int64_t x = 0; x < wp; ++x) xmap[x] = oracle_border_index((int)x, W, border_type);
    for (int64_t y = 0; y < hp; ++y) {
        const int ym = oracle_border_index((int)y, H, border_type);
        float* row = plane + (size_t)y * wp * C;
        for (int64_t x = 0; x < wp; ++x) {
            const int xm = xmap[x];
            for (int c = 0; c < C; ++c)
                row[x * C + c] = (ym < 0 || xm < 0) ? (float)border_const : (float)src[(size_t)ym * stride + (size_t)xm * C + c];
        }
    }
    free(xmap);
    int64_t h = hp, w = wp;
    for (int l = 0; l < depth; ++l) {
        const int64_t nh = h / 2, nw = w / 2;
        /* rows 2y, 2y+1 of the LL block are consumed before rows y and nh+y are written; y <= 2y, but nh+y may
         * lie ahead of the read position, so the whole level goes through a scratch copy of the LL block */
        float* ll = (float*)malloc((size_t)h * w * C * sizeof(float));
        if (!ll) return -1;
        for (int64_t y = 0; y < h; ++y) memcpy(ll + (size_t)y * w * C, plane + (size_t)y * wp * C, (size_t)w * C * sizeof(float));
        for (int64_t y = 0; y < nh; ++y) {
            const float* r0 = ll + (size_t)(2 * y) * w * C;
            const float* r1 = r0 + (size_t)w * C;
            float* top = plane + (size_t)y * wp * C;
            float* bot = plane + (size_t)(nh + y) * wp * C;
            for (int64_t x = 0; x < nw; ++x)
                for (int c = 0; c < C; ++c) {
                    const volatile float a = r0[(2 * x) * C + c], b = r0[(2 * x + 1) * C + c];
                    const volatile float cc = r1[(2 * x) * C + c], d = r1[(2 * x + 1) * C + c];
                    const volatile float rs0 = a + cc, rs1 = b + d, rd0 = a - cc, rd1 = b - d;
                    const volatile float s = rs0 + rs1, t = rs0 - rs1, u = rd0 + rd1, v = rd0 - rd1;
                    top[x * C + c] = s * 0.25f;               /* LL */
                    top[(nw + x) * C + c] = t * 0.25f;        /* HL */
                    bot[x * C + c] = u * 0.25f;               /* LH */
                    bot[(nw + x) * C + c] = v * 0.25f;        /* HH */
                }
        }
        free(ll);
        h = nh; w = nw;
    }
    return 0;
}

/* Synthesis:  a = (LL+HL)+(LH+HH)  b = (LL-HL)+(LH-HH)  c = (LL+HL)-(LH+HH)  d = (LL-HL)-(LH-HH)
 * (haar_oracle.haar_inverse).  plane (Hp, Wp, C) Mallat -> image (Hp, Wp, C) float32. */
int oracle_haar_inverse_f32(const float* plane, int Hp, int Wp, int C, int depth, float* image) {
    if (depth < 1 || (Hp & ((1 << depth) - 1)) || (Wp & ((1 << depth) - 1))) return -2;
    const size_t rowf = (size_t)Wp * C;
    memcpy(image, plane, (size_t)Hp * rowf * sizeof(float));
    for (int l = depth; l >= 1; --l) {
        const int64_t h = Hp >> l, w = Wp >> l;              /* LL_l extent; output 2h x 2w */
        float* blk = (float*)malloc((size_t)4 * h * w * C * sizeof(float));
        if (!blk) return -1;
        for (int64_t y = 0; y < 2 * h; ++y) memcpy(blk + (size_t)y * 2 * w * C, image + (size_t)y * rowf, (size_t)2 * w * C * sizeof(float));
        for (int64_t y = 0; y < h; ++y)
            for (int64_t x = 0; x < w; ++x)
                for (int c = 0; c < C; ++c) {
                    const volatile float ll = blk[((size_t)y * 2 * w + x) * C + c], hl = blk[((size_t)y * 2 * w + w + x) * C + c];
                    const volatile float lh = blk[((size_t)(h + y) * 2 * w + x) * C + c], hh = blk[((size_t)(h + y) * 2 * w + w + x) * C + c];
                    const volatile float p = ll + hl, q = ll - hl, s = lh + hh, t = lh - hh;
                    image[(size_t)(2 * y) * rowf + (2 * x) * C + c] = p + s;
                    image[(size_t)(2 * y) * rowf + (2 * x + 1) * C + c] = q + t;
                    image[(size_t)(2 * y + 1) * rowf + (2 * x) * C + c] = p - s;
                    image[(size_t)(2 * y + 1) * rowf + (2 * x + 1) * C + c] = q - t;
                }
        free(blk);
    }
    return 0;
}
