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
