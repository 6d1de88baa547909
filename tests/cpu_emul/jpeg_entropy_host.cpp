// jpeg_entropy_host.cpp - HOST entropy decoders for JPEG scans.  TEST INFRASTRUCTURE ONLY: the product decodes
// Huffman scans on the GPU (wicca_b200/csrc/jpeg_huffman.cu) and reports progressive / multi-scan files as
// WICCA_EUNSUPPORTED; these serial decoders are the checker the GPU decoder is compared with coefficient by
// coefficient (tests/test_gpu_jpeg.py), themselves pinned on oracle/jpeg_oracle.py (tests/test_oracle_jpeg.py).
// Built by tests/conftest.py together with the product's marker parser (wicca_b200/csrc/jpeg_host.cu, compiled as C++)
// into tests/cpu_emul/_build/libemul_jpeg.so; never linked into libwicca_b200.so.
#include <string.h>

#include <string>

#include "host_common.h"
#include "jpeg_host.h"

namespace wicca {

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// MSB-first bit reader: the next bit of the stream is bit 63 of `acc`; `bits` of them are valid.
struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc = 0;
    int bits = 0;

    // After refill() at least 57 bits are valid (zeros are fed in front of a marker / past the end).
    inline void refill() {
        if (p + 8 <= end) {
            uint64_t x;
            memcpy(&x, p, 8);
            const uint64_t inv = ~x;                                     // a 0xFF byte of x is a zero byte of inv
            if (!((inv - 0x0101010101010101ull) & ~inv & 0x8080808080808080ull)) {
                const int take = (64 - bits) >> 3;                       // whole bytes that fit
                acc |= (__builtin_bswap64(x) >> bits) & ~((1ull << (64 - bits - 8 * take)) - 1ull);
                p += take;
                bits += 8 * take;
                return;
            }
        }
        while (bits <= 56) {
            uint64_t b = 0;
            if (p < end) {
                b = *p;
                if (b == 0xFF) {
                    if (p + 1 < end && p[1] == 0) p += 2;                // stuffed zero
                    else b = 0;                                          // a marker: feed zeros, stay in front of it
                } else {
                    ++p;
                }
            }
            acc |= b << (56 - bits);
            bits += 8;
        }
    }
    inline uint32_t peek(int k) const { return (uint32_t)(acc >> (64 - k)); }
    inline void skip(int k) { acc <<= k; bits -= k; }
    inline int receive_extend(int s) {                                   // T.81 F.2.2.1 EXTEND, branch-free; s in 1..15
        const int v = (int)peek(s);
        skip(s);
        return v + (((v - (1 << (s - 1))) >> 31) & (1 - (1 << s)));
    }
    inline int decode(const JpegHuff& h) {
        const uint32_t e = h.look[peek(10)];
        if (e) { skip((int)(e >> 8)); return (int)(e & 255); }
        int len = 11;
        int32_t code = (int32_t)peek(11);
        while (code > h.maxcode[len]) { ++len; if (len > 16) return -1; code = (int32_t)peek(len); }
        skip(len);
        return h.symbols[(code + h.valoffset[len]) & 255];
    }
};

}  // namespace

int jpeg_decode_coefficients(const uint8_t* data, size_t len, const JpegFrame& f, int16_t* dst, std::string& why) {
    BitReader br;
    br.p = data + f.scan_offset;
    br.end = data + len;
    int pred[3] = {0, 0, 0};
    int64_t count = 0;
    for (int my = 0; my < f.mcuy; ++my) {
        for (int mx = 0; mx < f.mcux; ++mx, ++count) {
            if (f.restart_interval && count && count % f.restart_interval == 0) {
                // byte-align, step over the RSTn marker, reset the predictors
                const uint8_t* q = br.p;
                while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) ++q;
                if (q + 1 >= br.end) { why = "restart marker missing"; return WICCA_EINVAL; }
                br.p = q + 2; br.acc = 0; br.bits = 0;
                pred[0] = pred[1] = pred[2] = 0;
            }
            for (int c = 0; c < f.ncomp; ++c) {
                const JpegComponent& q = f.comp[c];
                const JpegHuff& hd = f.dc[q.td];
                const JpegHuff& ha = f.ac[q.ta];
                for (int by = 0; by < q.v; ++by) {
                    for (int bx = 0; bx < q.h; ++bx) {
                        int16_t* blk = dst + q.coef_offset + ((int64_t)(my * q.v + by) * q.blocks_w + (mx * q.h + bx)) * 64;
                        memset(blk, 0, 64 * sizeof(int16_t));
                        if (br.bits < 32) br.refill();
                        int s = br.decode(hd);
                        if (s < 0 || s > 15) { why = "corrupt JPEG data (DC code)"; return WICCA_EINVAL; }
                        if (s) pred[c] += br.receive_extend(s);
                        blk[0] = (int16_t)pred[c];
                        for (int k = 1; k < 64;) {
                            if (br.bits < 32) br.refill();
                            const int fast = ha.fast_ac[br.peek(10)];
                            if (fast) {                                     // code + value in one lookup
                                k += (fast >> 4) & 15;
                                if (k > 63) { why = "corrupt JPEG data (run past the block)"; return WICCA_EINVAL; }
                                br.skip(fast & 15);
                                blk[kZigzag[k]] = (int16_t)(fast >> 8);
                                ++k;
                                continue;
                            }
                            const int rs = br.decode(ha);
                            if (rs < 0) { why = "corrupt JPEG data (AC code)"; return WICCA_EINVAL; }
                            const int r = rs >> 4;
                            s = rs & 15;
                            if (s) {
                                k += r;
                                if (k > 63) { why = "corrupt JPEG data (run past the block)"; return WICCA_EINVAL; }
                                blk[kZigzag[k]] = (int16_t)br.receive_extend(s);
                                ++k;
                            } else {
                                if (r != 15) break;                        // end of block
                                k += 16;
                            }
                        }
                    }
                }
            }
        }
    }
    return 0;
}

int jpeg_decode_multiscan(const uint8_t* data, size_t len, const JpegFrame& f, int16_t* dst, std::string& why) {
    (void)len;
    memset(dst, 0, (size_t)f.total_coefs * sizeof(int16_t));
    for (const JpegScan& sc : f.scans) {
        BitReader br;
        br.p = data + sc.data_offset;
        br.end = data + sc.data_end;
        const int p1 = 1 << sc.al, m1 = -(1 << sc.al);
        int pred[3] = {0, 0, 0};
        int eobrun = 0;
        // units: whole MCUs for an interleaved scan, else the blocks of the component that cover real samples
        const bool inter = sc.ns > 1;
        const JpegComponent& q0 = f.comp[sc.comp[0]];
        const int ux = inter ? f.mcux : (q0.dw + 7) / 8, uy = inter ? f.mcuy : (q0.dh + 7) / 8;
        int64_t count = 0;
        for (int y = 0; y < uy; ++y) {
            for (int x = 0; x < ux; ++x, ++count) {
                if (sc.restart_interval && count && count % sc.restart_interval == 0) {
                    const uint8_t* q = br.p;
                    while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) ++q;
                    if (q + 1 >= br.end) { why = "restart marker missing"; return WICCA_EINVAL; }
                    br.p = q + 2; br.acc = 0; br.bits = 0;
                    pred[0] = pred[1] = pred[2] = 0;
                    eobrun = 0;
                }
                for (int k = 0; k < sc.ns; ++k) {
                    const int ci = sc.comp[k];
                    const JpegComponent& q = f.comp[ci];
                    const JpegHuff& hd = sc.dc[sc.td[k]];
                    const JpegHuff& ha = sc.ac[sc.ta[k]];
                    const int nby = inter ? q.v : 1, nbx = inter ? q.h : 1;
                    for (int by = 0; by < nby; ++by) {
                        for (int bx = 0; bx < nbx; ++bx) {
                            const int row = inter ? y * q.v + by : y, col = inter ? x * q.h + bx : x;
                            int16_t* blk = dst + q.coef_offset + ((int64_t)row * q.blocks_w + col) * 64;
                            if (!f.progressive) {
                                if (br.bits < 32) br.refill();
                                int s = br.decode(hd);
                                if (s < 0 || s > 15) { why = "corrupt JPEG data (DC code)"; return WICCA_EINVAL; }
                                if (s) pred[ci] += br.receive_extend(s);
                                blk[0] = (int16_t)pred[ci];
                                for (int z = 1; z < 64;) {
                                    if (br.bits < 32) br.refill();
                                    const int rs = br.decode(ha);
                                    if (rs < 0) { why = "corrupt JPEG data (AC code)"; return WICCA_EINVAL; }
                                    const int r = rs >> 4;
                                    s = rs & 15;
                                    if (s) {
                                        z += r;
                                        if (z > 63) { why = "corrupt JPEG data (run past the block)"; return WICCA_EINVAL; }
                                        blk[kZigzag[z]] = (int16_t)br.receive_extend(s);
                                        ++z;
                                    } else {
                                        if (r != 15) break;
                                        z += 16;
                                    }
                                }
                            } else if (sc.ss == 0 && sc.ah == 0) {                 // DC, first pass
                                if (br.bits < 32) br.refill();
                                const int s = br.decode(hd);
                                if (s < 0 || s > 15) { why = "corrupt JPEG data (DC code)"; return WICCA_EINVAL; }
                                if (s) pred[ci] += br.receive_extend(s);
                                blk[0] = (int16_t)(pred[ci] * p1);
                            } else if (sc.ss == 0) {                                // DC, refinement: one bit
                                if (br.bits < 32) br.refill();
                                if (br.peek(1)) blk[0] = (int16_t)(blk[0] | p1);
                                br.skip(1);
                            } else if (sc.ah == 0) {                                // AC band, first pass
                                if (eobrun > 0) { --eobrun; continue; }
                                for (int z = sc.ss; z <= sc.se; ++z) {
                                    if (br.bits < 32) br.refill();
                                    const int rs = br.decode(ha);
                                    if (rs < 0) { why = "corrupt JPEG data (AC code)"; return WICCA_EINVAL; }
                                    const int r = rs >> 4, s = rs & 15;
                                    if (s) {
                                        z += r;
                                        if (z > 63) { why = "corrupt JPEG data (run past the block)"; return WICCA_EINVAL; }
                                        blk[kZigzag[z]] = (int16_t)(br.receive_extend(s) * p1);
                                    } else if (r == 15) {
                                        z += 15;
                                    } else {                                        // end of band for 2^r + extra blocks
                                        eobrun = 1 << r;
                                        if (r) { eobrun += (int)br.peek(r); br.skip(r); }
                                        --eobrun;
                                        break;
                                    }
                                }
                            } else {                                                // AC band, refinement
                                int z = sc.ss;
                                auto refine = [&](int16_t& c) {                     // one correction bit for a nonzero coefficient
                                    if (br.bits < 32) br.refill();
                                    const uint32_t bit = br.peek(1);
                                    br.skip(1);
                                    if (bit && (c & p1) == 0) c = (int16_t)(c + (c >= 0 ? p1 : m1));
                                };
                                if (eobrun == 0) {
                                    for (; z <= sc.se; ++z) {
                                        if (br.bits < 32) br.refill();
                                        const int rs = br.decode(ha);
                                        if (rs < 0) { why = "corrupt JPEG data (AC code)"; return WICCA_EINVAL; }
                                        int r = rs >> 4, s = rs & 15;
                                        if (s) {
                                            s = br.peek(1) ? p1 : m1;              // a new +-1 coefficient
                                            br.skip(1);
                                        } else if (r != 15) {
                                            eobrun = 1 << r;
                                            if (r) { eobrun += (int)br.peek(r); br.skip(r); }
                                            break;
                                        }
                                        // skip r coefficients that are still zero, correcting the nonzero ones on the way
                                        for (; z <= sc.se; ++z) {
                                            int16_t& c = blk[kZigzag[z]];
                                            if (c != 0) refine(c);
                                            else if (--r < 0) break;
                                        }
                                        if (s && z <= 63) blk[kZigzag[z]] = (int16_t)s;
                                    }
                                }
                                if (eobrun > 0) {
                                    for (; z <= sc.se; ++z) {
                                        int16_t& c = blk[kZigzag[z]];
                                        if (c != 0) refine(c);
                                    }
                                    --eobrun;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    return 0;
}

}  // namespace wicca

// Dense quantised coefficients of a file (per component, blocks in raster order, 64 coefficients each in natural
// order), decoded on the host.  Returns 0 or a negative WICCA_E* code; `why` (nullable, 256 bytes) gets the reason.
extern "C" __attribute__((visibility("default"))) long long emul_jpeg_coeff_count(const uint8_t* data, size_t len) {
    wicca::JpegFrame f;
    std::string why;
    const int rc = wicca::jpeg_parse(data, len, f, why);
    return rc ? (long long)rc : (long long)f.total_coefs;
}

extern "C" __attribute__((visibility("default"))) int emul_jpeg_decode_coeffs(const uint8_t* data, size_t len, int16_t* dst,
                                                                              long long dst_count, int* blocks_w, int* blocks_h,
                                                                              uint16_t* qt, char* why_out) {
    wicca::JpegFrame f;
    std::string why;
    int rc = wicca::jpeg_parse(data, len, f, why);
    if (!rc && (!dst || dst_count < f.total_coefs)) { rc = WICCA_EINVAL; why = "coefficient buffer too small"; }
    if (!rc) {
        for (int k = 0; k < f.ncomp; ++k) {
            if (blocks_w) blocks_w[k] = f.comp[k].blocks_w;
            if (blocks_h) blocks_h[k] = f.comp[k].blocks_h;
            if (qt) memcpy(qt + 64 * k, f.qt[f.comp[k].tq], 64 * sizeof(uint16_t));
        }
        rc = f.multiscan ? wicca::jpeg_decode_multiscan(data, len, f, dst, why) : wicca::jpeg_decode_coefficients(data, len, f, dst, why);
    }
    if (why_out) { strncpy(why_out, why.c_str(), 255); why_out[255] = 0; }
    return rc;
}
