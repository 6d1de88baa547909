// jpeg_host.cu - JPEG marker parser and byte un-stuffing (host code; see jpeg_host.h).  Entropy decoding is done on
// the GPU (jpeg_huffman.cu); the host entropy decoders that check it live in tests/cpu_emul/jpeg_entropy_host.cpp.
// Subset: baseline / extended-sequential Huffman frames (SOF0, SOF1), 8-bit samples, one grey component or
// three YCbCr components in a single interleaved scan, integral sampling ratios, restart intervals.
// Anything else cv2.imread can read (progressive, arithmetic, CMYK ...) is reported as
// WICCA_EUNSUPPORTED so that the caller can route that file elsewhere - never decoded approximately.
#include "jpeg_host.h"

#include <string.h>

#include "host_common.h"

namespace wicca {

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

inline int be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

// EXIF orientation (tag 0x0112) of an APP1 segment, 1 when absent or unreadable.
int exif_orientation(const uint8_t* seg, size_t n) {
    if (n < 14 || memcmp(seg, "Exif\0\0", 6) != 0) return 1;
    const uint8_t* t = seg + 6;
    const size_t tn = n - 6;
    const bool le = t[0] == 'I' && t[1] == 'I';
    if (!le && !(t[0] == 'M' && t[1] == 'M')) return 1;
    auto u16 = [&](size_t o) -> uint32_t { return le ? (uint32_t)(t[o] | (t[o + 1] << 8)) : (uint32_t)((t[o] << 8) | t[o + 1]); };
    auto u32 = [&](size_t o) -> uint32_t {
        return le ? (uint32_t)t[o] | ((uint32_t)t[o + 1] << 8) | ((uint32_t)t[o + 2] << 16) | ((uint32_t)t[o + 3] << 24)
                  : ((uint32_t)t[o] << 24) | ((uint32_t)t[o + 1] << 16) | ((uint32_t)t[o + 2] << 8) | (uint32_t)t[o + 3];
    };
    const size_t ifd = u32(4);
    if (ifd + 2 > tn) return 1;
    const uint32_t entries = u16(ifd);
    for (uint32_t e = 0; e < entries; ++e) {
        const size_t o = ifd + 2 + 12 * (size_t)e;
        if (o + 12 > tn) break;
        if (u16(o) == 0x0112) return (int)u16(o + 8);
    }
    return 1;
}

int build_huff(const uint8_t* counts, const uint8_t* symbols, int n_symbols, JpegHuff& h) {
    memset(&h, 0, sizeof h);
    memcpy(h.symbols, symbols, (size_t)n_symbols);
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
        h.valoffset[len] = k - code;
        if (counts[len - 1]) {
            for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
                if (code >= (1 << len)) return -1;            // over-subscribed table
                if (len <= 10) {
                    const int first = code << (10 - len), span = 1 << (10 - len);
                    for (int j = 0; j < span; ++j) h.look[first + j] = (uint16_t)((len << 8) | symbols[k]);
                }
            }
            h.maxcode[len] = code - 1;
        } else {
            h.maxcode[len] = -1;
        }
        code <<= 1;
    }
    h.maxcode[17] = 0x7FFFFFFF;
    for (int i = 0; i < 1024; ++i) {
        const int e = h.look[i];
        if (!e) continue;
        const int len = e >> 8, run = (e >> 4) & 15, mag = e & 15;
        if (!mag || len + mag > 10) continue;
        int v = ((i << len) & 1023) >> (10 - mag);
        if (v < (1 << (mag - 1))) v += 1 - (1 << mag);
        if (v >= -128 && v <= 127) h.fast_ac[i] = (int16_t)(v * 256 + run * 16 + len + mag);
    }
    h.present = true;
    return 0;
}

}  // namespace

int jpeg_parse(const uint8_t* data, size_t len, JpegFrame& f, std::string& why) {
    f = JpegFrame();
    if (!data || len < 4 || data[0] != 0xFF || data[1] != 0xD8) { why = "not a JPEG stream (no SOI marker)"; return WICCA_EUNSUPPORTED; }
    bool have_frame = false, jfif = false, adobe = false;
    int adobe_transform = -1;
    size_t i = 2;
    for (;;) {
        if (i + 4 > len) {
            if (f.multiscan && i + 2 <= len) return 0;                     // the closing EOI (2 bytes) or a file cut short
            why = "truncated JPEG header"; return WICCA_EINVAL;
        }
        if (data[i] != 0xFF) { why = "JPEG marker expected"; return WICCA_EINVAL; }
        while (i + 1 < len && data[i + 1] == 0xFF) ++i;
        const int m = data[i + 1];
        i += 2;
        if (m == 0xD9) {
            if (f.multiscan) return 0;                                     // all scans collected
            why = "JPEG ends before any scan"; return WICCA_EINVAL;
        }
        if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;                  // stand-alone markers
        if (i + 2 > len) { why = "truncated JPEG header"; return WICCA_EINVAL; }
        const size_t seglen = (size_t)be16(data + i);
        if (seglen < 2 || i + seglen > len) { why = "bad JPEG segment length"; return WICCA_EINVAL; }
        const uint8_t* seg = data + i + 2;
        const size_t n = seglen - 2;
        i += seglen;
        switch (m) {
            case 0xDB: {
                size_t k = 0;
                while (k < n) {
                    const int pq = seg[k] >> 4, tq = seg[k] & 15;
                    ++k;
                    if (tq > 3 || k + (pq ? 128 : 64) > n) { why = "bad quantisation table"; return WICCA_EINVAL; }
                    for (int j = 0; j < 64; ++j) f.qt[tq][kZigzag[j]] = (uint16_t)(pq ? be16(seg + k + 2 * j) : seg[k + j]);
                    k += pq ? 128 : 64;
                    f.qt_present[tq] = true;
                }
                break;
            }
            case 0xC4: {
                size_t k = 0;
                while (k < n) {
                    if (k + 17 > n) { why = "bad Huffman table"; return WICCA_EINVAL; }
                    const int tc = seg[k] >> 4, th = seg[k] & 15;
                    int total = 0;
                    for (int j = 0; j < 16; ++j) total += seg[k + 1 + j];
                    if (tc > 1 || th > 3 || total > 256 || k + 17 + (size_t)total > n) { why = "bad Huffman table"; return WICCA_EINVAL; }
                    if (build_huff(seg + k + 1, seg + k + 17, total, tc ? f.ac[th] : f.dc[th])) { why = "bad Huffman table"; return WICCA_EINVAL; }
                    k += 17 + (size_t)total;
                }
                break;
            }
            case 0xC0: case 0xC1: case 0xC2: {
                if (have_frame) { why = "more than one frame header"; return WICCA_EINVAL; }
                f.progressive = (m == 0xC2);
                if (n < 6) { why = "bad frame header"; return WICCA_EINVAL; }
                if (seg[0] != 8) { why = "only 8-bit JPEG samples are supported"; return WICCA_EUNSUPPORTED; }
                f.height = be16(seg + 1); f.width = be16(seg + 3); f.ncomp = seg[5];
                if (f.height <= 0 || f.width <= 0) { why = "JPEG with zero size (DNL marker) is not supported"; return WICCA_EUNSUPPORTED; }
                if ((int64_t)f.height * f.width > ((int64_t)1 << 29)) { why = "JPEG larger than 512 megapixels"; return WICCA_EUNSUPPORTED; }
                if (f.ncomp != 1 && f.ncomp != 3) { why = "only 1- and 3-component JPEGs are supported"; return WICCA_EUNSUPPORTED; }
                if (n < 6 + 3 * (size_t)f.ncomp) { why = "bad frame header"; return WICCA_EINVAL; }
                for (int c = 0; c < f.ncomp; ++c) {
                    JpegComponent& q = f.comp[c];
                    q.id = seg[6 + 3 * c]; q.h = seg[7 + 3 * c] >> 4; q.v = seg[7 + 3 * c] & 15; q.tq = seg[8 + 3 * c];
                    if (q.h < 1 || q.h > 4 || q.v < 1 || q.v > 4 || q.tq > 3) { why = "bad component description"; return WICCA_EINVAL; }
                }
                have_frame = true;
                break;
            }
            case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                why = "lossless, differential and arithmetic-coded JPEGs are not supported";
                return WICCA_EUNSUPPORTED;
            case 0xDD:
                if (n < 2) { why = "bad restart interval"; return WICCA_EINVAL; }
                f.restart_interval = be16(seg);
                break;
            case 0xE0:
                if (n >= 5 && memcmp(seg, "JFIF\0", 5) == 0) jfif = true;
                break;
            case 0xE1:
                if (f.orientation == 1) {                                  // the first Exif segment decides, as in OpenCV
                    const int o = exif_orientation(seg, n);
                    if (o >= 2 && o <= 8) f.orientation = o;
                }
                break;
            case 0xEE:
                if (n >= 12 && memcmp(seg, "Adobe", 5) == 0) { adobe = true; adobe_transform = seg[11]; }
                break;
            case 0xDA: {
                if (!have_frame) { why = "scan before frame header"; return WICCA_EINVAL; }
                if (n < 1) { why = "bad scan header"; return WICCA_EINVAL; }
                const int ns = seg[0];
                if (ns < 1 || ns > f.ncomp || n < 1 + 2 * (size_t)ns + 3) { why = "bad scan header"; return WICCA_EINVAL; }
                if (f.scans.empty()) {
                    // geometry and colour space, once
                    if (f.ncomp == 3) {
                        // as libjpeg guesses it (jdapimin.c): JFIF => YCbCr; else Adobe transform 0 => RGB; else ids 'R','G','B' => RGB
                        const bool rgb_ids = f.comp[0].id == 'R' && f.comp[1].id == 'G' && f.comp[2].id == 'B';
                        if (!jfif && ((adobe && adobe_transform == 0) || (!adobe && rgb_ids))) { why = "RGB-coded JPEG (no YCbCr transform)"; return WICCA_EUNSUPPORTED; }
                    }
                    if (f.ncomp == 1) f.comp[0].h = f.comp[0].v = 1;      // a lone component is never interleaved
                    f.hmax = f.vmax = 1;
                    for (int c = 0; c < f.ncomp; ++c) { if (f.comp[c].h > f.hmax) f.hmax = f.comp[c].h; if (f.comp[c].v > f.vmax) f.vmax = f.comp[c].v; }
                    f.mcux = (f.width + 8 * f.hmax - 1) / (8 * f.hmax);
                    f.mcuy = (f.height + 8 * f.vmax - 1) / (8 * f.vmax);
                    int blocks_in_mcu = 0;
                    f.total_coefs = 0;
                    for (int c = 0; c < f.ncomp; ++c) {
                        JpegComponent& q = f.comp[c];
                        if (f.hmax % q.h || f.vmax % q.v) { why = "fractional chroma sampling ratio"; return WICCA_EUNSUPPORTED; }
                        if (!f.qt_present[q.tq]) { why = "frame refers to a missing quantisation table"; return WICCA_EINVAL; }
                        q.blocks_w = f.mcux * q.h; q.blocks_h = f.mcuy * q.v;
                        q.dw = (f.width * q.h + f.hmax - 1) / f.hmax; q.dh = (f.height * q.v + f.vmax - 1) / f.vmax;
                        q.coef_offset = f.total_coefs;
                        f.total_coefs += (int64_t)q.blocks_w * q.blocks_h * 64;
                        blocks_in_mcu += q.h * q.v;
                    }
                    if (blocks_in_mcu > 10) { why = "more than 10 blocks per MCU"; return WICCA_EINVAL; }
                }
                JpegScan sc;
                sc.ns = ns;
                for (int k = 0; k < ns; ++k) {
                    const int cid = seg[1 + 2 * k], tabs = seg[2 + 2 * k];
                    int ci = -1;
                    for (int c = 0; c < f.ncomp; ++c) if (f.comp[c].id == cid) ci = c;
                    if (ci < 0 || (k && ci <= sc.comp[k - 1])) { why = "scan components unknown or out of frame order"; return WICCA_EINVAL; }
                    sc.comp[k] = ci; sc.td[k] = tabs >> 4; sc.ta[k] = tabs & 15;
                    if (sc.td[k] > 3 || sc.ta[k] > 3) { why = "bad Huffman table id"; return WICCA_EINVAL; }
                }
                const uint8_t* tail = seg + 1 + 2 * ns;
                sc.ss = tail[0]; sc.se = tail[1]; sc.ah = tail[2] >> 4; sc.al = tail[2] & 15;
                sc.restart_interval = f.restart_interval;
                if (f.progressive) {
                    const bool dc_scan = sc.ss == 0;
                    if (sc.ss > sc.se || sc.se > 63 || (dc_scan && sc.se != 0) || (!dc_scan && ns != 1) || sc.al > 13 || (sc.ah && sc.ah != sc.al + 1)) { why = "bad progressive scan parameters"; return WICCA_EINVAL; }
                } else if (sc.ss != 0 || sc.se != 63 || sc.ah != 0 || sc.al != 0) {
                    why = "spectral selection / successive approximation in a sequential scan"; return WICCA_EINVAL;
                }
                for (int k = 0; k < ns; ++k) {
                    const bool need_dc = !f.progressive || sc.ss == 0, need_ac = !f.progressive || sc.ss > 0;
                    if ((need_dc && !(sc.ah && f.progressive) && !f.dc[sc.td[k]].present) || (need_ac && !f.ac[sc.ta[k]].present)) { why = "scan refers to a missing Huffman table"; return WICCA_EINVAL; }
                }
                if (!f.progressive && ns == f.ncomp && f.scans.empty()) {
                    // the common case: one interleaved sequential scan (the fast paths decode it in place)
                    for (int k = 0; k < ns; ++k) { f.comp[k].td = sc.td[k]; f.comp[k].ta = sc.ta[k]; }
                    f.scan_offset = i;
                    return 0;
                }
                // progressive / several scans: keep this scan with the tables in force, find where its data ends
                f.multiscan = true;
                for (int t = 0; t < 4; ++t) { sc.dc[t] = f.dc[t]; sc.ac[t] = f.ac[t]; }
                sc.data_offset = i;
                size_t j = i;
                for (;;) {                                                // to the next marker that is not RSTn
                    const uint8_t* q = (const uint8_t*)memchr(data + j, 0xFF, len - j);
                    if (!q || (size_t)(q - data) + 1 >= len) { j = len; break; }
                    j = (size_t)(q - data);
                    const uint8_t nx = data[j + 1];
                    if (nx == 0 || (nx >= 0xD0 && nx <= 0xD7)) { j += 2; continue; }   // stuffing / restart: scan data
                    if (nx == 0xFF) { j += 1; continue; }                             // fill byte
                    break;
                }
                sc.data_end = j < len ? j : len;
                f.scans.push_back(sc);
                if (f.scans.size() > 1000) { why = "too many scans"; return WICCA_EINVAL; }
                i = sc.data_end;
                if (i + 1 >= len) return 0;                               // no EOI: decode what is there
                break;
            }
            default:
                break;                                                    // APPn, COM, ... skipped
        }
    }
}

size_t jpeg_unstuff_scan(const uint8_t* data, size_t len, const JpegFrame& f, uint8_t* dst, std::vector<uint32_t>* interval_starts) {
    const uint8_t* p = data + f.scan_offset;
    const uint8_t* end = data + len;
    uint8_t* o = dst;
    while (p < end) {
        const uint8_t* q = (const uint8_t*)memchr(p, 0xFF, (size_t)(end - p));
        if (!q) q = end;
        memcpy(o, p, (size_t)(q - p));
        o += q - p;
        p = q;
        if (p >= end) break;
        if (p + 1 < end && p[1] == 0) { *o++ = 0xFF; p += 2; continue; }     // stuffed zero
        if (p + 1 < end && p[1] >= 0xD0 && p[1] <= 0xD7) {                     // RSTn: the next interval starts here
            if (interval_starts) interval_starts->push_back((uint32_t)(o - dst));
            p += 2;
            continue;
        }
        if (p + 1 < end && p[1] == 0xFF) { ++p; continue; }                    // fill byte in front of a marker
        break;                                                                 // any other marker (EOI): the scan ends here
    }
    memset(o, 0, 16);
    return (size_t)(o - dst);
}

}  // namespace wicca
