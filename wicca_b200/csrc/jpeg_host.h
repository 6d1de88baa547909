// jpeg_host.h - host half of the JPEG ingest path (row N2: cv2.imread + BGR2RGB, wicca/data_loader.py:53-58):
// marker parsing and byte un-stuffing only.  Huffman decoding (jpeg_huffman.cu), dequantisation, the inverse DCT,
// chroma upsampling and the colour transform (jpeg_kernels.cu) run on the GPU, bit-exact with libjpeg-turbo's
// defaults, which is what cv2.imread uses.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace wicca {

struct JpegComponent {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int blocks_w = 0, blocks_h = 0;        // whole MCUs
    int dw = 0, dh = 0;                    // real ("downsampled") samples
    int64_t coef_offset = 0;               // first coefficient of the component in the dense array (int16 units)
};

struct JpegHuff {
    // 10-bit lookahead: (length << 8) | symbol for codes of <= 10 bits, 0 otherwise; then the canonical tables
    uint16_t look[1024];
    // AC tables only: when a code and the value bits that follow it fit in the 10 lookahead bits and the value
    // fits in 8 bits, (value << 8) | (run << 4) | (bits consumed); 0 otherwise
    int16_t fast_ac[1024];
    int32_t maxcode[18];                   // largest code of each length (-1: none), [17] = sentinel
    int32_t valoffset[17];
    uint8_t symbols[256];
    bool present = false;
};

struct JpegScan {                          // one scan of a progressive / multi-scan file
    int ns = 0, comp[3] = {0, 0, 0}, td[3] = {0, 0, 0}, ta[3] = {0, 0, 0};
    int ss = 0, se = 63, ah = 0, al = 0;
    int restart_interval = 0;
    size_t data_offset = 0, data_end = 0;  // entropy-coded bytes
    JpegHuff dc[4], ac[4];                 // the tables in force when the scan starts
};

struct JpegFrame {
    int width = 0, height = 0, ncomp = 0;
    int hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
    int restart_interval = 0;
    int orientation = 1;                   // EXIF tag 0x0112 (1..8); cv2.imread applies it, so does this path
    JpegComponent comp[3];
    uint16_t qt[4][64];                    // natural (row-major) order
    bool qt_present[4] = {false, false, false, false};
    JpegHuff dc[4], ac[4];
    size_t scan_offset = 0;                // first entropy-coded byte (single-scan files)
    int64_t total_coefs = 0;               // int16 count of the dense coefficient array
    bool progressive = false;              // SOF2
    bool multiscan = false;                // progressive, or sequential with more than one scan: see `scans`
    std::vector<JpegScan> scans;
};

// Size of the image the caller receives: the frame size, transposed for EXIF orientations 5..8.
inline void jpeg_output_size(const JpegFrame& f, int* H, int* W) {
    const bool swap = f.orientation >= 5;
    *H = swap ? f.width : f.height;
    *W = swap ? f.height : f.width;
}

// 0 on success; a negative WICCA_* code otherwise, with the reason in `why`.  WICCA_EUNSUPPORTED marks valid
// JPEGs outside the subset (progressive, arithmetic coding, 12-bit, CMYK, several scans ...).
int jpeg_parse(const uint8_t* data, size_t len, JpegFrame& f, std::string& why);

// --- the two host entropy decoders below are NOT part of libwicca_b200.so: they are implemented in
// --- tests/cpu_emul/jpeg_entropy_host.cpp and exist only as the checker of the GPU Huffman decoder.
// All scans of a progressive / multi-scan file accumulated into dst[total_coefs] (same layout as below); this is
// what libjpeg holds when the whole file has been read, before it outputs the first row.
int jpeg_decode_multiscan(const uint8_t* data, size_t len, const JpegFrame& f, int16_t* dst, std::string& why);

// Huffman-decode the scan into dst[total_coefs]: per component, blocks in raster order, 64 coefficients each
// in natural order, not dequantised.  Every element of dst is written (no pre-zeroing needed).
int jpeg_decode_coefficients(const uint8_t* data, size_t len, const JpegFrame& f, int16_t* dst, std::string& why);

// Copy the entropy-coded bytes of the scan to dst with the byte stuffing (0xFF 0x00 -> 0xFF) and the RSTn markers
// removed, stopping at the first other marker; dst needs len - scan_offset + 16 bytes and is zero padded by 16
// bytes.  interval_starts receives the byte offset (in dst) at which every restart interval but the first begins.
// Returns the byte count.
size_t jpeg_unstuff_scan(const uint8_t* data, size_t len, const JpegFrame& f, uint8_t* dst, std::vector<uint32_t>* interval_starts);

}  // namespace wicca
