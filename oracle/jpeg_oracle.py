"""TEST INFRASTRUCTURE ONLY - never imported by the product (wicca_b200/).

CPU restatement of what ``cv2.imread`` + ``cv2.cvtColor(BGR2RGB)`` (the reference's ``load_image``,
wicca/data_loader.py:53-58) does to a baseline JPEG file.  The arithmetic lives in a third-party dependency that
is not in /root/reference: libjpeg-turbo, bundled in the pinned ``opencv-python==4.12.0.88`` wheel
(requirements.txt:91; the wheel installed here, 4.13.0, bundles libjpeg-turbo 3.1.2).  OpenCV leaves the
decompressor at its defaults, so the published algorithm restated here is:

  * Huffman decoding of a sequential DCT frame (ITU T.81 F.2.2), DC prediction per component, restart markers;
    for progressive frames (SOF2) the DC / AC first and refinement scans of T.81 G.1.2 (jdphuff.c), all scans
    accumulated before anything is output, as libjpeg does when the whole file is available;
  * dequantisation + the accurate integer IDCT ``jpeg_idct_islow`` (jidctint.c: CONST_BITS = 13,
    PASS1_BITS = 2, the Loeffler-Ligtenberg-Moschytz factorisation);
  * "fancy" triangle-filter upsampling for 2:1 horizontal (``h2v1_fancy_upsample``) and 2:1 x 2:1
    (``h2v2_fancy_upsample``) chroma when the chroma plane is wider than 2 samples, box replication
    otherwise (jdsample.c); rows above / below the image are copies of the first / last real row (jdmainct.c);
  * YCbCr -> RGB with the 16-bit fixed-point tables of jdcolor.c.

Pinning: ``tests/test_oracle_jpeg.py`` checks this restatement (i) against ``tests/golden/jpeg_golden.npz`` -
JPEG files and what the reference's own ``load_image`` returned for them, written by
``tests/golden/make_golden.py jpeg`` in the build container - and (ii) against ``cv2.imdecode`` run live, bit
for bit over sizes, qualities, chroma samplings, restart intervals, EXIF orientations and grayscale files.
The Huffman decoder is a plain Python loop: small images only.
"""
from __future__ import annotations

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
                   21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53,
                   60, 61, 54, 47, 55, 62, 63], dtype=np.int64)


class Unsupported(ValueError):
    """A JPEG flavour outside the restated subset (progressive, arithmetic, 12-bit, CMYK, odd samplings ...)."""


def parse(data: bytes) -> dict:
    """Markers up to and including the (single) scan: frame, tables, restart interval, entropy-coded bytes."""
    if data[:2] != b"\xff\xd8":
        raise Unsupported("not a JPEG (no SOI)")
    qt = {}
    ht = {}
    frame = None
    restart = 0
    i = 2
    while True:
        if data[i] != 0xFF:
            raise Unsupported("marker expected")
        while data[i + 1] == 0xFF:
            i += 1
        m = data[i + 1]
        i += 2
        if m == 0xD9:
            raise Unsupported("EOI before any scan")
        seglen = (data[i] << 8) | data[i + 1]
        seg = data[i + 2:i + seglen]
        i += seglen
        if m == 0xDB:
            k = 0
            while k < len(seg):
                pq, tq = seg[k] >> 4, seg[k] & 15
                k += 1
                if pq:
                    vals = [(seg[k + 2 * j] << 8) | seg[k + 2 * j + 1] for j in range(64)]
                    k += 128
                else:
                    vals = list(seg[k:k + 64])
                    k += 64
                t = np.zeros(64, dtype=np.int64)
                t[ZIGZAG] = vals                       # tables are stored in zigzag order
                qt[tq] = t
        elif m == 0xC4:
            k = 0
            while k < len(seg):
                tc, th = seg[k] >> 4, seg[k] & 15
                counts = list(seg[k + 1:k + 17])
                n = sum(counts)
                ht[(tc, th)] = (counts, list(seg[k + 17:k + 17 + n]))
                k += 17 + n
        elif m in (0xC0, 0xC1):
            if seg[0] != 8:
                raise Unsupported("only 8-bit samples")
            h, w, nc = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            comps = [{"id": seg[6 + 3 * c], "h": seg[7 + 3 * c] >> 4, "v": seg[7 + 3 * c] & 15, "tq": seg[8 + 3 * c]} for c in range(nc)]
            frame = {"h": h, "w": w, "comps": comps}
        elif m in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise Unsupported("only baseline / extended sequential Huffman frames")
        elif m == 0xDD:
            restart = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            ns = seg[0]
            if frame is None or ns != len(frame["comps"]):
                raise Unsupported("only one interleaved scan")
            for c in range(ns):
                cid, tabs = seg[1 + 2 * c], seg[2 + 2 * c]
                comp = [x for x in frame["comps"] if x["id"] == cid][0]
                comp["td"], comp["ta"] = tabs >> 4, tabs & 15
            return {"frame": frame, "qt": qt, "ht": ht, "restart": restart, "scan": data[i:]}
        # every other segment (APPn, COM ...) is skipped


def _huff_lookup(counts, symbols):
    """code length / value -> symbol (T.81 Annex C)."""
    table = {}
    code = 0
    k = 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            table[(length, code)] = symbols[k]
            code += 1
            k += 1
        code <<= 1
    return table


class _Bits:
    def __init__(self, data: bytes):
        self.d, self.i, self.acc, self.n = data, 0, 0, 0

    def bit(self) -> int:
        if self.n == 0:
            b = self.d[self.i] if self.i < len(self.d) else 0
            self.i += 1
            if b == 0xFF:
                nxt = self.d[self.i] if self.i < len(self.d) else 0
                if nxt == 0:
                    self.i += 1                      # stuffed zero
                else:
                    self.i -= 1                      # a marker: feed zeros (libjpeg does the same)
                    b = 0
            self.acc, self.n = b, 8
        self.n -= 1
        return (self.acc >> self.n) & 1

    def bits(self, k: int) -> int:
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v

    def restart(self):
        """Skip to just after the next RSTn marker."""
        self.n = 0
        while not (self.d[self.i] == 0xFF and 0xD0 <= self.d[self.i + 1] <= 0xD7):
            self.i += 1
        self.i += 2


def _decode_symbol(br: _Bits, table) -> int:
    code = 0
    for length in range(1, 17):
        code = (code << 1) | br.bit()
        s = table.get((length, code))
        if s is not None:
            return s
    raise ValueError("bad Huffman code")


def _extend(v: int, s: int) -> int:
    return v if v >= (1 << (s - 1)) else v - (1 << s) + 1


def decode_coefficients(data: bytes) -> dict:
    """Quantised coefficients per component, natural (row-major) order: arrays (blocks_y, blocks_x, 64) int32
    covering whole MCUs, plus the frame geometry and the quantisation tables."""
    p = parse(data)
    fr = p["frame"]
    comps = fr["comps"]
    hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
    mcux, mcuy = -(-fr["w"] // (8 * hmax)), -(-fr["h"] // (8 * vmax))
    if len(comps) == 1:                                # a single-component scan is never interleaved: 1 block per MCU
        comps[0]["h"] = comps[0]["v"] = 1
        hmax = vmax = 1
        mcux, mcuy = -(-fr["w"] // 8), -(-fr["h"] // 8)
    tabs = {k: _huff_lookup(*v) for k, v in p["ht"].items()}
    coefs = [np.zeros((mcuy * c["v"], mcux * c["h"], 64), dtype=np.int32) for c in comps]
    br = _Bits(p["scan"])
    pred = [0] * len(comps)
    count = 0
    for my in range(mcuy):
        for mx in range(mcux):
            if p["restart"] and count and count % p["restart"] == 0:
                br.restart()
                pred = [0] * len(comps)
            count += 1
            for ci, c in enumerate(comps):
                dc_t, ac_t = tabs[(0, c["td"])], tabs[(1, c["ta"])]
                for by in range(c["v"]):
                    for bx in range(c["h"]):
                        blk = coefs[ci][my * c["v"] + by, mx * c["h"] + bx]
                        s = _decode_symbol(br, dc_t)
                        if s:
                            pred[ci] += _extend(br.bits(s), s)
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = _decode_symbol(br, ac_t)
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = _extend(br.bits(s), s)
                            k += 1
    return {"w": fr["w"], "h": fr["h"], "comps": comps, "hmax": hmax, "vmax": vmax, "coefs": coefs,
            "qt": [p["qt"][c["tq"]] for c in comps]}


# ------------------------------------------------------------------ multi-scan files: progressive (jdphuff.c)
def parse_all(data: bytes) -> dict:
    """Every scan of the file with the tables in force when it starts (tables may be redefined between scans)."""
    if data[:2] != b"\xff\xd8":
        raise Unsupported("not a JPEG (no SOI)")
    qt, ht = {}, {}
    frame, restart, scans = None, 0, []
    i = 2
    while i + 4 <= len(data):
        if data[i] != 0xFF:
            raise Unsupported("marker expected")
        while data[i + 1] == 0xFF:
            i += 1
        m = data[i + 1]
        i += 2
        if m == 0xD9:
            break
        seglen = (data[i] << 8) | data[i + 1]
        seg = data[i + 2:i + seglen]
        i += seglen
        if m == 0xDB:
            k = 0
            while k < len(seg):
                pq, tq = seg[k] >> 4, seg[k] & 15
                k += 1
                vals = [(seg[k + 2 * j] << 8) | seg[k + 2 * j + 1] for j in range(64)] if pq else list(seg[k:k + 64])
                k += 128 if pq else 64
                t = np.zeros(64, dtype=np.int64)
                t[ZIGZAG] = vals
                qt[tq] = t
        elif m == 0xC4:
            k = 0
            while k < len(seg):
                tc, th = seg[k] >> 4, seg[k] & 15
                counts = list(seg[k + 1:k + 17])
                n = sum(counts)
                ht[(tc, th)] = _huff_lookup(counts, list(seg[k + 17:k + 17 + n]))
                k += 17 + n
        elif m in (0xC0, 0xC1, 0xC2):
            if seg[0] != 8:
                raise Unsupported("only 8-bit samples")
            h, w, nc = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            comps = [{"id": seg[6 + 3 * c], "h": seg[7 + 3 * c] >> 4, "v": seg[7 + 3 * c] & 15, "tq": seg[8 + 3 * c]} for c in range(nc)]
            frame = {"h": h, "w": w, "comps": comps, "progressive": m == 0xC2}
        elif m in (0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise Unsupported("lossless / differential / arithmetic frames")
        elif m == 0xDD:
            restart = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            ns = seg[0]
            sel = []
            for c in range(ns):
                cid, tabs = seg[1 + 2 * c], seg[2 + 2 * c]
                ci = [x["id"] for x in frame["comps"]].index(cid)
                sel.append((ci, tabs >> 4, tabs & 15))
            ss, se, ahal = seg[1 + 2 * ns], seg[2 + 2 * ns], seg[3 + 2 * ns]
            j = i                                      # entropy-coded data runs to the next marker that is not RSTn
            while not (data[j] == 0xFF and data[j + 1] != 0 and not 0xD0 <= data[j + 1] <= 0xD7):
                j += 1
            scans.append({"comps": sel, "ss": ss, "se": se, "ah": ahal >> 4, "al": ahal & 15, "data": data[i:j],
                          "ht": dict(ht), "restart": restart})
            i = j
    return {"frame": frame, "qt": qt, "scans": scans}


def decode_coefficients_multiscan(data: bytes) -> dict:
    """Coefficients of a progressive (or multi-scan sequential) file: all scans accumulated, as libjpeg does before
    it outputs anything when the whole file is available (no block smoothing then)."""
    p = parse_all(data)
    fr = p["frame"]
    comps = fr["comps"]
    if len(comps) == 1:
        comps[0]["h"] = comps[0]["v"] = 1
    hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
    mcux, mcuy = -(-fr["w"] // (8 * hmax)), -(-fr["h"] // (8 * vmax))
    coefs = [np.zeros((mcuy * c["v"], mcux * c["h"], 64), dtype=np.int32) for c in comps]
    for sc in p["scans"]:
        br = _Bits(sc["data"])
        ss, se, ah, al = sc["ss"], sc["se"], sc["ah"], sc["al"]
        sel = sc["comps"]
        if len(sel) > 1:                               # interleaved: whole MCUs
            units = [(my, mx) for my in range(mcuy) for mx in range(mcux)]
            def blocks_of(u):
                my, mx = u
                return [(ci, td, ta, my * comps[ci]["v"] + by, mx * comps[ci]["h"] + bx)
                        for (ci, td, ta) in sel for by in range(comps[ci]["v"]) for bx in range(comps[ci]["h"])]
        else:                                          # one component: only the blocks that cover real samples
            ci, td, ta = sel[0]
            bw = -(-(-(-fr["w"] * comps[ci]["h"] // hmax)) // 8)
            bh = -(-(-(-fr["h"] * comps[ci]["v"] // vmax)) // 8)
            units = [(by, bx) for by in range(bh) for bx in range(bw)]
            def blocks_of(u, ci=ci, td=td, ta=ta):
                return [(ci, td, ta, u[0], u[1])]
        pred = [0] * len(comps)
        eobrun = 0
        p1, m1 = 1 << al, -(1 << al)
        for n, u in enumerate(units):
            if sc["restart"] and n and n % sc["restart"] == 0:
                br.restart()
                pred = [0] * len(comps)
                eobrun = 0
            for (ci, td, ta, by, bx) in blocks_of(u):
                blk = coefs[ci][by, bx]
                if not fr["progressive"]:              # sequential scan of a multi-scan file
                    s = _decode_symbol(br, sc["ht"][(0, td)])
                    if s:
                        pred[ci] += _extend(br.bits(s), s)
                    blk[0] = pred[ci]
                    k = 1
                    while k < 64:
                        rs = _decode_symbol(br, sc["ht"][(1, ta)])
                        r, s = rs >> 4, rs & 15
                        if s == 0:
                            if r != 15:
                                break
                            k += 16
                            continue
                        k += r
                        blk[ZIGZAG[k]] = _extend(br.bits(s), s)
                        k += 1
                elif ss == 0 and ah == 0:              # DC first
                    s = _decode_symbol(br, sc["ht"][(0, td)])
                    if s:
                        pred[ci] += _extend(br.bits(s), s)
                    blk[0] = pred[ci] << al
                elif ss == 0:                          # DC refinement
                    if br.bit():
                        blk[0] |= p1
                elif ah == 0:                          # AC first
                    if eobrun > 0:
                        eobrun -= 1
                        continue
                    k = ss
                    while k <= se:
                        rs = _decode_symbol(br, sc["ht"][(1, ta)])
                        r, s = rs >> 4, rs & 15
                        if s:
                            k += r
                            blk[ZIGZAG[k]] = _extend(br.bits(s), s) << al
                        elif r == 15:
                            k += 15
                        else:
                            eobrun = (1 << r) + (br.bits(r) if r else 0) - 1
                            break
                        k += 1
                else:                                  # AC refinement
                    k = ss
                    if eobrun == 0:
                        while k <= se:
                            rs = _decode_symbol(br, sc["ht"][(1, ta)])
                            r, s = rs >> 4, rs & 15
                            if s:
                                s = p1 if br.bit() else m1
                            elif r != 15:
                                eobrun = (1 << r) + (br.bits(r) if r else 0)
                                break
                            while k <= se:
                                z = ZIGZAG[k]
                                if blk[z] != 0:
                                    if br.bit() and (blk[z] & p1) == 0:
                                        blk[z] += p1 if blk[z] >= 0 else m1
                                else:
                                    r -= 1
                                    if r < 0:
                                        break
                                k += 1
                            if s:
                                blk[ZIGZAG[k]] = s
                            k += 1
                    if eobrun > 0:
                        while k <= se:
                            z = ZIGZAG[k]
                            if blk[z] != 0 and br.bit() and (blk[z] & p1) == 0:
                                blk[z] += p1 if blk[z] >= 0 else m1
                            k += 1
                        eobrun -= 1
    return {"w": fr["w"], "h": fr["h"], "comps": comps, "hmax": hmax, "vmax": vmax, "coefs": coefs,
            "qt": [p["qt"][c["tq"]] for c in comps]}


# ------------------------------------------------------------------ jidctint.c: jpeg_idct_islow
_F = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137,
          f1_961=16069, f2_053=16819, f2_562=20995, f3_072=25172)


def _idct_1d(v, shift):
    """One pass over the leading axis of v (8, ...) int64; returns the 8 outputs descaled by `shift`."""
    z2, z3 = v[2], v[6]
    z1 = (z2 + z3) * _F["f0_541"]
    tmp2 = z1 - z3 * _F["f1_847"]
    tmp3 = z1 + z2 * _F["f0_765"]
    tmp0 = (v[0] + v[4]) << 13
    tmp1 = (v[0] - v[4]) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = v[7], v[5], v[3], v[1]
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * _F["f1_175"]
    t0 = t0 * _F["f0_298"]; t1 = t1 * _F["f2_053"]; t2 = t2 * _F["f3_072"]; t3 = t3 * _F["f1_501"]
    z1 = -z1 * _F["f0_899"]; z2 = -z2 * _F["f2_562"]; z3 = -z3 * _F["f1_961"] + z5; z4 = -z4 * _F["f0_390"] + z5
    t0 = t0 + z1 + z3; t1 = t1 + z2 + z4; t2 = t2 + z2 + z3; t3 = t3 + z1 + z4
    half = 1 << (shift - 1)
    out = [tmp10 + t3, tmp11 + t2, tmp12 + t1, tmp13 + t0, tmp13 - t0, tmp12 - t1, tmp11 - t2, tmp10 - t3]
    return np.stack([(o + half) >> shift for o in out])


def _range_limit(x):
    """sample_range_limit + CENTERJSAMPLE indexed with (x & RANGE_MASK): clamp(x + 128) inside [-512, 511], and the
    table's wrap-around outside (only corrupt data gets there)."""
    idx = x & 1023
    out = np.where(idx < 128, idx + 128, np.where(idx < 512, 255, np.where(idx < 896, 0, idx - 896)))
    return out.astype(np.uint8)


def idct_islow(coefs: np.ndarray, qt: np.ndarray) -> np.ndarray:
    """(by, bx, 64) quantised coefficients -> (by*8, bx*8) uint8 samples."""
    by, bx, _ = coefs.shape
    v = (coefs.astype(np.int64) * qt.astype(np.int64)).reshape(by, bx, 8, 8)       # [.., row, col]
    ws = _idct_1d(np.moveaxis(v, 2, 0), 13 - 2)                                  # pass 1: along columns (over rows)
    ws = np.moveaxis(ws, 0, 2)                                                   # back to [.., row, col]
    res = _idct_1d(np.moveaxis(ws, 3, 0), 13 + 2 + 3)                            # pass 2: along rows (over cols)
    res = np.moveaxis(res, 0, 3)
    return _range_limit(res).transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)


# ------------------------------------------------------------------ jdsample.c
def _h2v1_fancy(p: np.ndarray) -> np.ndarray:
    """(rows, n) -> (rows, 2n): 3/4 nearer + 1/4 farther, biases 1 and 2; the two end samples are copied."""
    x = p.astype(np.int32)
    n = x.shape[1]
    out = np.empty((x.shape[0], 2 * n), dtype=np.int32)
    left = np.concatenate([x[:, :1], x[:, :-1]], axis=1)
    right = np.concatenate([x[:, 1:], x[:, -1:]], axis=1)
    out[:, 0::2] = (3 * x + left + 1) >> 2
    out[:, 1::2] = (3 * x + right + 2) >> 2
    out[:, 0] = x[:, 0]
    out[:, -1] = x[:, -1]
    return out.astype(np.uint8)


def _h2v2_fancy(p: np.ndarray) -> np.ndarray:
    """(n_rows, n) -> (2 n_rows, 2n): vertical 3:1 blend with the nearer / farther row (edge rows replicated),
    then horizontal 3:1 with biases 8 and 7 on the 16x column sums."""
    x = p.astype(np.int32)
    up = np.concatenate([x[:1], x[:-1]], axis=0)
    down = np.concatenate([x[1:], x[-1:]], axis=0)
    rows = np.empty((2 * x.shape[0], x.shape[1]), dtype=np.int32)
    rows[0::2] = 3 * x + up
    rows[1::2] = 3 * x + down
    n = x.shape[1]
    out = np.empty((rows.shape[0], 2 * n), dtype=np.int32)
    last = np.concatenate([rows[:, :1], rows[:, :-1]], axis=1)
    nxt = np.concatenate([rows[:, 1:], rows[:, -1:]], axis=1)
    out[:, 0::2] = (3 * rows + last + 8) >> 4
    out[:, 1::2] = (3 * rows + nxt + 7) >> 4
    out[:, 0] = (rows[:, 0] * 4 + 8) >> 4
    out[:, -1] = (rows[:, -1] * 4 + 7) >> 4
    return out.astype(np.uint8)


def _h1v2_fancy(p: np.ndarray) -> np.ndarray:
    """(n_rows, n) -> (2 n_rows, n): 3:1 vertical blend, bias 1 for the upper output row and 2 for the lower."""
    x = p.astype(np.int32)
    up = np.concatenate([x[:1], x[:-1]], axis=0)
    down = np.concatenate([x[1:], x[-1:]], axis=0)
    out = np.empty((2 * x.shape[0], x.shape[1]), dtype=np.int32)
    out[0::2] = (3 * x + up + 1) >> 2
    out[1::2] = (3 * x + down + 2) >> 2
    return out.astype(np.uint8)


def upsample(plane: np.ndarray, hf: int, vf: int) -> np.ndarray:
    """plane: the component's real samples (downsampled_height, downsampled_width); hf, vf: expansion factors."""
    if hf == 1 and vf == 1:
        return plane
    fancy = plane.shape[1] > 2
    if hf == 2 and vf == 1:
        return _h2v1_fancy(plane) if fancy else np.repeat(plane, 2, axis=1)
    if hf == 2 and vf == 2:
        return _h2v2_fancy(plane) if fancy else np.repeat(np.repeat(plane, 2, axis=0), 2, axis=1)
    if hf == 1 and vf == 2:
        return _h1v2_fancy(plane)
    return np.repeat(np.repeat(plane, vf, axis=0), hf, axis=1)


# ------------------------------------------------------------------ jdcolor.c
def ycc_to_rgb(y: np.ndarray, cb: np.ndarray, cr: np.ndarray) -> np.ndarray:
    def fix(x):
        return int(x * 65536 + 0.5)
    half = 1 << 15
    yb, b, r = y.astype(np.int64), cb.astype(np.int64) - 128, cr.astype(np.int64) - 128
    rr = yb + ((fix(1.40200) * r + half) >> 16)
    bb = yb + ((fix(1.77200) * b + half) >> 16)
    gg = yb + ((-fix(0.34414) * b + half - fix(0.71414) * r) >> 16)
    return np.clip(np.stack([rr, gg, bb], axis=-1), 0, 255).astype(np.uint8)


def decode_rgb(data: bytes) -> np.ndarray:
    """What ``cv2.cvtColor(cv2.imdecode(data, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)`` returns (no EXIF rotation)."""
    try:
        d = decode_coefficients(data)
    except Unsupported:
        d = decode_coefficients_multiscan(data)        # progressive / several scans
    planes = []
    for c, coefs, qt in zip(d["comps"], d["coefs"], d["qt"]):
        full = idct_islow(coefs, qt)
        dh = -(-d["h"] * c["v"] // d["vmax"])
        dw = -(-d["w"] * c["h"] // d["hmax"])
        up = upsample(full[:dh, :dw], d["hmax"] // c["h"], d["vmax"] // c["v"])
        planes.append(up[:d["h"], :d["w"]])
    if len(planes) == 1:
        return np.repeat(planes[0][:, :, None], 3, axis=2)
    return ycc_to_rgb(*planes)


def exif_orientation(data: bytes) -> int:
    """EXIF tag 0x0112 of the first Exif APP1 segment (1 when absent), as OpenCV's ExifReader finds it."""
    i = 2
    while i + 4 <= len(data) and data[i] == 0xFF:
        m = data[i + 1]
        if m == 0xDA or m == 0xD9:
            break
        n = (data[i + 2] << 8) | data[i + 3]
        seg = data[i + 4:i + 2 + n]
        i += 2 + n
        if m == 0xE1 and seg[:6] == b"Exif\x00\x00":
            t = seg[6:]
            order = "little" if t[:2] == b"II" else "big"
            ifd = int.from_bytes(t[4:8], order)
            for e in range(int.from_bytes(t[ifd:ifd + 2], order)):
                o = ifd + 2 + 12 * e
                if int.from_bytes(t[o:o + 2], order) == 0x0112:
                    v = int.from_bytes(t[o + 8:o + 10], order)
                    return v if 1 <= v <= 8 else 1
            return 1
    return 1


def decode_rgb_oriented(data: bytes) -> np.ndarray:
    """``decode_rgb`` followed by OpenCV's ApplyExifOrientation - what ``cv2.imread`` / ``imdecode`` return."""
    img = decode_rgb(data)
    o = exif_orientation(data)
    if o == 2:
        return img[:, ::-1]
    if o == 3:
        return img[::-1, ::-1]
    if o == 4:
        return img[::-1]
    t = img.transpose(1, 0, 2)
    if o == 5:
        return t
    if o == 6:
        return t[:, ::-1]
    if o == 7:
        return t[::-1, ::-1]
    if o == 8:
        return t[::-1]
    return img
