"""TEST INFRASTRUCTURE ONLY -- a restatement of the pixel pipeline of jpeg-decoder 0.1.22 (the crate behind
`image 0.23.12`'s `image::open` for the reference's .jpg inputs; un-vendored, restated from its published algorithm).

Why: the reference's golden hashes (lib/tests/diff.rs:163-252) were produced from images decoded by that crate.  Entropy
decoding is lossless -- every conforming decoder recovers the same quantised DCT coefficients -- but the inverse DCT and the
YCbCr -> RGB conversion are not: libjpeg (Pillow) and jpeg-decoder differ by +-1 LSB in a few per cent of the samples, and three of
the nine configurations threshold a JPEG mask at exactly 255 / 0 or feed JPEG guides.  This module decodes

  * baseline and progressive Huffman JPEGs (SOF0 / SOF2) with 8-bit samples, three components, no chroma subsampling
    (all of the reference's imgs/*.jpg that the nine configurations read, except tom.jpg -- 4:2:0 -- which stays a Pillow decode),
  * entropy decoding after ITU T.81 (progressive: spectral selection + successive approximation, Annex G),
  * dequantisation + inverse DCT as jpeg-decoder's idct.rs does it: the 12-bit fixed-point 1-D kernel of stb_image
    (columns first with 2 extra bits kept, `+512 >> 10`; rows with `+65536 + (128 << 17) >> 17`, clamped to 0..255),
  * YCbCr -> RGB as jpeg-decoder's decoder.rs does it in this version: f32 arithmetic with the BT.601 constants
    1.402 / 0.34414 / 0.71414 / 1.772, `+0.5`, truncating cast, clamp.

Pinned by tests/test_oracle_pin.py: with these decodes (tests/golden/ref_imgs_jpegport.npz, made by
tests/golden/make_ref_inputs.py) and the restated rstar order the oracle's hash distances to the reference's constants are
asserted there.  Nothing in the product uses this file.
"""
import struct

import numpy as np

ZIGZAG = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])


class _Bits:
    """MSB-first bit reader over one scan's entropy-coded segment (byte stuffing removed; stops at the next marker)."""

    def __init__(self, data, pos):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def _fill(self):
        d, p = self.d, self.p
        b = d[p] if p < len(d) else 0
        if b == 0xFF:
            nxt = d[p + 1] if p + 1 < len(d) else 0xD9
            if nxt == 0x00:
                p += 2
            else:  # a marker: feed zero bits, do not advance
                b = 0
        else:
            p += 1
        self.p = p
        self.acc = ((self.acc << 8) | b) & 0xFFFFFFFF
        self.n += 8

    def bit(self):
        if self.n == 0:
            self._fill()
        self.n -= 1
        return (self.acc >> self.n) & 1

    def bits(self, k):
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v

    def restart(self):
        """Byte-align and step over an RSTn marker."""
        self.n = 0
        self.acc = 0
        d = self.d
        while self.p + 1 < len(d) and not (d[self.p] == 0xFF and 0xD0 <= d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _huff_table(counts, symbols):
    """(length, code) -> symbol, canonical Huffman codes of T.81 Annex C."""
    table, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            table[(length, code)] = symbols[k]
            code += 1
            k += 1
        code <<= 1
    return table


def _decode_symbol(br, table):
    code = 0
    for length in range(1, 17):
        code = (code << 1) | br.bit()
        s = table.get((length, code))
        if s is not None:
            return s
    raise ValueError("bad Huffman code")


def _extend(v, t):
    return v if v >= (1 << (t - 1)) else v - (1 << t) + 1


def _coefficients(data):
    """Parses the file; returns (width, height, [per component: int32 array (blocks_y, blocks_x, 64) in NATURAL order],
    [per component: 64 quantisation values in natural order], adobe_transform or None)."""
    assert data[0:2] == b"\xff\xd8"
    pos = 2
    qt, hd, ha = {}, {}, {}
    frame = None
    restart_interval = 0
    adobe = None
    coefs = None
    while pos < len(data):
        assert data[pos] == 0xFF, hex(pos)
        m = data[pos + 1]
        if m == 0xFF:
            pos += 1
            continue
        if m == 0xD9:
            break
        L = struct.unpack(">H", data[pos + 2:pos + 4])[0]
        seg = data[pos + 4:pos + 2 + L]
        if m == 0xDB:
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                i += 1
                t = np.zeros(64, np.int32)
                for k in range(64):
                    if pq:
                        t[ZIGZAG[k]] = struct.unpack(">H", seg[i:i + 2])[0]
                        i += 2
                    else:
                        t[ZIGZAG[k]] = seg[i]
                        i += 1
                qt[tq] = t
        elif m == 0xC4:
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                counts = list(seg[i + 1:i + 17])
                n = sum(counts)
                syms = list(seg[i + 17:i + 17 + n])
                (ha if tc else hd)[th] = _huff_table(counts, syms)
                i += 17 + n
        elif m in (0xC0, 0xC1, 0xC2):
            p, h, w, nc = struct.unpack(">BHHB", seg[:6])
            assert p == 8
            comps = []
            for c in range(nc):
                cid, hv, tq = seg[6 + 3 * c], seg[7 + 3 * c], seg[8 + 3 * c]
                assert hv == 0x11, "only unsubsampled images are restated"
                comps.append(dict(id=cid, tq=tq))
            frame = dict(w=w, h=h, comps=comps, progressive=(m == 0xC2))
            bx, by = (w + 7) // 8, (h + 7) // 8
            coefs = [np.zeros((by, bx, 64), np.int32) for _ in comps]
        elif m == 0xDD:
            restart_interval = struct.unpack(">H", seg[:2])[0]
        elif m == 0xEE and seg[:5] == b"Adobe":
            adobe = seg[11]
        elif m == 0xDA:
            ns = seg[0]
            sel = []
            for c in range(ns):
                cid, tt = seg[1 + 2 * c], seg[2 + 2 * c]
                ci = [i for i, cc in enumerate(frame["comps"]) if cc["id"] == cid][0]
                sel.append((ci, tt >> 4, tt & 15))
            ss, se, ahl = seg[1 + 2 * ns], seg[2 + 2 * ns], seg[3 + 2 * ns]
            ah, al = ahl >> 4, ahl & 15
            if not frame["progressive"]:
                ss, se, ah, al = 0, 63, 0, 0
            pos = _scan(data, pos + 2 + L, frame, coefs, sel, ss, se, ah, al, hd, ha, restart_interval)
            continue
        pos += 2 + L
    quant = [qt[c["tq"]] for c in frame["comps"]]
    return frame["w"], frame["h"], coefs, quant, adobe


def _scan(data, pos, frame, coefs, sel, ss, se, ah, al, hd, ha, restart_interval):
    br = _Bits(data, pos)
    bx, by = (frame["w"] + 7) // 8, (frame["h"] + 7) // 8
    pred = [0] * len(frame["comps"])
    eobrun = 0
    p1, m1 = 1 << al, -1 << al
    n_units = bx * by
    for unit in range(n_units):
        if restart_interval and unit and unit % restart_interval == 0:
            br.restart()
            pred = [0] * len(frame["comps"])
            eobrun = 0
        y, x = divmod(unit, bx)
        for ci, td, ta in sel:
            blk = coefs[ci][y, x]
            if not frame["progressive"]:
                t = _decode_symbol(br, hd[td])
                pred[ci] += _extend(br.bits(t), t) if t else 0
                blk[0] = pred[ci]
                k = 1
                tab = ha[ta]
                while k < 64:
                    rs = _decode_symbol(br, tab)
                    r, s = rs >> 4, rs & 15
                    if s == 0:
                        if r != 15:
                            break
                        k += 16
                        continue
                    k += r
                    blk[ZIGZAG[k]] = _extend(br.bits(s), s)
                    k += 1
                continue
            if ss == 0:  # DC scan
                if ah == 0:
                    t = _decode_symbol(br, hd[td])
                    pred[ci] += _extend(br.bits(t), t) if t else 0
                    blk[0] = pred[ci] * (1 << al)
                elif br.bit():
                    blk[0] |= p1
                continue
            tab = ha[ta]
            if ah == 0:  # AC first scan (G.1.2.2)
                if eobrun > 0:
                    eobrun -= 1
                    continue
                k = ss
                while k <= se:
                    rs = _decode_symbol(br, tab)
                    r, s = rs >> 4, rs & 15
                    if s == 0:
                        if r < 15:
                            eobrun = (1 << r) - 1
                            if r:
                                eobrun += br.bits(r)
                            break
                        k += 16
                        continue
                    k += r
                    blk[ZIGZAG[k]] = _extend(br.bits(s), s) * (1 << al)
                    k += 1
                continue
            # AC refinement scan (G.1.2.3)
            k = ss
            if eobrun == 0:
                while k <= se:
                    rs = _decode_symbol(br, tab)
                    r, s = rs >> 4, rs & 15
                    if s:
                        s = p1 if br.bit() else m1
                    elif r != 15:
                        eobrun = 1 << r
                        if r:
                            eobrun += br.bits(r)
                        break
                    while k <= se:
                        idx = ZIGZAG[k]
                        v = blk[idx]
                        if v != 0:
                            if br.bit() and (v & p1) == 0:
                                blk[idx] = v + (p1 if v >= 0 else m1)
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
                    idx = ZIGZAG[k]
                    v = blk[idx]
                    if v != 0 and br.bit() and (v & p1) == 0:
                        blk[idx] = v + (p1 if v >= 0 else m1)
                    k += 1
                eobrun -= 1
    # position of the next marker
    p = br.p
    while p + 1 < len(data) and not (data[p] == 0xFF and data[p + 1] != 0x00 and not (0xD0 <= data[p + 1] <= 0xD7)):
        p += 1
    return p


def _f2f(x):
    return int(x * 4096.0 + 0.5)


def _idct_1d(s0, s1, s2, s3, s4, s5, s6, s7):
    """The 1-D kernel of stb_image's stbi__idct_block as ported by jpeg-decoder's idct.rs; returns (x0..x3, t0..t3)."""
    p2, p3 = s2, s6
    p1 = (p2 + p3) * _f2f(0.5411961)
    t2 = p1 + p3 * _f2f(-1.847759065)
    t3 = p1 + p2 * _f2f(0.765366865)
    p2, p3 = s0, s4
    t0 = (p2 + p3) * 4096
    t1 = (p2 - p3) * 4096
    x0, x3, x1, x2 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    t0, t1, t2, t3 = s7, s5, s3, s1
    p3, p4, p1, p2 = t0 + t2, t1 + t3, t0 + t3, t1 + t2
    p5 = (p3 + p4) * _f2f(1.175875602)
    t0 = t0 * _f2f(0.298631336)
    t1 = t1 * _f2f(2.053119869)
    t2 = t2 * _f2f(3.072711026)
    t3 = t3 * _f2f(1.501321110)
    p1 = p5 + p1 * _f2f(-0.899976223)
    p2 = p5 + p2 * _f2f(-2.562915447)
    p3 = p3 * _f2f(-1.961570560)
    p4 = p4 * _f2f(-0.390180644)
    t3 = t3 + p1 + p4
    t2 = t2 + p2 + p3
    t1 = t1 + p2 + p4
    t0 = t0 + p1 + p3
    return x0, x1, x2, x3, t0, t1, t2, t3


def _idct_blocks(coef, quant):
    """coef (by, bx, 64) natural order -> samples (by*8, bx*8) uint8."""
    by, bx, _ = coef.shape
    d = (coef.astype(np.int64) * quant.astype(np.int64)).reshape(by, bx, 8, 8)  # [row v][column u]
    # columns: for each column i the eight rows d[0..7][i].  (stb's all-zero-AC shortcut, dcterm << 2, gives the same values.)
    s = [d[:, :, r, :] for r in range(8)]
    x0, x1, x2, x3, t0, t1, t2, t3 = _idct_1d(*s)
    x0, x1, x2, x3 = x0 + 512, x1 + 512, x2 + 512, x3 + 512
    tmp = np.empty_like(d)
    tmp[:, :, 0, :] = (x0 + t3) >> 10
    tmp[:, :, 7, :] = (x0 - t3) >> 10
    tmp[:, :, 1, :] = (x1 + t2) >> 10
    tmp[:, :, 6, :] = (x1 - t2) >> 10
    tmp[:, :, 2, :] = (x2 + t1) >> 10
    tmp[:, :, 5, :] = (x2 - t1) >> 10
    tmp[:, :, 3, :] = (x3 + t0) >> 10
    tmp[:, :, 4, :] = (x3 - t0) >> 10
    # rows
    s = [tmp[:, :, :, c] for c in range(8)]
    x0, x1, x2, x3, t0, t1, t2, t3 = _idct_1d(*s)
    bias = 65536 + (128 << 17)
    x0, x1, x2, x3 = x0 + bias, x1 + bias, x2 + bias, x3 + bias
    out = np.empty_like(d)
    out[:, :, :, 0] = (x0 + t3) >> 17
    out[:, :, :, 7] = (x0 - t3) >> 17
    out[:, :, :, 1] = (x1 + t2) >> 17
    out[:, :, :, 6] = (x1 - t2) >> 17
    out[:, :, :, 2] = (x2 + t1) >> 17
    out[:, :, :, 5] = (x2 - t1) >> 17
    out[:, :, :, 3] = (x3 + t0) >> 17
    out[:, :, :, 4] = (x3 - t0) >> 17
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)


def _ycbcr_to_rgb(y, cb, cr):
    y = y.astype(np.float32)
    cb = cb.astype(np.float32) - np.float32(128.0)
    cr = cr.astype(np.float32) - np.float32(128.0)
    r = y + np.float32(1.40200) * cr
    g = y - np.float32(0.34414) * cb - np.float32(0.71414) * cr
    b = y + np.float32(1.77200) * cb
    half = np.float32(0.5)
    return [np.clip((c + half).astype(np.int32), 0, 255).astype(np.uint8) for c in (r, g, b)]  # `as i32` truncates toward zero


def decode(path_or_bytes):
    """RGB uint8 array (h, w, 3) as jpeg-decoder 0.1.22 would deliver it."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    w, h, coefs, quant, adobe = _coefficients(data)
    planes = [_idct_blocks(c, q)[:h, :w] for c, q in zip(coefs, quant)]
    assert len(planes) == 3
    if adobe == 0:  # Adobe marker says RGB
        return np.stack(planes, axis=-1)
    return np.stack(_ycbcr_to_rgb(*planes), axis=-1)
