"""A tiny baseline-JPEG *writer* for crafting edge-case streams (test tooling, pure Python).

It entropy-codes caller-supplied QUANTISED coefficient blocks (scan order) with the JPEG Annex K
tables, so tests control exactly which symbols appear: full blocks without EOB, ZRL chains, DC
category 0 and 11, maximum AC magnitudes, stuffed FF bytes, fill bytes before RSTn, restart
interval 1, and so on. Small images only.
"""
import numpy as np

# Annex K.3 typical Huffman tables
DC_LUMA_BITS = [0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0]
DC_LUMA_VALS = list(range(12))
DC_CHROMA_BITS = [0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0]
DC_CHROMA_VALS = list(range(12))
AC_LUMA_BITS = [0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7D]
AC_LUMA_VALS = [
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xA1, 0x08, 0x23, 0x42, 0xB1, 0xC1, 0x15, 0x52, 0xD1, 0xF0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0A, 0x16, 0x17, 0x18, 0x19, 0x1A, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3,
    0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3,
    0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA, 0xE1, 0xE2,
    0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF1, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA]
AC_CHROMA_BITS = [0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77]
AC_CHROMA_VALS = [
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xA1, 0xB1, 0xC1, 0x09, 0x23, 0x33, 0x52, 0xF0, 0x15, 0x62, 0x72, 0xD1,
    0x0A, 0x16, 0x24, 0x34, 0xE1, 0x25, 0xF1, 0x17, 0x18, 0x19, 0x1A, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A,
    0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA,
    0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA,
    0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA]

STD_TABLES = {
    (0, 0): (DC_LUMA_BITS, DC_LUMA_VALS), (0, 1): (DC_CHROMA_BITS, DC_CHROMA_VALS),
    (1, 0): (AC_LUMA_BITS, AC_LUMA_VALS), (1, 1): (AC_CHROMA_BITS, AC_CHROMA_VALS),
}


def canonical_codes(bits, vals):
    """symbol -> (code, length)"""
    out, code, k = {}, 0, 0
    for l in range(1, 17):
        for _ in range(bits[l - 1]):
            out[vals[k]] = (code, l)
            code += 1
            k += 1
        code <<= 1
    return out


class BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def put(self, value, nbits):
        if nbits == 0:
            return
        self.acc = (self.acc << nbits) | (value & ((1 << nbits) - 1))
        self.n += nbits
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0x00)   # byte stuffing
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)   # pad with ones

    def raw(self, data):
        assert self.n == 0
        self.out += bytes(data)


def _category(v):
    return 0 if v == 0 else int(abs(int(v))).bit_length()


def _value_bits(v, size):
    return v if v >= 0 else v + (1 << size) - 1


def encode_block(bw, blk, pred, dc_codes, ac_codes, zrl_only_tail=False):
    """blk: 64 quantised coefficients in SCAN order; blk[0] is the absolute DC value."""
    diff = int(blk[0]) - pred
    s = _category(diff)
    code, l = dc_codes[s]
    bw.put(code, l)
    bw.put(_value_bits(diff, s), s)
    run = 0
    last_nz = max([i for i in range(1, 64) if blk[i] != 0], default=0)
    for i in range(1, 64):
        v = int(blk[i])
        if v == 0:
            if i > last_nz and not zrl_only_tail:
                code, l = ac_codes[0x00]
                bw.put(code, l)
                return int(blk[0])
            run += 1
            if run == 16:
                code, l = ac_codes[0xF0]
                bw.put(code, l)
                run = 0
            continue
        s = _category(v)
        code, l = ac_codes[(run << 4) | s]
        bw.put(code, l)
        bw.put(_value_bits(v, s), s)
        run = 0
    if run and not zrl_only_tail:
        code, l = ac_codes[0x00]
        bw.put(code, l)
    return int(blk[0])


def _seg(marker, payload):
    return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + bytes(payload)


def build_jpeg(width, height, sampling, blocks, qtabs, restart_interval=0, fill_before_rst=0,
               tables=None, comp_tables=((0, 0), (1, 1), (1, 1)), with_app0=True, trailing=b"", dqt16=False):
    """sampling: luma (h, v) with 1x1 chroma, or ((hy, vy), (hu, vu), (hv, vv)). blocks: int array [n_blocks, 64] in MCU order and
    SCAN (zig-zag) coefficient order, blocks[:,0] = absolute DC. qtabs: two 64-lists (zig-zag order)
    for luma and chroma. tables: dict like STD_TABLES. Returns the file bytes."""
    tables = tables or STD_TABLES
    if isinstance(sampling[0], (tuple, list)):
        samp = [tuple(x) for x in sampling]          # ((hy, vy), (hu, vu), (hv, vv))
    else:
        samp = [tuple(sampling), (1, 1), (1, 1)]
    h, v = samp[0]
    nblk = [a * b for a, b in samp]
    tot = sum(nblk)
    mcu_w, mcu_h = 8 * max(a for a, _ in samp), 8 * max(b for _, b in samp)
    n_mcu = ((width + mcu_w - 1) // mcu_w) * ((height + mcu_h - 1) // mcu_h)
    blocks = np.asarray(blocks)
    assert blocks.shape == (n_mcu * tot, 64), (blocks.shape, n_mcu * tot)
    out = bytearray(b"\xFF\xD8")
    if with_app0:
        out += _seg(0xE0, b"JFIF\x00\x01\x01\x00\x00\x01\x00\x01\x00\x00")
    if dqt16:   # precision 1: 64 big-endian 16-bit entries (the reference reads them WITHOUT a byte swap, parser.cpp:81-87)
        out += _seg(0xDB, bytes([0x10]) + b"".join(int(q).to_bytes(2, "big") for q in qtabs[0]))
        out += _seg(0xDB, bytes([0x11]) + b"".join(int(q).to_bytes(2, "big") for q in qtabs[1]))
    else:
        out += _seg(0xDB, bytes([0]) + bytes(qtabs[0]))
        out += _seg(0xDB, bytes([1]) + bytes(qtabs[1]))
    sof = bytes([8]) + height.to_bytes(2, "big") + width.to_bytes(2, "big") + bytes([3])
    sof += bytes([1, (samp[0][0] << 4) | samp[0][1], 0, 2, (samp[1][0] << 4) | samp[1][1], 1, 3, (samp[2][0] << 4) | samp[2][1], 1])
    out += _seg(0xC0, sof)
    for (tc, th), (bits, vals) in sorted(tables.items()):
        out += _seg(0xC4, bytes([(tc << 4) | th]) + bytes(bits) + bytes(vals))
    if restart_interval:
        out += _seg(0xDD, restart_interval.to_bytes(2, "big"))
    sos = bytes([3])
    for c in range(3):
        sos += bytes([c + 1, (comp_tables[c][0] << 4) | comp_tables[c][1]])
    sos += bytes([0, 0x3F, 0])
    out += _seg(0xDA, sos)
    codes = {k: canonical_codes(*tv) for k, tv in tables.items()}
    bw = BitWriter()
    pred = [0, 0, 0]
    rst = 0
    bi = 0
    for m in range(n_mcu):
        if restart_interval and m and m % restart_interval == 0:
            bw.flush()
            bw.raw(b"\xFF" * fill_before_rst + bytes([0xFF, 0xD0 + (rst & 7)]))
            rst += 1
            pred = [0, 0, 0]
        for c in range(3):
            for _ in range(nblk[c]):
                dc_codes = codes[(0, comp_tables[c][0])]
                ac_codes = codes[(1, comp_tables[c][1])]
                pred[c] = encode_block(bw, blocks[bi], pred[c], dc_codes, ac_codes)
                bi += 1
    bw.flush()
    out += bw.out
    out += b"\xFF\xD9" + trailing
    return bytes(out)
