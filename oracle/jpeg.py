"""TEST INFRASTRUCTURE ONLY (CPU oracle): numpy restatement of the baseline JPEG encoder behind `cv2.imwrite("*.jpg", bgr_u8)` - the
output stage of the reference's path (test_autoencoder.py:88-93 writes every compressed image with cv2.imwrite; GAN_functions.py:41-50
`save_image`, called at GAN_test.py:390).  SURVEY.md 8 f3 ("uint8/JPEG output stage").

The algorithm lives in a third-party dependency (OpenCV's bundled libjpeg-turbo; the reference pins no version, this image has
opencv-python 4.13).  What is restated, from the published libjpeg design (IJG libjpeg 6b, which libjpeg-turbo reproduces bit for bit):
  * file layout as OpenCV writes it: SOI, APP0/JFIF 1.01, two DQT, SOF0 (8 bit, 3 components, Y 2x2 / Cb 1x1 / Cr 1x1 = 4:2:0), four DHT
    (the Annex K tables), SOS, entropy-coded data, EOI; quality 95 by default, no restart markers, no optimised tables;
  * quantisation tables: Annex K base tables scaled by (200 - 2 q) for q >= 50 (5000 / q below), +50, / 100, clamped to [1, 255];
  * colour conversion: 16-bit fixed point Y / Cb / Cr (jccolor.c), rounding constants ONE_HALF and ONE_HALF - 1;
  * chroma down-sampling: 2x2 box with the alternating bias 1, 2, 1, 2 ... along a row (jcsample.c h2v2_downsample); the right edge
    is replicated before down-sampling, the bottom edge only up to an even row count - below that the last down-sampled row is
    replicated (jcprepct.c); luma edges are replicated;
  * forward DCT: the accurate integer ("islow") 8x8 DCT of jfdctint.c - 13-bit constants, 2 extra bits after the row pass, output
    scaled by 8 - then division by 8 * q with rounding half away from zero (jcdctmgr.c);
  * dummy blocks completing an MCU past the component's last block column / row: AC = 0, DC = DC of the previous block (jccoefct.c);
  * Huffman coding with DC prediction per component, ZRL / EOB, 0xFF byte stuffing, final padding with 1 bits (jchuff.c).

Pinned: `tests/test_oracle_extras.py` compares `encode_bgr` byte for byte with `cv2.imencode(".jpg", img)` (the real library) on
several sizes, including sizes that are not multiples of 16.
"""
from __future__ import annotations

import numpy as np

# ---- Annex K tables --------------------------------------------------------------------------------------------------------------
STD_LUMA_Q = np.array([
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99],
    np.int32)
STD_CHROMA_Q = np.array([
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99],
    np.int32)
ZIGZAG = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63], np.int32)

DC_LUMA_BITS = [0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0]
DC_CHROMA_BITS = [0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0]
DC_VALS = list(range(12))
AC_LUMA_BITS = [0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7D]
AC_LUMA_VALS = [
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xA1, 0x08,
    0x23, 0x42, 0xB1, 0xC1, 0x15, 0x52, 0xD1, 0xF0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0A, 0x16, 0x17, 0x18, 0x19, 0x1A, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2A, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6,
    0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA, 0xE1, 0xE2,
    0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF1, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA]
AC_CHROMA_BITS = [0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77]
AC_CHROMA_VALS = [
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
    0xA1, 0xB1, 0xC1, 0x09, 0x23, 0x33, 0x52, 0xF0, 0x15, 0x62, 0x72, 0xD1, 0x0A, 0x16, 0x24, 0x34, 0xE1, 0x25, 0xF1, 0x17, 0x18, 0x19, 0x1A, 0x26,
    0x27, 0x28, 0x29, 0x2A, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4,
    0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA,
    0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA]


def quant_table(base: np.ndarray, quality: int) -> np.ndarray:
    """jcparam.c jpeg_quality_scaling + jpeg_add_quant_table (force_baseline): natural (row-major) order"""
    q = min(max(int(quality), 1), 100)
    scale = 5000 // q if q < 50 else 200 - 2 * q
    return np.clip((base.astype(np.int64) * scale + 50) // 100, 1, 255).astype(np.int32)


def huff_codes(bits, vals):
    """Annex C: (code, length) per symbol value"""
    code, k, out = 0, 0, {}
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            out[vals[k]] = (code, length)
            code += 1
            k += 1
        code <<= 1
    return out


def header(h: int, w: int, quality: int = 95) -> bytes:
    ql, qc = quant_table(STD_LUMA_Q, quality), quant_table(STD_CHROMA_Q, quality)
    out = bytearray(b"\xFF\xD8")
    out += b"\xFF\xE0" + (16).to_bytes(2, "big") + b"JFIF\x00\x01\x01\x00\x00\x01\x00\x01\x00\x00"
    for idx, q in ((0, ql), (1, qc)):
        out += b"\xFF\xDB" + (67).to_bytes(2, "big") + bytes([idx]) + bytes(int(q[z]) for z in ZIGZAG)
    out += b"\xFF\xC0" + (17).to_bytes(2, "big") + b"\x08" + h.to_bytes(2, "big") + w.to_bytes(2, "big") + b"\x03" \
        + b"\x01\x22\x00" + b"\x02\x11\x01" + b"\x03\x11\x01"
    for tc_th, bits, vals in ((0x00, DC_LUMA_BITS, DC_VALS), (0x10, AC_LUMA_BITS, AC_LUMA_VALS),
                              (0x01, DC_CHROMA_BITS, DC_VALS), (0x11, AC_CHROMA_BITS, AC_CHROMA_VALS)):
        out += b"\xFF\xC4" + (3 + 16 + len(vals)).to_bytes(2, "big") + bytes([tc_th]) + bytes(bits) + bytes(vals)
    out += b"\xFF\xDA" + (12).to_bytes(2, "big") + b"\x03" + b"\x01\x00" + b"\x02\x11" + b"\x03\x11" + b"\x00\x3F\x00"
    return bytes(out)


HEADER_BYTES = len(header(16, 16))

# ---- colour conversion and down-sampling ----------------------------------------------------------------------------------------------
_FIX = lambda x: int(x * 65536 + 0.5)  # noqa: E731


def bgr_to_ycc(img: np.ndarray):
    """jccolor.c rgb_ycc_convert (SCALEBITS 16): uint8 planes"""
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    half = 1 << 15
    y = (_FIX(0.29900) * r + _FIX(0.58700) * g + _FIX(0.11400) * b + half) >> 16
    cb = (-_FIX(0.16874) * r - _FIX(0.33126) * g + _FIX(0.50000) * b + (128 << 16) + half - 1) >> 16
    cr = (_FIX(0.50000) * r - _FIX(0.41869) * g - _FIX(0.08131) * b + (128 << 16) + half - 1) >> 16
    return y.astype(np.int32), cb.astype(np.int32), cr.astype(np.int32)


def _pad_edge(p: np.ndarray, hh: int, ww: int) -> np.ndarray:
    return np.pad(p, ((0, hh - p.shape[0]), (0, ww - p.shape[1])), mode="edge")


def downsample_h2v2(p: np.ndarray) -> np.ndarray:
    """jcsample.c h2v2_downsample: (a + b + c + d + bias) >> 2 with bias 1, 2, 1, 2, ... along the output row"""
    s = p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2]
    bias = 1 + (np.arange(s.shape[1]) & 1)
    return (s + bias[None, :]) >> 2


# ---- DCT and quantisation --------------------------------------------------------------------------------------------------------
def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _dct_1d(d, first_pass: bool):
    """one pass of jfdctint.c over the last axis of d (int64, 8 wide)"""
    c = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137, f1_961=16069,
             f2_053=16819, f2_562=20995, f3_072=25172)
    t0, t7 = d[..., 0] + d[..., 7], d[..., 0] - d[..., 7]
    t1, t6 = d[..., 1] + d[..., 6], d[..., 1] - d[..., 6]
    t2, t5 = d[..., 2] + d[..., 5], d[..., 2] - d[..., 5]
    t3, t4 = d[..., 3] + d[..., 4], d[..., 3] - d[..., 4]
    t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    out = np.empty_like(d)
    n = 13 - 2 if first_pass else 13 + 2
    if first_pass:
        out[..., 0] = (t10 + t11) << 2
        out[..., 4] = (t10 - t11) << 2
    else:
        out[..., 0] = _descale(t10 + t11, 2)
        out[..., 4] = _descale(t10 - t11, 2)
    z1 = (t12 + t13) * c["f0_541"]
    out[..., 2] = _descale(z1 + t13 * c["f0_765"], n)
    out[..., 6] = _descale(z1 - t12 * c["f1_847"], n)
    z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
    z5 = (z3 + z4) * c["f1_175"]
    t4, t5, t6, t7 = t4 * c["f0_298"], t5 * c["f2_053"], t6 * c["f3_072"], t7 * c["f1_501"]
    z1, z2, z3, z4 = -z1 * c["f0_899"], -z2 * c["f2_562"], -z3 * c["f1_961"] + z5, -z4 * c["f0_390"] + z5
    out[..., 7] = _descale(t4 + z1 + z3, n)
    out[..., 5] = _descale(t5 + z2 + z4, n)
    out[..., 3] = _descale(t6 + z2 + z3, n)
    out[..., 1] = _descale(t7 + z1 + z4, n)
    return out


def fdct_islow(blocks: np.ndarray) -> np.ndarray:
    """blocks (..., 8, 8) of level-shifted samples -> DCT coefficients scaled by 8"""
    x = _dct_1d(blocks.astype(np.int64), True)                    # rows
    return np.swapaxes(_dct_1d(np.swapaxes(x, -1, -2), False), -1, -2)   # columns


def quantize(coef: np.ndarray, q: np.ndarray) -> np.ndarray:
    """jcdctmgr.c: divide by (q << 3), rounding half away from zero"""
    qv = (q.reshape(8, 8).astype(np.int64)) << 3
    a = np.abs(coef)
    return (np.sign(coef) * ((a + (qv >> 1)) // qv)).astype(np.int32)


def component_blocks(plane: np.ndarray, q: np.ndarray) -> np.ndarray:
    """plane (multiple of 8 in both axes) -> quantised blocks (rows, cols, 64) in natural order"""
    hb, wb = plane.shape[0] // 8, plane.shape[1] // 8
    b = plane.reshape(hb, 8, wb, 8).transpose(0, 2, 1, 3).astype(np.int64) - 128
    return quantize(fdct_islow(b), q).reshape(hb, wb, 64)


def mcu_blocks(img: np.ndarray, quality: int = 95) -> np.ndarray:
    """(H, W, 3) BGR uint8 -> quantised coefficients (mcu_rows, mcu_cols, 6, 64), natural order within a block, blocks Y00 Y01 Y10 Y11
    Cb Cr; includes libjpeg's dummy-block rule for Y blocks beyond the component's block grid."""
    h, w = img.shape[:2]
    mh, mw = (h + 15) // 16, (w + 15) // 16
    ql, qc = quant_table(STD_LUMA_Q, quality), quant_table(STD_CHROMA_Q, quality)
    y, cb, cr = bgr_to_ycc(img)
    yb = component_blocks(_pad_edge(y, mh * 16, mw * 16), ql)     # (2 mh, 2 mw, 64)
    # chroma: columns are replicated BEFORE down-sampling (h2v2_downsample's expand_right_edge), rows only up to an even count
    # (jcprepct.c pads the input to whole row groups); the rest of the iMCU is filled by replicating the last DOWN-SAMPLED row
    he = h + (h & 1)
    cbb = component_blocks(_pad_edge(downsample_h2v2(_pad_edge(cb, he, mw * 16)), mh * 8, mw * 8), qc)
    crb = component_blocks(_pad_edge(downsample_h2v2(_pad_edge(cr, he, mw * 16)), mh * 8, mw * 8), qc)
    out = np.zeros((mh, mw, 6, 64), np.int32)
    for dy in range(2):
        for dx in range(2):
            out[:, :, 2 * dy + dx] = yb[dy::2, dx::2]
    out[:, :, 4], out[:, :, 5] = cbb, crb
    # dummy blocks (jccoefct.c compress_data): Y blocks past width_in_blocks / height_in_blocks of the component
    ybw, ybh = (w + 7) // 8, (h + 7) // 8
    if ybw % 2:                                                    # right edge: block (., 1) of the last MCU column
        for dy in range(2):
            out[:, mw - 1, 2 * dy + 1, :] = 0
            out[:, mw - 1, 2 * dy + 1, 0] = out[:, mw - 1, 2 * dy, 0]
    if ybh % 2:                                                    # bottom edge: the second block row of the last MCU row
        out[mh - 1, :, 2:4, :] = 0
        out[mh - 1, :, 2, 0] = out[mh - 1, :, 1, 0]
        out[mh - 1, :, 3, 0] = out[mh - 1, :, 1, 0]
    return out


# ---- entropy coding ------------------------------------------------------------------------------------------------------------------
class _Bits:
    def __init__(self):
        self.out, self.acc, self.n = bytearray(), 0, 0

    def put(self, code: int, length: int):
        self.acc = (self.acc << length) | (code & ((1 << length) - 1))
        self.n += length
        while self.n >= 8:
            byte = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(byte)
            if byte == 0xFF:
                self.out.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            self.put(0x7F, 8 - self.n)       # jchuff.c flush_bits: fill the last byte with 1 bits


def _nbits(v: int) -> int:
    return int(abs(v)).bit_length()


def block_bits(block: np.ndarray, pred: int, dc_tab, ac_tab):
    """[(code, length)] of one block (natural order in, zig-zag scan), jchuff.c encode_one_block"""
    out = []
    diff = int(block[0]) - pred
    n = _nbits(diff)
    out.append(dc_tab[n])
    if n:
        out.append(((diff if diff >= 0 else diff - 1) & ((1 << n) - 1), n))
    run = 0
    for k in range(1, 64):
        v = int(block[ZIGZAG[k]])
        if v == 0:
            run += 1
            continue
        while run > 15:
            out.append(ac_tab[0xF0])
            run -= 16
        n = _nbits(v)
        out.append(ac_tab[(run << 4) | n])
        out.append(((v if v >= 0 else v - 1) & ((1 << n) - 1), n))
        run = 0
    if run:
        out.append(ac_tab[0x00])
    return out


def entropy_encode(mcus: np.ndarray) -> bytes:
    dc_l, ac_l = huff_codes(DC_LUMA_BITS, DC_VALS), huff_codes(AC_LUMA_BITS, AC_LUMA_VALS)
    dc_c, ac_c = huff_codes(DC_CHROMA_BITS, DC_VALS), huff_codes(AC_CHROMA_BITS, AC_CHROMA_VALS)
    bits = _Bits()
    pred = [0, 0, 0]
    for row in mcus:
        for mcu in row:
            for b in range(6):
                comp = 0 if b < 4 else b - 3
                for code, length in block_bits(mcu[b], pred[comp], dc_l if comp == 0 else dc_c, ac_l if comp == 0 else ac_c):
                    bits.put(code, length)
                pred[comp] = int(mcu[b][0])
    bits.flush()
    return bytes(bits.out)


def encode_bgr(img: np.ndarray, quality: int = 95) -> bytes:
    """== bytes(cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality])[1]) for (H, W, 3) uint8 BGR"""
    img = np.asarray(img)
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
        raise ValueError("encode_bgr: expected (H, W, 3) uint8")
    h, w = img.shape[:2]
    return header(h, w, quality) + entropy_encode(mcu_blocks(img, quality)) + b"\xFF\xD9"
