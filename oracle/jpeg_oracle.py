"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, integer arithmetic) of the baseline JPEG round trip that the
reference's JPEG stage performs through libjpeg-turbo.

Follows ``models/utils/turbo_jpeg_compression.py:17-77``: ``TurboJPEG.encode(img_np, quality=q)`` with PyTurboJPEG
1.7.7's defaults (``pixel_format=TJPF_BGR`` although the array is RGB, ``jpeg_subsample=TJSAMP_422``, baseline
sequential Huffman with the Annex K tables, no restart markers) followed by ``TurboJPEG.decode`` (BGR out, ISLOW IDCT,
fancy up-sampling).  libjpeg-turbo itself is absent from /root/reference (un-vendored system library, ``setup.sh:25``);
its algorithm is restated here from the published libjpeg sources (jccolor.c, jcsample.c h2v1_downsample,
jfdctint.c, jcdctmgr.c quantize, jchuff.c encode_one_block, jdhuff.c, jidctint.c, jdsample.c h2v1_fancy_upsample,
jdcolor.c, jcparam.c quality scaling / std tables, jcmarker.c header order).

PINNED: ``tests/test_jpeg_oracle.py`` checks ``encode`` byte-for-byte and ``decode_coefficients`` pixel-for-pixel
against OpenCV's bundled libjpeg-turbo 3.1.2 (``cv2.imencode`` / ``cv2.imdecode``; the same library PyTurboJPEG wraps)
when cv2 is importable, and against committed fixtures made by it (``tests/golden/jpeg_*.npz``) otherwise.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
import numpy as np

# ---- Annex K tables (jcparam.c std_luminance_quant_tbl / std_chrominance_quant_tbl, natural order) ----
STD_LUM_Q = np.array([
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
    14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int64)
STD_CHR_Q = np.array([
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
    47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99], dtype=np.int64)
# zig-zag position -> natural (row-major) index (jutils.c jpeg_natural_order)
ZIGZAG = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,
    7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
    39, 46, 53, 60, 61, 54, 47, 55, 62, 63], dtype=np.int64)
DC_LUM_BITS = [0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0]
DC_CHR_BITS = [0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0]
DC_VALS = list(range(12))
AC_LUM_BITS = [0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d]
AC_LUM_VALS = [
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa]
AC_CHR_BITS = [0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77]
AC_CHR_VALS = [
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa]

# jfdctint.c / jidctint.c fixed-point constants (CONST_BITS = 13)
CONST_BITS, PASS1_BITS = 13, 2
F_0_298, F_0_390, F_0_541, F_0_765, F_0_899, F_1_175 = 2446, 3196, 4433, 6270, 7373, 9633
F_1_501, F_1_847, F_1_961, F_2_053, F_2_562, F_3_072 = 12299, 15137, 16069, 16819, 20995, 25172


def quant_tables(quality):
    """jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline=TRUE): two [64] tables, natural order."""
    q = min(max(int(quality), 1), 100)
    scale = 5000 // q if q < 50 else 200 - 2 * q
    out = []
    for base in (STD_LUM_Q, STD_CHR_Q):
        t = (base * scale + 50) // 100
        out.append(np.clip(t, 1, 255))
    return out


def huff_codes(bits, vals):
    """jchuff.c jpeg_make_c_derived_tbl: symbol -> (code, length)."""
    code, k, table = 0, 0, {}
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            table[vals[k]] = (code, length)
            code += 1
            k += 1
        code <<= 1
    return table


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _fdct_pass(d, first):
    """One 1-D pass of jpeg_fdct_islow along the last axis of d[..., 8] (int64)."""
    t0, t7 = d[..., 0] + d[..., 7], d[..., 0] - d[..., 7]
    t1, t6 = d[..., 1] + d[..., 6], d[..., 1] - d[..., 6]
    t2, t5 = d[..., 2] + d[..., 5], d[..., 2] - d[..., 5]
    t3, t4 = d[..., 3] + d[..., 4], d[..., 3] - d[..., 4]
    t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    o = np.empty_like(d)
    if first:
        o[..., 0] = (t10 + t11) << PASS1_BITS
        o[..., 4] = (t10 - t11) << PASS1_BITS
        sh = CONST_BITS - PASS1_BITS
    else:
        o[..., 0] = _descale(t10 + t11, PASS1_BITS)
        o[..., 4] = _descale(t10 - t11, PASS1_BITS)
        sh = CONST_BITS + PASS1_BITS
    z1 = (t12 + t13) * F_0_541
    o[..., 2] = _descale(z1 + t13 * F_0_765, sh)
    o[..., 6] = _descale(z1 - t12 * F_1_847, sh)
    z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
    z5 = (z3 + z4) * F_1_175
    t4, t5, t6, t7 = t4 * F_0_298, t5 * F_2_053, t6 * F_3_072, t7 * F_1_501
    z1, z2, z3, z4 = -z1 * F_0_899, -z2 * F_2_562, -z3 * F_1_961 + z5, -z4 * F_0_390 + z5
    o[..., 7] = _descale(t4 + z1 + z3, sh)
    o[..., 5] = _descale(t5 + z2 + z4, sh)
    o[..., 3] = _descale(t6 + z2 + z3, sh)
    o[..., 1] = _descale(t7 + z1 + z4, sh)
    return o


def fdct_quant(plane, qt):
    """[H, W] uint8 plane -> [H/8, W/8, 64] quantised coefficients (natural order), jfdctint.c + jcdctmgr.c."""
    H, W = plane.shape
    blk = plane.astype(np.int64).reshape(H // 8, 8, W // 8, 8).transpose(0, 2, 1, 3) - 128  # [by, bx, row, col]
    d = _fdct_pass(blk, True)                              # rows
    d = _fdct_pass(d.swapaxes(-1, -2), False).swapaxes(-1, -2)  # columns
    d = d.reshape(H // 8, W // 8, 64)
    div = qt.astype(np.int64) << 3                          # the FDCT output is scaled up by 8
    a = np.abs(d)
    q = (a + (div >> 1)) // div
    return np.where(d < 0, -q, q)


def _idct_pass(c, first):
    """One 1-D pass of jpeg_idct_islow along the last axis of c[..., 8] (int64)."""
    z2, z3 = c[..., 2], c[..., 6]
    z1 = (z2 + z3) * F_0_541
    t2 = z1 - z3 * F_1_847
    t3 = z1 + z2 * F_0_765
    t0 = (c[..., 0] + c[..., 4]) << CONST_BITS
    t1 = (c[..., 0] - c[..., 4]) << CONST_BITS
    t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    t0, t1, t2, t3 = c[..., 7], c[..., 5], c[..., 3], c[..., 1]
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * F_1_175
    t0, t1, t2, t3 = t0 * F_0_298, t1 * F_2_053, t2 * F_3_072, t3 * F_1_501
    z1, z2, z3, z4 = -z1 * F_0_899, -z2 * F_2_562, -z3 * F_1_961 + z5, -z4 * F_0_390 + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    sh = CONST_BITS - PASS1_BITS if first else CONST_BITS + PASS1_BITS + 3
    o = np.empty_like(c)
    o[..., 0], o[..., 7] = _descale(t10 + t3, sh), _descale(t10 - t3, sh)
    o[..., 1], o[..., 6] = _descale(t11 + t2, sh), _descale(t11 - t2, sh)
    o[..., 2], o[..., 5] = _descale(t12 + t1, sh), _descale(t12 - t1, sh)
    o[..., 3], o[..., 4] = _descale(t13 + t0, sh), _descale(t13 - t0, sh)
    return o


def dequant_idct(coef, qt):
    """[H/8, W/8, 64] coefficients -> [H, W] uint8 plane, jidctint.c jpeg_idct_islow + range limit."""
    by, bx, _ = coef.shape
    c = (coef.astype(np.int64) * qt.astype(np.int64)).reshape(by, bx, 8, 8)
    w = _idct_pass(c.swapaxes(-1, -2), True).swapaxes(-1, -2)  # columns first
    o = _idct_pass(w, False)                                    # then rows
    o = np.clip(o + 128, 0, 255)
    return o.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8).astype(np.uint8)


def color_forward(img):
    """[H, W, 3] uint8 in libjpeg's (B, G, R) reading of the array -> Y [H, W], Cb, Cr [H, W/2] uint8.
    jccolor.c rgb_ycc_convert (16-bit fixed point) + jcsample.c h2v1_downsample (alternating 0/1 bias)."""
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    half = 1 << 15
    y = (19595 * r + 38470 * g + 7471 * b + half) >> 16
    cb = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + half - 1) >> 16
    cr = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + half - 1) >> 16
    bias = np.arange(cb.shape[1] // 2, dtype=np.int64) & 1
    cbs = (cb[:, 0::2] + cb[:, 1::2] + bias) >> 1
    crs = (cr[:, 0::2] + cr[:, 1::2] + bias) >> 1
    return y.astype(np.uint8), cbs.astype(np.uint8), crs.astype(np.uint8)


def color_inverse(y, cbs, crs):
    """Y [H, W], Cb, Cr [H, W/2] uint8 -> [H, W, 3] uint8 in (B, G, R) order.
    jdsample.c h2v1_fancy_upsample (3/4 - 1/4 triangle filter) + jdcolor.c ycc_rgb_convert."""
    def up(c):
        c = c.astype(np.int64)
        left = np.concatenate([c[:, :1], c[:, :-1]], axis=1)
        right = np.concatenate([c[:, 1:], c[:, -1:]], axis=1)
        o = np.empty((c.shape[0], c.shape[1] * 2), dtype=np.int64)
        o[:, 0::2] = (3 * c + left + 1) >> 2
        o[:, 1::2] = (3 * c + right + 2) >> 2
        o[:, 0] = c[:, 0]
        o[:, -1] = c[:, -1]
        return o
    yy = y.astype(np.int64)
    cb, cr = up(cbs) - 128, up(crs) - 128
    half = 1 << 15
    r = yy + ((91881 * cr + half) >> 16)
    b = yy + ((116130 * cb + half) >> 16)
    g = yy + ((-22554 * cb - 46802 * cr + half) >> 16)
    return np.clip(np.stack([b, g, r], axis=-1), 0, 255).astype(np.uint8)


def coefficients(img, quality):
    """[H, W, 3] uint8 (H % 8 == 0, W % 16 == 0) -> (Y, Cb, Cr) coefficient arrays [H/8, w/8, 64] + tables."""
    H, W, _ = img.shape
    if H % 8 or W % 16:
        raise ValueError("jpeg oracle: H must be a multiple of 8 and W of 16 (no edge padding in this restatement)")
    ql, qc = quant_tables(quality)
    y, cb, cr = color_forward(img)
    return (fdct_quant(y, ql), fdct_quant(cb, qc), fdct_quant(cr, qc)), (ql, qc)


def decode_coefficients(coefs, tables):
    (cy, ccb, ccr), (ql, qc) = coefs, tables
    return color_inverse(dequant_idct(cy, ql), dequant_idct(ccb, qc), dequant_idct(ccr, qc))


def roundtrip(img, quality):
    """What encode -> decode returns, without the entropy coder in between: [H, W, 3] uint8."""
    coefs, tables = coefficients(img, quality)
    return decode_coefficients(coefs, tables)


# ---- entropy coder + markers (jchuff.c, jcmarker.c) ----
class _BitWriter:
    def __init__(self, stuff=True):
        self.out = bytearray()
        self.acc = 0
        self.n = 0
        self.stuff = stuff
        self.nbits = 0

    def put(self, code, length):
        self.acc = (self.acc << length) | code
        self.n += length
        self.nbits += length
        while self.n >= 8:
            byte = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(byte)
            if byte == 0xFF and self.stuff:
                self.out.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)


def _encode_block(bw, blk, last_dc, dc_tab, ac_tab):
    """jchuff.c encode_one_block; blk is [64] in natural order."""
    diff = int(blk[0]) - last_dc
    t, t2 = (diff, diff) if diff >= 0 else (-diff, diff - 1)
    nb = t.bit_length()
    bw.put(*dc_tab[nb])
    if nb:
        bw.put(t2 & ((1 << nb) - 1), nb)
    run = 0
    zz = blk[ZIGZAG]
    for k in range(1, 64):
        v = int(zz[k])
        if v == 0:
            run += 1
            continue
        while run > 15:
            bw.put(*ac_tab[0xF0])
            run -= 16
        t, t2 = (v, v) if v >= 0 else (-v, v - 1)
        nb = t.bit_length()
        bw.put(*ac_tab[(run << 4) + nb])
        bw.put(t2 & ((1 << nb) - 1), nb)
        run = 0
    if run:
        bw.put(*ac_tab[0])
    return int(blk[0])


def header(H, W, quality):
    """SOI, APP0 (JFIF 1.01, density 1:1), DQT x2, SOF0, DHT x4, SOS -- the order jcmarker.c writes them in."""
    ql, qc = quant_tables(quality)
    o = bytearray(b"\xff\xd8")
    o += b"\xff\xe0" + (16).to_bytes(2, "big") + b"JFIF\x00\x01\x01\x00\x00\x01\x00\x01\x00\x00"
    for i, t in enumerate((ql, qc)):
        o += b"\xff\xdb" + (67).to_bytes(2, "big") + bytes([i]) + bytes(int(t[z]) for z in ZIGZAG)
    o += b"\xff\xc0" + (17).to_bytes(2, "big") + b"\x08" + H.to_bytes(2, "big") + W.to_bytes(2, "big") + b"\x03"
    o += bytes([1, 0x21, 0, 2, 0x11, 1, 3, 0x11, 1])
    for tc_th, bits, vals in ((0x00, DC_LUM_BITS, DC_VALS), (0x10, AC_LUM_BITS, AC_LUM_VALS),
                              (0x01, DC_CHR_BITS, DC_VALS), (0x11, AC_CHR_BITS, AC_CHR_VALS)):
        o += b"\xff\xc4" + (3 + 16 + len(vals)).to_bytes(2, "big") + bytes([tc_th]) + bytes(bits) + bytes(vals)
    o += b"\xff\xda" + (12).to_bytes(2, "big") + b"\x03" + bytes([1, 0x00, 2, 0x11, 3, 0x11]) + b"\x00\x3f\x00"
    return bytes(o)


def scan_bits(coefs):
    """The scan as the device kernels leave it (hyres_jpeg_forward's scan_words / scan_bits): the raw bit string,
    no byte stuffing, no padding -> (big-endian uint32 words, number of bits)."""
    bw = _scan(coefs, stuff=False)
    nbits = bw.nbits
    if bw.n:
        bw.out.append((bw.acc << (8 - bw.n)) & 0xFF)
    raw = bytes(bw.out) + b"\x00" * (-len(bw.out) % 4)
    return np.frombuffer(raw, dtype=">u4").astype(np.uint32), nbits


def entropy_segment(coefs):
    """The interleaved 4:2:2 scan (MCU = Y Y Cb Cr), byte-stuffed, final byte padded with one bits."""
    bw = _scan(coefs, stuff=True)
    bw.flush()
    return bytes(bw.out)


def _scan(coefs, stuff):
    cy, ccb, ccr = coefs
    dcl, acl = huff_codes(DC_LUM_BITS, DC_VALS), huff_codes(AC_LUM_BITS, AC_LUM_VALS)
    dcc, acc = huff_codes(DC_CHR_BITS, DC_VALS), huff_codes(AC_CHR_BITS, AC_CHR_VALS)
    bw = _BitWriter(stuff)
    last = [0, 0, 0]
    for my in range(cy.shape[0]):
        for mx in range(ccb.shape[1]):
            last[0] = _encode_block(bw, cy[my, 2 * mx], last[0], dcl, acl)
            last[0] = _encode_block(bw, cy[my, 2 * mx + 1], last[0], dcl, acl)
            last[1] = _encode_block(bw, ccb[my, mx], last[1], dcc, acc)
            last[2] = _encode_block(bw, ccr[my, mx], last[2], dcc, acc)
    return bw


def encode(img, quality):
    """[H, W, 3] uint8 -> the JPEG file libjpeg-turbo writes for it (cv2.imencode / TurboJPEG.encode, 4:2:2)."""
    coefs, _ = coefficients(img, quality)
    return header(img.shape[0], img.shape[1], quality) + entropy_segment(coefs) + b"\xff\xd9"


def stage_forward(x, quality):
    """``TurboJPEGCompression.forward`` (models/utils/turbo_jpeg_compression.py:62-77) on a [B,3,H,W] float array in
    [0,1]: returns (decoded [B,3,H,W] float32 = u8 / 255, bpp, per-image byte counts)."""
    x = np.asarray(x, dtype=np.float32)
    u8 = (np.clip(x, 0, 1).transpose(0, 2, 3, 1) * np.float32(255)).astype(np.uint8)  # .byte() truncates
    dec, sizes = [], []
    for img in u8:
        coefs, tables = coefficients(img, quality)
        sizes.append(len(header(img.shape[0], img.shape[1], quality)) + len(entropy_segment(coefs)) + 2)
        dec.append(decode_coefficients(coefs, tables))
    d = np.stack(dec).transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255.0)
    B, _, H, W = x.shape
    return d, 8.0 * sum(sizes) / (B * H * W), sizes
