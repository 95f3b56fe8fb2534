"""TEST INFRASTRUCTURE ONLY (oracle): baseline JPEG decoder in numpy / pure Python.

Restates what `tf.image.decode_image` (utils/tfdata.py:426-484) does for baseline JPEG through
libjpeg(-turbo) with its defaults - dct_method ISLOW, fancy upsampling on - so that the split decoder of
the engine (host Huffman in csrc/host_io.cc, device IDCT / upsampling / colour in csrc/jpeg.cu) has a
bit-exact checker.  The algorithm lives in a third-party dependency that is absent from /root/reference
(libjpeg-turbo, linked by TensorFlow); its published algorithm is restated here:
  * entropy decoding: ITU-T T.81 Annex F (sequential DCT, Huffman), restart markers (E.1.4)
  * inverse DCT: jidctint.c `jpeg_idct_islow` (CONST_BITS 13, PASS1_BITS 2, range limit to [0, 255])
  * chroma upsampling: jdsample.c `h2v1_fancy_upsample` / `h2v2_fancy_upsample` (triangle filter)
  * colour: jdcolor.c YCbCr -> RGB with 16-bit fixed-point tables
Pinned in tests/test_jpeg.py against PIL (libjpeg-turbo) on the reference fixture's images and on
synthetic 4:4:4 / 4:2:2 / 4:2:0 / grey / restart-interval JPEGs: bit exact.
"""
import numpy as np

ZIGZAG = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
    21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60,
    61, 54, 47, 55, 62, 63], np.int32)


class JpegError(ValueError):
  pass


def _u16(b, p):
  return (b[p] << 8) | b[p + 1]


def parse_headers(data):
  """Returns dict(width, height, comps=[(id, h, v, tq, td, ta)], qt={id: [64] natural order},
  dc/ac huffman tables {id: (counts[16], symbols)}, restart_interval, scan_offset)."""
  b = data
  if len(b) < 4 or b[0] != 0xFF or b[1] != 0xD8:
    raise JpegError('not a JPEG (no SOI)')
  p = 2
  info = {'qt': {}, 'dc': {}, 'ac': {}, 'restart_interval': 0}
  while True:
    if p + 4 > len(b):
      raise JpegError('truncated before SOS')
    if b[p] != 0xFF:
      raise JpegError('marker expected at %d' % p)
    while b[p + 1] == 0xFF:
      p += 1
    m = b[p + 1]
    p += 2
    if m == 0xD8 or (0xD0 <= m <= 0xD7) or m == 0x01:
      continue
    n = _u16(b, p)
    seg = b[p + 2:p + n]
    if m == 0xDB:    # DQT
      q = 0
      while q < len(seg):
        pq, tq = seg[q] >> 4, seg[q] & 15
        q += 1
        tbl = np.zeros(64, np.int32)
        for i in range(64):
          if pq:
            tbl[ZIGZAG[i]] = (seg[q] << 8) | seg[q + 1]
            q += 2
          else:
            tbl[ZIGZAG[i]] = seg[q]
            q += 1
        info['qt'][tq] = tbl
    elif m in (0xC0, 0xC1):   # SOF0 / SOF1: baseline / extended sequential, Huffman
      if seg[0] != 8:
        raise JpegError('only 8-bit precision is supported')
      info['height'], info['width'] = _u16(seg, 1), _u16(seg, 3)
      nc = seg[5]
      info['comps'] = [[seg[6 + 3 * i], seg[7 + 3 * i] >> 4, seg[7 + 3 * i] & 15, seg[8 + 3 * i], 0, 0]
                       for i in range(nc)]
    elif m in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
      raise JpegError('unsupported JPEG process (SOF marker 0x%02X): only baseline sequential' % m)
    elif m == 0xC4:  # DHT
      q = 0
      while q < len(seg):
        tc, th = seg[q] >> 4, seg[q] & 15
        counts = list(seg[q + 1:q + 17])
        total = sum(counts)
        symbols = list(seg[q + 17:q + 17 + total])
        info['ac' if tc else 'dc'][th] = (counts, symbols)
        q += 17 + total
    elif m == 0xDD:  # DRI
      info['restart_interval'] = _u16(seg, 0)
    elif m == 0xDA:  # SOS
      ns = seg[0]
      if 'comps' not in info or ns != len(info['comps']):
        raise JpegError('only single-scan (interleaved) JPEGs are supported')
      for i in range(ns):
        cid, t = seg[1 + 2 * i], seg[2 + 2 * i]
        for c in info['comps']:
          if c[0] == cid:
            c[4], c[5] = t >> 4, t & 15
      info['scan_offset'] = p + n
      return info
    p += n


def _build_lookup(counts, symbols):
  """code length / value tables per T.81 C.2 / F.2.2.3 as a {(length, code): symbol} dict."""
  table, code, k = {}, 0, 0
  for length in range(1, 17):
    for _ in range(counts[length - 1]):
      table[(length, code)] = symbols[k]
      code += 1
      k += 1
    code <<= 1
  return table


class _Bits(object):

  def __init__(self, data, pos):
    self.d, self.p, self.acc, self.n = data, pos, 0, 0

  def bit(self):
    if self.n == 0:
      if self.p >= len(self.d):
        byte = 0
      else:
        byte = self.d[self.p]
        self.p += 1
        if byte == 0xFF:
          nxt = self.d[self.p] if self.p < len(self.d) else 0
          if nxt == 0:
            self.p += 1
          else:       # a marker inside the scan: feed zeros (libjpeg behaviour), do not consume it
            self.p -= 1
            byte = 0
      self.acc, self.n = byte, 8
    self.n -= 1
    return (self.acc >> self.n) & 1

  def bits(self, k):
    v = 0
    for _ in range(k):
      v = (v << 1) | self.bit()
    return v

  def decode(self, table):
    code = 0
    for length in range(1, 17):
      code = (code << 1) | self.bit()
      s = table.get((length, code))
      if s is not None:
        return s
    raise JpegError('bad Huffman code')

  def restart(self):
    """Byte-align and skip the RSTn marker."""
    self.n = 0
    while self.p + 1 < len(self.d) and not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
      self.p += 1
    self.p += 2


def _extend(v, t):
  return v - (1 << t) + 1 if t and v < (1 << (t - 1)) else v


def decode_coefficients(data):
  """Entropy decoding (T.81 F.2).  Returns (info, [per component int16 [blocks_h, blocks_w, 64] natural order])."""
  info = parse_headers(data)
  comps = info['comps']
  hmax, vmax = max(c[1] for c in comps), max(c[2] for c in comps)
  mcux = -(-info['width'] // (8 * hmax))
  mcuy = -(-info['height'] // (8 * vmax))
  coefs = [np.zeros((mcuy * c[2], mcux * c[1], 64), np.int16) for c in comps]
  dc_t = {k: _build_lookup(*v) for k, v in info['dc'].items()}
  ac_t = {k: _build_lookup(*v) for k, v in info['ac'].items()}
  bits = _Bits(data, info['scan_offset'])
  pred = [0] * len(comps)
  ri, count = info['restart_interval'], 0
  for my in range(mcuy):
    for mx in range(mcux):
      if ri and count and count % ri == 0:
        bits.restart()
        pred = [0] * len(comps)
      count += 1
      for ci, c in enumerate(comps):
        for by in range(c[2]):
          for bx in range(c[1]):
            blk = coefs[ci][my * c[2] + by, mx * c[1] + bx]
            t = bits.decode(dc_t[c[4]])
            pred[ci] += _extend(bits.bits(t), t)
            blk[0] = pred[ci]
            k = 1
            while k < 64:
              rs = bits.decode(ac_t[c[5]])
              r, s = rs >> 4, rs & 15
              if s == 0:
                if r != 15:
                  break
                k += 16
                continue
              k += r
              if k > 63:
                raise JpegError('AC index out of range')
              blk[ZIGZAG[k]] = _extend(bits.bits(s), s)
              k += 1
  info['mcux'], info['mcuy'], info['hmax'], info['vmax'] = mcux, mcuy, hmax, vmax
  return info, coefs


# ---- jidctint.c: jpeg_idct_islow ------------------------------------------------------------------
CONST_BITS, PASS1_BITS = 13, 2
FIX_0_298631336, FIX_0_390180644, FIX_0_541196100, FIX_0_765366865 = 2446, 3196, 4433, 6270
FIX_0_899976223, FIX_1_175875602, FIX_1_501321110, FIX_1_847759065 = 7373, 9633, 12299, 15137
FIX_1_961570560, FIX_2_053119869, FIX_2_562915447, FIX_3_072711026 = 16069, 16819, 20995, 25172


def _descale(x, n):
  return (x + (1 << (n - 1))) >> n


def _idct_1d(d, shift, pass1):
  """One pass of jpeg_idct_islow over axis -1 of int64 array d [..., 8]; the even part scales by
  2^CONST_BITS, outputs are descaled by `shift`."""
  z2, z3 = d[..., 2], d[..., 6]
  z1 = (z2 + z3) * FIX_0_541196100
  tmp2 = z1 + z3 * (-FIX_1_847759065)
  tmp3 = z1 + z2 * FIX_0_765366865
  z2, z3 = d[..., 0], d[..., 4]
  tmp0 = (z2 + z3) << CONST_BITS
  tmp1 = (z2 - z3) << CONST_BITS
  tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
  tmp0, tmp1, tmp2, tmp3 = d[..., 7], d[..., 5], d[..., 3], d[..., 1]
  z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
  z5 = (z3 + z4) * FIX_1_175875602
  tmp0 = tmp0 * FIX_0_298631336
  tmp1 = tmp1 * FIX_2_053119869
  tmp2 = tmp2 * FIX_3_072711026
  tmp3 = tmp3 * FIX_1_501321110
  z1 = z1 * (-FIX_0_899976223)
  z2 = z2 * (-FIX_2_562915447)
  z3 = z3 * (-FIX_1_961570560) + z5
  z4 = z4 * (-FIX_0_390180644) + z5
  tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
  out = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0,
                  tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3], -1)
  return _descale(out, shift)


def idct_islow(coef, qt):
  """coef int16 [..., 64] (natural order), qt int32 [64] -> uint8 samples [..., 8, 8]."""
  d = (coef.astype(np.int64) * qt.astype(np.int64)).reshape(coef.shape[:-1] + (8, 8))
  # pass 1: columns (axis -2) -> work array scaled by 2^PASS1_BITS
  ws = _idct_1d(np.swapaxes(d, -1, -2), CONST_BITS - PASS1_BITS, True)
  ws = np.swapaxes(ws, -1, -2)
  # pass 2: rows
  out = _idct_1d(ws, CONST_BITS + PASS1_BITS + 3, False)
  return np.clip(out + 128, 0, 255).astype(np.uint8)   # range_limit table: (x + CENTERJSAMPLE) clamped


def _planes(info, coefs):
  planes = []
  for c, cf in zip(info['comps'], coefs):
    blk = idct_islow(cf, info['qt'][c[3]])                         # [bh, bw, 8, 8]
    bh, bw = blk.shape[:2]
    planes.append(blk.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8))
  return planes


def _h2v1_fancy(p):
  """jdsample.c h2v1_fancy_upsample: out[2i] = (3*in[i] + in[i-1] + 1) >> 2, out[2i+1] = (3*in[i] + in[i+1] + 2) >> 2,
  first / last columns replicated."""
  x = p.astype(np.int32)
  left = np.concatenate([x[:, :1], x[:, :-1]], 1)
  right = np.concatenate([x[:, 1:], x[:, -1:]], 1)
  out = np.empty((x.shape[0], x.shape[1] * 2), np.int32)
  out[:, 0::2] = (3 * x + left + 1) >> 2
  out[:, 1::2] = (3 * x + right + 2) >> 2
  out[:, 0] = x[:, 0]
  out[:, -1] = x[:, -1]
  return out.astype(np.uint8)


def _h2v2_fancy(p, rows_valid):
  """jdsample.c h2v2_fancy_upsample: vertical 3:1 blend with the nearer neighbour row (the context rows at
  the image top / bottom replicate the edge row), then horizontal 3:1 with rounding 8 / 7 and >> 4."""
  x = p[:rows_valid].astype(np.int32)
  up = np.concatenate([x[:1], x[:-1]], 0)
  down = np.concatenate([x[1:], x[-1:]], 0)
  out = np.empty((x.shape[0] * 2, x.shape[1] * 2), np.int32)
  for v, other in ((0, up), (1, down)):
    col = 3 * x + other                                   # "thiscolsum"
    left = np.concatenate([col[:, :1], col[:, :-1]], 1)
    right = np.concatenate([col[:, 1:], col[:, -1:]], 1)
    even = (3 * col + left + 8) >> 4
    odd = (3 * col + right + 7) >> 4
    even[:, 0] = (4 * col[:, 0] + 8) >> 4
    odd[:, -1] = (4 * col[:, -1] + 7) >> 4
    out[v::2, 0::2] = even
    out[v::2, 1::2] = odd
  return out.astype(np.uint8)


def _ycc_to_rgb(y, cb, cr):
  """jdcolor.c build_ycc_rgb_table / ycc_rgb_convert (SCALEBITS 16)."""
  one_half = 1 << 15
  x = np.arange(256, dtype=np.int64) - 128
  cr_r = (91881 * x + one_half) >> 16            # FIX(1.40200)
  cb_b = (116130 * x + one_half) >> 16           # FIX(1.77200)
  cr_g = -46802 * x                               # -FIX(0.71414)
  cb_g = -22554 * x + one_half                    # -FIX(0.34414) + ONE_HALF
  yy = y.astype(np.int64)
  r = yy + cr_r[cr]
  g = yy + ((cb_g[cb] + cr_g[cr]) >> 16)
  b = yy + cb_b[cb]
  return np.clip(np.stack([r, g, b], -1), 0, 255).astype(np.uint8)


def decode(data, channels=3):
  """bytes -> uint8 [H, W, channels] exactly as libjpeg-turbo (ISLOW, fancy upsampling) produces it."""
  info, coefs = decode_coefficients(bytes(data))
  h, w = info['height'], info['width']
  planes = _planes(info, coefs)
  comps = info['comps']
  if len(comps) == 1:
    grey = planes[0][:h, :w]
    return grey[..., None] if channels == 1 else np.repeat(grey[..., None], 3, -1)
  if len(comps) != 3:
    raise JpegError('only greyscale and YCbCr JPEGs are supported')
  hmax, vmax = info['hmax'], info['vmax']
  full = []
  for c, pl in zip(comps, planes):
    hs, vs = hmax // c[1], vmax // c[2]
    cw = -(-w * c[1] // hmax)          # downsampled_width
    ch = -(-h * c[2] // vmax)
    pl = pl[:, :cw]
    if (hs, vs) == (1, 1):
      up = pl
    elif cw <= 2:
      # jdsample.c jinit_upsampler: the fancy (triangle) filters are only selected when downsampled_width > 2;
      # narrower components are upsampled by replication (h2v1_upsample / h2v2_upsample)
      up = np.repeat(np.repeat(pl[:ch] if vs == 2 else pl, hs, 1), vs, 0)
    elif (hs, vs) == (2, 1):
      up = _h2v1_fancy(pl)
    elif (hs, vs) == (2, 2):
      up = _h2v2_fancy(pl, ch)
    else:
      raise JpegError('unsupported sampling factors %dx%d' % (hs, vs))
    full.append(up[:h, :w])
  rgb = _ycc_to_rgb(full[0], full[1], full[2])
  if channels == 1:
    # tf.image.decode_image(channels=1) asks libjpeg for JCS_GRAYSCALE: the Y plane
    return full[0][..., None]
  return rgb
