// jpeg_host.cc — host half of the split JPEG decoder: header parsing and Huffman entropy decoding of
// baseline (SOF0 / SOF1, 8-bit, single interleaved scan, optional restart intervals) JPEG streams into
// quantised DCT coefficient blocks.  The device half (jpeg.cu) does dequantisation, the ISLOW inverse
// DCT, fancy chroma upsampling and YCbCr -> RGB, bit-exactly like libjpeg(-turbo) with its defaults,
// which is what tf.image.decode_image runs for utils/tfdata.py:426-484.
//
// Entropy decoding is inherently serial per image (ITU-T T.81 Annex F), so it stays on host threads
// (one image per task); everything that is data parallel moves to the GPU and only int16 coefficients
// (already ~the size of the decoded image, but written once into pinned memory) cross PCIe.
#include <stdint.h>
#include <unistd.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/t2r_b200.h"

namespace t2r {
void set_error(const char* fmt, ...);
}

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

constexpr int kFastAcBits = 10;

struct Huff {
  // 9-bit lookahead: entry = (length << 8) | symbol for codes of length <= 9, 0 otherwise
  uint16_t fast[512];
  int32_t maxcode[18];  // largest code of each length (-1 if none), maxcode[17] = sentinel
  int32_t valptr[17];
  int32_t mincode[17];
  uint8_t symbols[256];
  // AC tables only: for a kFastAcBits lookahead that holds a whole (code, magnitude bits) pair with a small
  // value, (value << 8) | (run << 4) | total_bits; 0 otherwise (stb_image's "fast AC" idea)
  int16_t fast_ac[1 << kFastAcBits];
  bool present = false;
};

struct Parsed {
  T2RJpegInfo info;
  Huff dc[4], ac[4];
  int td[3] = {-1, -1, -1}, ta[3] = {-1, -1, -1};
  uint64_t scan_offset = 0;
};

// Frames beyond this are refused before anything is allocated for them (a corrupted SOF can claim 65535 x 65535).
constexpr int kMaxSide = 16384;
constexpr int64_t kMaxPixels = int64_t(1) << 26;

bool build_huff(const uint8_t* counts, const uint8_t* symbols, int total, Huff* h) {
  memset(h->fast, 0, sizeof(h->fast));
  memcpy(h->symbols, symbols, size_t(total));
  int code = 0, k = 0;
  for (int len = 1; len <= 16; ++len) {
    h->valptr[len] = k;
    h->mincode[len] = code;
    for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
      if (code >= (1 << len)) return false;          // over-subscribed table: would index past the lookahead
      if (len <= 9) {
        const int shift = 9 - len;
        for (int f = 0; f < (1 << shift); ++f) h->fast[(code << shift) | f] = uint16_t((len << 8) | symbols[k]);
      }
    }
    h->maxcode[len] = counts[len - 1] ? code - 1 : -1;
    if (code > (1 << len)) return false;
    code <<= 1;
  }
  h->maxcode[17] = 0x7fffffff;
  for (int i = 0; i < (1 << kFastAcBits); ++i) {
    h->fast_ac[i] = 0;
    const uint16_t f = h->fast[i >> (kFastAcBits - 9)];
    if (!f) continue;
    const int len = f >> 8, rs = f & 0xFF, run = rs >> 4, mag = rs & 15;
    if (mag == 0 || len + mag > kFastAcBits) continue;
    int k = (i >> (kFastAcBits - len - mag)) & ((1 << mag) - 1);   // the magnitude bits that follow the code
    if (k < (1 << (mag - 1))) k += 1 - (1 << mag);                  // EXTEND
    if (k >= -128 && k <= 127) h->fast_ac[i] = int16_t(k * 256 + run * 16 + len + mag);
  }
  h->present = true;
  return true;
}

inline uint32_t be16(const uint8_t* p) { return (uint32_t(p[0]) << 8) | p[1]; }

int parse_headers(const uint8_t* b, uint64_t len, Parsed* out) {
  memset(&out->info, 0, sizeof(out->info));
  out->info.struct_size = sizeof(T2RJpegInfo);
  if (len < 4 || b[0] != 0xFF || b[1] != 0xD8) { t2r::set_error("jpeg: no SOI marker"); return T2R_ERR_PARSE; }
  uint64_t p = 2;
  bool have_sof = false;
  for (;;) {
    if (p + 4 > len) { t2r::set_error("jpeg: truncated before SOS"); return T2R_ERR_PARSE; }
    if (b[p] != 0xFF) { t2r::set_error("jpeg: marker expected at byte %llu", (unsigned long long)p); return T2R_ERR_PARSE; }
    while (p + 1 < len && b[p + 1] == 0xFF) ++p;
    if (p + 2 > len) { t2r::set_error("jpeg: truncated marker"); return T2R_ERR_PARSE; }   // a run of fill bytes up to the end
    const uint8_t m = b[p + 1];
    p += 2;
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
    if (p + 2 > len) { t2r::set_error("jpeg: truncated segment"); return T2R_ERR_PARSE; }
    const uint32_t n = be16(b + p);
    if (n < 2 || p + n > len) { t2r::set_error("jpeg: bad segment length"); return T2R_ERR_PARSE; }
    const uint8_t* seg = b + p + 2;
    const uint32_t sl = n - 2;
    if (m == 0xDB) {
      uint32_t q = 0;
      while (q < sl) {
        const int pq = seg[q] >> 4, tq = seg[q] & 15;
        ++q;
        if (tq > 3 || q + (pq ? 128u : 64u) > sl) { t2r::set_error("jpeg: bad DQT"); return T2R_ERR_PARSE; }
        for (int i = 0; i < 64; ++i) {
          out->info.qt[tq][kZigzag[i]] = pq ? uint16_t(be16(seg + q)) : seg[q];
          q += pq ? 2 : 1;
        }
      }
    } else if (m == 0xC0 || m == 0xC1) {
      if (sl < 6 || seg[0] != 8) { t2r::set_error("jpeg: only 8-bit precision is supported"); return T2R_ERR_PARSE; }
      out->info.height = int32_t(be16(seg + 1));
      out->info.width = int32_t(be16(seg + 3));
      out->info.ncomp = seg[5];
      if ((out->info.ncomp != 1 && out->info.ncomp != 3) || sl < 6u + 3u * out->info.ncomp) {
        t2r::set_error("jpeg: %d components unsupported", out->info.ncomp);
        return T2R_ERR_PARSE;
      }
      for (int i = 0; i < out->info.ncomp; ++i) {
        out->info.comp_id[i] = seg[6 + 3 * i];
        out->info.h[i] = seg[7 + 3 * i] >> 4;
        out->info.v[i] = seg[7 + 3 * i] & 15;
        out->info.tq[i] = seg[8 + 3 * i];
        if (out->info.tq[i] > 3) { t2r::set_error("jpeg: bad quantisation table id"); return T2R_ERR_PARSE; }
      }
      have_sof = true;
    } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      t2r::set_error("jpeg: unsupported process (SOF marker 0x%02X): only baseline sequential Huffman", m);
      return T2R_ERR_PARSE;
    } else if (m == 0xC4) {
      uint32_t q = 0;
      while (q + 17 <= sl) {
        const int tc = seg[q] >> 4, th = seg[q] & 15;
        int total = 0;
        for (int i = 0; i < 16; ++i) total += seg[q + 1 + i];
        if (th > 3 || total > 256 || q + 17 + total > sl) { t2r::set_error("jpeg: bad DHT"); return T2R_ERR_PARSE; }
        if (!build_huff(seg + q + 1, seg + q + 17, total, tc ? &out->ac[th] : &out->dc[th])) {
          t2r::set_error("jpeg: inconsistent Huffman table");
          return T2R_ERR_PARSE;
        }
        q += 17 + total;
      }
    } else if (m == 0xDD) {
      if (sl < 2) { t2r::set_error("jpeg: bad DRI"); return T2R_ERR_PARSE; }
      out->info.restart_interval = int32_t(be16(seg));
    } else if (m == 0xDA) {
      if (!have_sof || sl < 1 || seg[0] != out->info.ncomp || sl < 1u + 2u * seg[0]) {
        t2r::set_error("jpeg: only single-scan (interleaved) streams are supported");
        return T2R_ERR_PARSE;
      }
      for (int i = 0; i < out->info.ncomp; ++i) {
        const int cid = seg[1 + 2 * i], t = seg[2 + 2 * i];
        for (int c = 0; c < out->info.ncomp; ++c)
          if (out->info.comp_id[c] == cid) { out->td[c] = t >> 4; out->ta[c] = t & 15; }
      }
      for (int c = 0; c < out->info.ncomp; ++c)
        if (out->td[c] < 0 || out->td[c] > 3 || out->ta[c] < 0 || out->ta[c] > 3) {
          t2r::set_error("jpeg: scan header does not name a valid Huffman table for every frame component");
          return T2R_ERR_PARSE;
        }
      out->scan_offset = p + n;
      break;
    }
    p += n;
  }
  T2RJpegInfo& in = out->info;
  if (in.width <= 0 || in.height <= 0) { t2r::set_error("jpeg: empty image"); return T2R_ERR_PARSE; }
  if (in.width > kMaxSide || in.height > kMaxSide || int64_t(in.width) * in.height > kMaxPixels) {
    t2r::set_error("jpeg: %d x %d frame exceeds the supported size", in.width, in.height);
    return T2R_ERR_PARSE;
  }
  int hmax = 1, vmax = 1;
  for (int c = 0; c < in.ncomp; ++c) {
    if (in.h[c] < 1 || in.h[c] > 2 || in.v[c] < 1 || in.v[c] > 2) {
      t2r::set_error("jpeg: sampling factors %dx%d unsupported", in.h[c], in.v[c]);
      return T2R_ERR_PARSE;
    }
    hmax = in.h[c] > hmax ? in.h[c] : hmax;
    vmax = in.v[c] > vmax ? in.v[c] : vmax;
  }
  if (in.ncomp == 1) { in.h[0] = in.v[0] = 1; hmax = vmax = 1; }   // a single component is never interleaved
  if (in.ncomp == 3 && (in.h[0] != hmax || in.v[0] != vmax || in.h[1] != 1 || in.v[1] != 1 || in.h[2] != 1 ||
                        in.v[2] != 1 || (hmax == 1 && vmax == 2))) {
    t2r::set_error("jpeg: unsupported sampling layout (luma %dx%d chroma %dx%d)", in.h[0], in.v[0], in.h[1], in.v[1]);
    return T2R_ERR_PARSE;
  }
  in.hmax = hmax; in.vmax = vmax;
  in.mcux = (in.width + 8 * hmax - 1) / (8 * hmax);
  in.mcuy = (in.height + 8 * vmax - 1) / (8 * vmax);
  int64_t off = 0;
  for (int c = 0; c < in.ncomp; ++c) {
    in.coef_offset[c] = off;
    off += int64_t(in.mcux) * in.h[c] * in.mcuy * in.v[c] * 64;
  }
  in.coef_count = off;
  return T2R_OK;
}

struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t acc = 0;
  int n = 0;
  bool hit_marker = false;
  bool exhausted = false;     // the input ended inside the scan (no marker): TF / libjpeg's JERR_INPUT_EOF
  inline void fill() {
    // fast path: four stream bytes at once when none of them is 0xFF (no stuffing, no marker)
    if (n <= 32 && !hit_marker && end - p >= 4) {
      const uint32_t v = (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3];
      const uint32_t inv = ~v;
      if (!((inv - 0x01010101u) & ~inv & 0x80808080u)) {   // no byte of v equals 0xFF
        acc |= uint64_t(v) << (32 - n);
        n += 32;
        p += 4;
        return;
      }
    }
    while (n <= 56) {
      uint32_t byte = 0;
      if (!hit_marker && p >= end) exhausted = true;
      if (!hit_marker && p < end) {
        byte = *p++;
        if (byte == 0xFF) {
          if (p >= end) exhausted = true;
          const uint8_t nx = p < end ? *p : 0;
          if (nx == 0) ++p;
          else { --p; byte = 0; hit_marker = true; }   // marker inside the scan: feed zeros, do not consume
        }
      }
      acc |= uint64_t(byte) << (56 - n);
      n += 8;
    }
  }
  inline uint32_t peek(int k) { return uint32_t(acc >> (64 - k)); }
  inline void skip(int k) { acc <<= k; n -= k; }
  inline void clamp() {}
  void restart() {
    n = 0; acc = 0; hit_marker = false;
    while (p + 1 < end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) ++p;
    p += 2;
  }
};

// Reader over an UN-STUFFED copy of the entropy-coded segment (FF00 -> FF, cut at the first marker, zero padded): the
// refill is one unconditional 4-byte big-endian load.  Used for scans without restart markers (what PIL / libjpeg
// write by default and what the replay records hold); scans with a restart interval keep the byte-wise BitReader.
struct FastBits {
  const uint8_t* p;            // first byte that is not yet completely in `acc`
  const uint8_t* lim;          // end of the real (un-stuffed) bytes; kScanPad zero bytes follow
  uint64_t acc = 0;            // bit buffer, next bit at the top
  int n = 0;                   // valid bits in `acc`
  bool past_end = false;       // the decoder consumed bits beyond `lim`
  bool drained() const { return past_end || p >= lim; }   // every real byte has been loaded
  // Branch-free refill: OR the next eight stream bytes in below the valid bits and advance by the whole bytes that fit.
  // Afterwards 56 <= n <= 63: one fill covers a 16-bit code plus its magnitude bits several times over.
  inline void fill() {
    uint64_t next;
    memcpy(&next, p, 8);
    next = __builtin_bswap64(next);
    acc |= next >> n;
    p += (63 - n) >> 3;
    n |= 56;
  }
  // Called once per block: a block consumes at most 64 * 31 bits = 248 bytes, the padding is larger.
  inline void clamp() {
    if (p > lim && (p - lim) * 8 > n) { p = lim; past_end = true; acc = 0; n = 0; }   // bits beyond the data were consumed
  }
  inline uint32_t peek(int k) { return uint32_t(acc >> (64 - k)); }
  inline void skip(int k) { acc <<= k; n -= k; }
  void restart() {}
};

constexpr size_t kScanPad = 512;

// Copies the scan that starts at data[0] into `out` with byte stuffing removed; stops at the first marker.  Returns
// true when a marker ended the segment, false when the input simply ran out.
bool unstuff_scan(const uint8_t* data, uint64_t len, std::vector<uint8_t>* out) {
  out->resize(size_t(len) + kScanPad);
  uint8_t* dst = out->data();
  const uint8_t* p = data;
  const uint8_t* end = data + len;
  bool marker = false;
  while (p < end) {
    const uint8_t* ff = static_cast<const uint8_t*>(memchr(p, 0xFF, size_t(end - p)));
    if (!ff) {
      memcpy(dst, p, size_t(end - p));
      dst += end - p;
      break;
    }
    memcpy(dst, p, size_t(ff - p));
    dst += ff - p;
    if (ff + 1 < end && ff[1] != 0) { marker = true; break; }     // FF xx: a marker, not data
    *dst++ = 0xFF;                                                  // FF 00 (or a lone trailing FF): the data byte FF
    p = ff + 2;
  }
  const size_t n = size_t(dst - out->data());
  memset(dst, 0, kScanPad);
  out->resize(n + kScanPad);
  return marker;
}

// The bits after a fill(): Huffman code (<= 16 bits) then `s` magnitude bits, no further refill needed.
template <class BR>
inline int decode_symbol(BR& br, const Huff& h) {
  const uint16_t f = h.fast[br.peek(9)];
  if (f) { br.skip(f >> 8); return f & 0xFF; }
  int32_t code = int32_t(br.peek(9));
  int len = 9;
  br.skip(9);
  while (len < 17 && (h.maxcode[len] < 0 || code > h.maxcode[len])) {
    code = (code << 1) | int32_t(br.peek(1));
    br.skip(1);
    ++len;
  }
  if (len > 16) return -1;
  return h.symbols[h.valptr[len] + code - h.mincode[len]];
}

template <class BR>
inline int32_t receive_extend(BR& br, int s) {
  if (!s) return 0;
  const int32_t v = int32_t(br.peek(s));
  br.skip(s);
  return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

template <class BR>
int entropy_decode_with(BR& br, const Parsed& ps, int16_t* coef) {
  const T2RJpegInfo& in = ps.info;
  int pred[3] = {0, 0, 0};
  int64_t count = 0;
  for (int my = 0; my < in.mcuy; ++my)
    for (int mx = 0; mx < in.mcux; ++mx) {
      if (in.restart_interval && count && count % in.restart_interval == 0) {
        br.restart();
        pred[0] = pred[1] = pred[2] = 0;
      }
      ++count;
      for (int c = 0; c < in.ncomp; ++c) {
        const int bw = in.mcux * in.h[c];
        const Huff& hdc = ps.dc[ps.td[c]];
        const Huff& hac = ps.ac[ps.ta[c]];
        for (int by = 0; by < in.v[c]; ++by)
          for (int bx = 0; bx < in.h[c]; ++bx) {
            int16_t* blk = coef + in.coef_offset[c] + (int64_t(my * in.v[c] + by) * bw + (mx * in.h[c] + bx)) * 64;
            br.clamp();
            br.fill();
            const int t = decode_symbol(br, hdc);
            if (t < 0 || t > 11) { t2r::set_error("jpeg: bad DC Huffman code"); return T2R_ERR_PARSE; }
            pred[c] += receive_extend(br, t);
            blk[0] = int16_t(pred[c]);
            for (int k = 1; k < 64;) {
              br.fill();
              const int fa = hac.fast_ac[br.peek(kFastAcBits)];
              if (fa) {                                   // code + magnitude bits in one lookup
                k += (fa >> 4) & 15;
                if (k > 63) { t2r::set_error("jpeg: AC coefficient index out of range"); return T2R_ERR_PARSE; }
                br.skip(fa & 15);
                blk[kZigzag[k++]] = int16_t(fa >> 8);
                continue;
              }
              const int rs = decode_symbol(br, hac);
              if (rs < 0) { t2r::set_error("jpeg: bad AC Huffman code"); return T2R_ERR_PARSE; }
              const int r = rs >> 4, s = rs & 15;
              if (s == 0) {
                if (r != 15) break;
                k += 16;
                continue;
              }
              k += r;
              if (k > 63) { t2r::set_error("jpeg: AC coefficient index out of range"); return T2R_ERR_PARSE; }
              blk[kZigzag[k]] = int16_t(receive_extend(br, s));
              ++k;
            }
          }
      }
    }
  return T2R_OK;
}

int entropy_decode(const uint8_t* data, uint64_t len, const Parsed& ps, int16_t* coef) {
  const T2RJpegInfo& in = ps.info;
  memset(coef, 0, size_t(in.coef_count) * sizeof(int16_t));
  for (int c = 0; c < in.ncomp; ++c)
    if (!ps.dc[ps.td[c]].present || !ps.ac[ps.ta[c]].present) {
      t2r::set_error("jpeg: scan refers to a Huffman table that was never defined");
      return T2R_ERR_PARSE;
    }
  bool exhausted;
  if (in.restart_interval == 0) {
    static thread_local std::vector<uint8_t> scan;
    const bool marker = unstuff_scan(data + ps.scan_offset, len - ps.scan_offset, &scan);
    FastBits br{scan.data(), scan.data() + scan.size() - kScanPad};
    const int rc = entropy_decode_with(br, ps, coef);
    if (rc != T2R_OK) return rc;
    br.clamp();
    exhausted = !marker && br.drained();     // no marker ends the segment and the reader ran into its end
  } else {
    BitReader br{data + ps.scan_offset, data + len};
    const int rc = entropy_decode_with(br, ps, coef);
    if (rc != T2R_OK) return rc;
    exhausted = br.exhausted;
  }
  if (exhausted) {      // what tf.image.decode_image reports for a truncated file (try_recover_truncated = False)
    t2r::set_error("jpeg: premature end of data inside the entropy-coded segment");
    return T2R_ERR_PARSE;
  }
  return T2R_OK;
}


// ---------------------------------------------------------------------------------------------
// Complete host decode (the 'host' image decoder of the record parser): the same integer arithmetic as the device half
// in jpeg.cu - jidctint.c ISLOW, jdsample.c fancy upsampling, jdcolor.c tables - so host, device and libjpeg-turbo
// outputs are bit-identical.  PIL's JPEG plugin keeps the GIL while decoding, so its thread pool does not scale;
// these run on plain C++ threads.
// ---------------------------------------------------------------------------------------------
constexpr int CONST_BITS = 13, PASS1_BITS = 2;
constexpr int F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270;
constexpr int F_0_899976223 = 7373, F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137;
constexpr int F_1_961570560 = 16069, F_2_053119869 = 16819, F_2_562915447 = 20995, F_3_072711026 = 25172;

inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// One 1-D pass of jpeg_idct_islow.  Products are formed in 32-bit like libjpeg's INT32 (values stay below 2^31 for
// 8-bit data); intermediate wrap-around cannot occur for coefficients a conforming stream can hold.
inline void idct8(const int* d, int stride, int* o, int ostride, int shift) {
  int z2 = d[2 * stride], z3 = d[6 * stride];
  int z1 = (z2 + z3) * F_0_541196100;
  const int t2e = z1 + z3 * (-F_1_847759065);
  const int t3e = z1 + z2 * F_0_765366865;
  const int t0e = int(uint32_t(d[0] + d[4 * stride]) << CONST_BITS);
  const int t1e = int(uint32_t(d[0] - d[4 * stride]) << CONST_BITS);
  const int tmp10 = t0e + t3e, tmp13 = t0e - t3e, tmp11 = t1e + t2e, tmp12 = t1e - t2e;
  int tmp0 = d[7 * stride], tmp1 = d[5 * stride], tmp2 = d[3 * stride], tmp3 = d[1 * stride];
  z1 = tmp0 + tmp3;
  z2 = tmp1 + tmp2;
  z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * F_1_175875602;
  tmp0 *= F_0_298631336;
  tmp1 *= F_2_053119869;
  tmp2 *= F_3_072711026;
  tmp3 *= F_1_501321110;
  z1 *= -F_0_899976223;
  z2 *= -F_2_562915447;
  z3 = z3 * (-F_1_961570560) + z5;
  z4 = z4 * (-F_0_390180644) + z5;
  tmp0 += z1 + z3;
  tmp1 += z2 + z4;
  tmp2 += z2 + z3;
  tmp3 += z1 + z4;
  o[0 * ostride] = descale(tmp10 + tmp3, shift); o[7 * ostride] = descale(tmp10 - tmp3, shift);
  o[1 * ostride] = descale(tmp11 + tmp2, shift); o[6 * ostride] = descale(tmp11 - tmp2, shift);
  o[2 * ostride] = descale(tmp12 + tmp1, shift); o[5 * ostride] = descale(tmp12 - tmp1, shift);
  o[3 * ostride] = descale(tmp13 + tmp0, shift); o[4 * ostride] = descale(tmp13 - tmp0, shift);
}

inline uint8_t clamp255(int v) { return uint8_t(v < 0 ? 0 : (v > 255 ? 255 : v)); }

void idct_block(const int16_t* c, const uint16_t* q, uint8_t* dst, int pw) {
  int in[64], ws[64], o[8];
  for (int i = 0; i < 64; ++i) in[i] = int(c[i]) * int(q[i]);
  for (int col = 0; col < 8; ++col) {                                                              // columns
    const int* d = in + col;
    if ((d[8] | d[16] | d[24] | d[32] | d[40] | d[48] | d[56]) == 0) {
      // jidctint.c's shortcut for a column without AC terms; the full butterfly gives exactly d[0] << PASS1_BITS too
      const int dc = int(uint32_t(d[0]) << PASS1_BITS);
      for (int r = 0; r < 8; ++r) ws[r * 8 + col] = dc;
      continue;
    }
    idct8(d, 8, ws + col, 8, CONST_BITS - PASS1_BITS);
  }
  for (int r = 0; r < 8; ++r) {                                                                    // rows
    const int* w = ws + r * 8;
    uint8_t* out = dst + size_t(r) * pw;
    if ((w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7]) == 0) {
      const uint8_t v = clamp255(descale(w[0], PASS1_BITS + 3) + 128);     // identical to the butterfly on zeros
      for (int k = 0; k < 8; ++k) out[k] = v;
      continue;
    }
    idct8(w, 1, o, 1, CONST_BITS + PASS1_BITS + 3);
    for (int k = 0; k < 8; ++k) out[k] = clamp255(o[k] + 128);
  }
}

// jdcolor.c build_ycc_rgb_table (SCALEBITS = 16): per-sample terms of the YCbCr -> RGB conversion.
struct YccTables {
  int cr_r[256], cb_b[256], cr_g[256], cb_g[256];
  YccTables() {
    for (int i = 0; i < 256; ++i) {
      const int x = i - 128;
      cr_r[i] = (91881 * x + 32768) >> 16;
      cb_b[i] = (116130 * x + 32768) >> 16;
      cr_g[i] = -46802 * x;
      cb_g[i] = -22554 * x + 32768;
    }
  }
};
const YccTables kYcc;

// One full-resolution row of a chroma plane (jdsample.c: h2v2 / h2v1 fancy upsampling, or a copy for 4:4:4).
void upsample_row(const uint8_t* pl, int pw, int cw, int ch, int hs, int vs, int y, int W, int* col, uint8_t* row) {
  if (hs == 1) {
    memcpy(row, pl + size_t(y) * pw, size_t(W));
    return;
  }
  if (cw <= 2) {     // jinit_upsampler picks the fancy filters only for downsampled_width > 2: replication otherwise
    const uint8_t* r0 = pl + size_t(vs == 2 ? y >> 1 : y) * pw;
    for (int x = 0; x < W; ++x) row[x] = r0[x >> 1];
    return;
  }
  if (vs == 1) {                                            // h2v1: 3/4 nearer + 1/4 further sample
    const uint8_t* r0 = pl + size_t(y) * pw;
    for (int cx = 0; cx < cw; ++cx) col[cx] = r0[cx];
    const int last = (W - 1) >> 1;                          // chroma column of the last pixel
    for (int cx = 1; cx < last; ++cx) {                     // interior: no edge cases, vectorisable
      row[2 * cx] = uint8_t((3 * col[cx] + col[cx - 1] + 1) >> 2);
      row[2 * cx + 1] = uint8_t((3 * col[cx] + col[cx + 1] + 2) >> 2);
    }
    for (int x = 0; x < W; ++x) {                           // the first and last chroma columns
      const int cx = x >> 1;
      if (cx > 0 && cx < last) { x = 2 * last - 1; continue; }
      const int v0 = col[cx];
      if (x & 1) row[x] = uint8_t(cx == cw - 1 ? v0 : (3 * v0 + col[cx + 1] + 2) >> 2);
      else row[x] = uint8_t(cx == 0 ? v0 : (3 * v0 + col[cx - 1] + 1) >> 2);
    }
    return;
  }
  const int cy = y >> 1;                                    // h2v2: vertical 3:1 first, then horizontal 3:1
  const int oy = (y & 1) ? (cy + 1 < ch ? cy + 1 : ch - 1) : (cy > 0 ? cy - 1 : 0);
  const uint8_t* r0 = pl + size_t(cy) * pw;
  const uint8_t* r1 = pl + size_t(oy) * pw;
  for (int cx = 0; cx < cw; ++cx) col[cx] = 3 * r0[cx] + r1[cx];
  const int last = (W - 1) >> 1;
  for (int cx = 1; cx < last; ++cx) {                       // interior: no edge cases, vectorisable
    row[2 * cx] = uint8_t((3 * col[cx] + col[cx - 1] + 8) >> 4);
    row[2 * cx + 1] = uint8_t((3 * col[cx] + col[cx + 1] + 7) >> 4);
  }
  for (int x = 0; x < W; ++x) {                             // the first and last chroma columns
    const int cx = x >> 1;
    if (cx > 0 && cx < last) { x = 2 * last - 1; continue; }
    if (x & 1) row[x] = uint8_t(cx == cw - 1 ? (4 * col[cx] + 7) >> 4 : (3 * col[cx] + col[cx + 1] + 7) >> 4);
    else row[x] = uint8_t(cx == 0 ? (4 * col[cx] + 8) >> 4 : (3 * col[cx] + col[cx - 1] + 8) >> 4);
  }
}

inline int chroma_at(const uint8_t* pl, int pw, int cw, int ch, int hs, int vs, int y, int x) {
  if (hs == 1) return pl[size_t(y) * pw + x];
  const int cx = x >> 1;
  if (vs == 1) {                                            // h2v1_fancy_upsample
    const int v0 = pl[size_t(y) * pw + cx];
    if (x & 1) return cx == cw - 1 ? v0 : (3 * v0 + pl[size_t(y) * pw + cx + 1] + 2) >> 2;
    return cx == 0 ? v0 : (3 * v0 + pl[size_t(y) * pw + cx - 1] + 1) >> 2;
  }
  const int cy = y >> 1;                                    // h2v2_fancy_upsample
  const int oy = (y & 1) ? (cy + 1 < ch ? cy + 1 : ch - 1) : (cy > 0 ? cy - 1 : 0);
  const uint8_t* r0 = pl + size_t(cy) * pw;
  const uint8_t* r1 = pl + size_t(oy) * pw;
  const int col = 3 * r0[cx] + r1[cx];
  if (x & 1) return cx == cw - 1 ? (4 * col + 7) >> 4 : (3 * col + 3 * r0[cx + 1] + r1[cx + 1] + 7) >> 4;
  return cx == 0 ? (4 * col + 8) >> 4 : (3 * col + 3 * r0[cx - 1] + r1[cx - 1] + 8) >> 4;
}

// A persistent worker pool: threads are created once (per process) and reused by every batch call, so a call does not
// pay thread creation and its workers are already spread over the cores.  run(n, fn) executes fn(0..n-1), the caller
// participating; nested or concurrent calls serialise on the pool mutex.  After a fork() the child rebuilds the pool.
class WorkerPool {
 public:
  static WorkerPool& instance() {
    static WorkerPool* pool = new WorkerPool();     // intentionally leaked: workers may outlive static destructors
    return *pool;
  }

  void run(int n, const std::function<void(int)>& fn) {
    if (n <= 1) {
      for (int i = 0; i < n; ++i) fn(i);
      return;
    }
    std::lock_guard<std::mutex> call_lock(call_mutex_);
    ensure_threads(n - 1);
    {
      std::lock_guard<std::mutex> lock(mutex_);
      fn_ = &fn;
      total_ = n;
      next_ = 0;
      pending_ = n;
      ++generation_;
    }
    wake_.notify_all();
    work_until_empty();
    std::unique_lock<std::mutex> lock(mutex_);
    done_.wait(lock, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void ensure_threads(int wanted) {
    const pid_t pid = getpid();
    if (pid != owner_) {              // first use, or a forked child: the parent's threads do not exist here
      threads_ = 0;
      owner_ = pid;
    }
    unsigned hw = std::thread::hardware_concurrency();
    const int cap = int(hw ? (hw > 16 ? 16 : hw) : 1) - 1;
    wanted = wanted < cap ? wanted : cap;
    while (threads_ < wanted) {
      std::thread(&WorkerPool::worker_loop, this).detach();
      ++threads_;
    }
  }

  void work_until_empty() {
    for (;;) {
      int i;
      const std::function<void(int)>* fn;
      {
        std::lock_guard<std::mutex> lock(mutex_);
        if (fn_ == nullptr || next_ >= total_) return;
        i = next_++;
        fn = fn_;
      }
      (*fn)(i);
      std::lock_guard<std::mutex> lock(mutex_);
      if (--pending_ == 0) done_.notify_all();
    }
  }

  void worker_loop() {
    unsigned long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lock(mutex_);
        wake_.wait(lock, [&] { return generation_ != seen; });
        seen = generation_;
      }
      work_until_empty();
    }
  }

  std::mutex call_mutex_, mutex_;
  std::condition_variable wake_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int total_ = 0, next_ = 0, pending_ = 0, threads_ = 0;
  unsigned long generation_ = 0;
  pid_t owner_ = 0;
};

struct Scratch {
  std::vector<int16_t> coef;
  std::vector<uint8_t> planes;
};
// The free list (like the pool) lives on the heap for the life of the process: reused buffers keep their pages.
std::mutex g_scratch_mutex;
std::vector<Scratch*>* g_scratch_free = new std::vector<Scratch*>();

Scratch* acquire_scratch() {
  std::lock_guard<std::mutex> lock(g_scratch_mutex);
  if (g_scratch_free->empty()) return new Scratch();
  Scratch* s = g_scratch_free->back();
  g_scratch_free->pop_back();
  return s;
}

void release_scratch(Scratch* s) {
  std::lock_guard<std::mutex> lock(g_scratch_mutex);
  if (g_scratch_free->size() < 64) g_scratch_free->push_back(s);
  else delete s;
}

int decode_one_host(const uint8_t* data, uint64_t len, int H, int W, int channels, uint8_t* out, std::vector<int16_t>* coef,
                    std::vector<uint8_t>* planes) {
  Parsed ps;
  int rc = parse_headers(data, len, &ps);
  if (rc != T2R_OK) return rc;
  const T2RJpegInfo& in = ps.info;
  if (in.width != W || in.height != H) {
    t2r::set_error("jpeg: stream is %d x %d, the caller expects %d x %d", in.width, in.height, W, H);
    return T2R_ERR_INVALID_ARG;
  }
  coef->resize(size_t(in.coef_count));
  planes->resize(size_t(in.coef_count));
  rc = entropy_decode(data, len, ps, coef->data());
  if (rc != T2R_OK) return rc;
  const int ncomp_needed = channels == 1 ? 1 : in.ncomp;
  int bw[3] = {0, 0, 0};
  for (int c = 0; c < in.ncomp; ++c) bw[c] = in.mcux * in.h[c];
  for (int c = 0; c < ncomp_needed; ++c) {
    const int bh = in.mcuy * in.v[c], pw = bw[c] * 8;
    const int16_t* cc = coef->data() + in.coef_offset[c];
    uint8_t* pl = planes->data() + in.coef_offset[c];
    for (int by = 0; by < bh; ++by)
      for (int bx = 0; bx < bw[c]; ++bx)
        idct_block(cc + (size_t(by) * bw[c] + bx) * 64, in.qt[in.tq[c]], pl + size_t(by) * 8 * pw + bx * 8, pw);
  }
  const uint8_t* yp = planes->data() + in.coef_offset[0];
  const int ypw = bw[0] * 8;
  if (in.ncomp == 1 || channels == 1) {
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const uint8_t v = yp[size_t(y) * ypw + x];
        uint8_t* o = out + (size_t(y) * W + x) * channels;
        for (int k = 0; k < channels; ++k) o[k] = v;
      }
    return T2R_OK;
  }
  const int hs = in.hmax / in.h[1], vs = in.vmax / in.v[1];
  const int cw = (W * in.h[1] + in.hmax - 1) / in.hmax, ch = (H * in.v[1] + in.vmax - 1) / in.vmax;
  const uint8_t* cbp = planes->data() + in.coef_offset[1];
  const uint8_t* crp = planes->data() + in.coef_offset[2];
  std::vector<int> col(size_t(cw) + 1);
  std::vector<uint8_t> cbrow(static_cast<size_t>(W)), crrow(static_cast<size_t>(W));
  for (int y = 0; y < H; ++y) {
    upsample_row(cbp, bw[1] * 8, cw, ch, hs, vs, y, W, col.data(), cbrow.data());
    upsample_row(crp, bw[2] * 8, cw, ch, hs, vs, y, W, col.data(), crrow.data());
    const uint8_t* yrow = yp + size_t(y) * ypw;
    uint8_t* o = out + size_t(y) * W * 3;
    for (int x = 0; x < W; ++x, o += 3) {
      const int yy = yrow[x], cb = cbrow[x], cr = crrow[x];
      o[0] = clamp255(yy + kYcc.cr_r[cr]);
      o[1] = clamp255(yy + ((kYcc.cb_g[cb] + kYcc.cr_g[cr]) >> 16));
      o[2] = clamp255(yy + kYcc.cb_b[cb]);
    }
  }
  return T2R_OK;
}

}  // namespace

extern "C" int32_t t2r_jpeg_parse(const uint8_t* data, uint64_t len, T2RJpegInfo* info) {
  if (!data || !info) { t2r::set_error("jpeg_parse: null pointer"); return T2R_ERR_INVALID_ARG; }
  Parsed ps;
  const int rc = parse_headers(data, len, &ps);
  if (rc == T2R_OK) *info = ps.info;
  return rc;
}

extern "C" int32_t t2r_jpeg_entropy_decode_batch(const uint8_t* const* data, const uint64_t* lens, int32_t B,
                                                 T2RJpegInfo* infos, int16_t* coef, int64_t coef_stride) {
  if (!data || !lens || !infos || !coef || B <= 0) { t2r::set_error("jpeg_entropy_decode_batch: bad args"); return T2R_ERR_INVALID_ARG; }
  std::vector<int> rcs(size_t(B), T2R_OK);
  std::vector<std::string> msgs(static_cast<size_t>(B));
  WorkerPool::instance().run(B, [&](int b) {
    Parsed ps;
    int rc = parse_headers(data[b], lens[b], &ps);
    if (rc == T2R_OK && ps.info.coef_count > coef_stride) {
      t2r::set_error("jpeg: image %d needs %lld coefficients, stride is %lld", b, (long long)ps.info.coef_count,
                     (long long)coef_stride);
      rc = T2R_ERR_INVALID_ARG;
    }
    if (rc == T2R_OK) rc = entropy_decode(data[b], lens[b], ps, coef + int64_t(b) * coef_stride);
    if (rc != T2R_OK) msgs[size_t(b)] = t2r_last_error();
    infos[b] = ps.info;
    rcs[size_t(b)] = rc;
  });
  for (int b = 0; b < B; ++b)
    if (rcs[size_t(b)] != T2R_OK) {
      t2r::set_error("%s", msgs[size_t(b)].c_str());
      return rcs[size_t(b)];
    }
  return T2R_OK;
}

extern "C" int32_t t2r_jpeg_decode_host_batch(const uint8_t* const* data, const uint64_t* lens, int32_t B, int32_t H, int32_t W,
                                              int32_t channels, uint8_t* out) {
  if (!data || !lens || !out || B <= 0 || H <= 0 || W <= 0 || (channels != 1 && channels != 3)) {
    t2r::set_error("jpeg_decode_host_batch: bad args");
    return T2R_ERR_INVALID_ARG;
  }
  std::vector<int> rcs(size_t(B), T2R_OK);
  std::vector<std::string> msgs(static_cast<size_t>(B));
  const size_t frame = size_t(H) * W * channels;
  WorkerPool::instance().run(B, [&](int b) {
    Scratch* scratch = acquire_scratch();      // coefficient / plane buffers persist across calls (no page faults)
    const int rc = decode_one_host(data[b], lens[b], H, W, channels, out + size_t(b) * frame, &scratch->coef,
                                   &scratch->planes);
    if (rc != T2R_OK) msgs[size_t(b)] = t2r_last_error();
    rcs[size_t(b)] = rc;
    release_scratch(scratch);
  });
  for (int b = 0; b < B; ++b)
    if (rcs[size_t(b)] != T2R_OK) {
      t2r::set_error("%s", msgs[size_t(b)].c_str());
      return rcs[size_t(b)];
    }
  return T2R_OK;
}
