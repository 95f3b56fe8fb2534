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
#include <string.h>

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

struct Huff {
  // 9-bit lookahead: entry = (length << 8) | symbol for codes of length <= 9, 0 otherwise
  uint16_t fast[512];
  int32_t maxcode[18];  // largest code of each length (-1 if none), maxcode[17] = sentinel
  int32_t valptr[17];
  int32_t mincode[17];
  uint8_t symbols[256];
  // AC tables only: for a 9-bit lookahead that holds a whole (code, magnitude bits) pair with a small
  // value, (value << 8) | (run << 4) | total_bits; 0 otherwise (stb_image's "fast AC" idea)
  int16_t fast_ac[512];
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
  for (int i = 0; i < 512; ++i) {
    h->fast_ac[i] = 0;
    const uint16_t f = h->fast[i];
    if (!f) continue;
    const int len = f >> 8, rs = f & 0xFF, run = rs >> 4, mag = rs & 15;
    if (mag == 0 || len + mag > 9) continue;
    int k = ((i << len) & 511) >> (9 - mag);          // the magnitude bits that follow the code
    if (k < (1 << (mag - 1))) k += 1 - (1 << mag);    // EXTEND
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
      if (!hit_marker && p < end) {
        byte = *p++;
        if (byte == 0xFF) {
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
  inline int32_t receive_extend(int s) {
    if (!s) return 0;
    fill();
    const int32_t v = int32_t(peek(s));
    skip(s);
    return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
  }
  inline int decode(const Huff& h) {
    fill();
    const uint16_t f = h.fast[peek(9)];
    if (f) { skip(f >> 8); return f & 0xFF; }
    int32_t code = int32_t(peek(9));
    int len = 9;
    skip(9);
    while (len < 17 && (h.maxcode[len] < 0 || code > h.maxcode[len])) {
      code = (code << 1) | int32_t(peek(1));
      skip(1);
      ++len;
    }
    if (len > 16) return -1;
    return h.symbols[h.valptr[len] + code - h.mincode[len]];
  }
  void restart() {
    n = 0; acc = 0; hit_marker = false;
    while (p + 1 < end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) ++p;
    p += 2;
  }
};

int entropy_decode(const uint8_t* data, uint64_t len, const Parsed& ps, int16_t* coef) {
  const T2RJpegInfo& in = ps.info;
  memset(coef, 0, size_t(in.coef_count) * sizeof(int16_t));
  for (int c = 0; c < in.ncomp; ++c)
    if (!ps.dc[ps.td[c]].present || !ps.ac[ps.ta[c]].present) {
      t2r::set_error("jpeg: scan refers to a Huffman table that was never defined");
      return T2R_ERR_PARSE;
    }
  BitReader br{data + ps.scan_offset, data + len};
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
        for (int by = 0; by < in.v[c]; ++by)
          for (int bx = 0; bx < in.h[c]; ++bx) {
            int16_t* blk = coef + in.coef_offset[c] + (int64_t(my * in.v[c] + by) * bw + (mx * in.h[c] + bx)) * 64;
            const int t = br.decode(ps.dc[ps.td[c]]);
            if (t < 0 || t > 11) { t2r::set_error("jpeg: bad DC Huffman code"); return T2R_ERR_PARSE; }
            pred[c] += br.receive_extend(t);
            blk[0] = int16_t(pred[c]);
            const Huff& hac = ps.ac[ps.ta[c]];
            for (int k = 1; k < 64;) {
              br.fill();
              const int fa = hac.fast_ac[br.peek(9)];
              if (fa) {                                   // code + magnitude bits in one lookup
                k += (fa >> 4) & 15;
                if (k > 63) { t2r::set_error("jpeg: AC coefficient index out of range"); return T2R_ERR_PARSE; }
                br.skip(fa & 15);
                blk[kZigzag[k++]] = int16_t(fa >> 8);
                continue;
              }
              const int rs = br.decode(hac);
              if (rs < 0) { t2r::set_error("jpeg: bad AC Huffman code"); return T2R_ERR_PARSE; }
              const int r = rs >> 4, s = rs & 15;
              if (s == 0) {
                if (r != 15) break;
                k += 16;
                continue;
              }
              k += r;
              if (k > 63) { t2r::set_error("jpeg: AC coefficient index out of range"); return T2R_ERR_PARSE; }
              blk[kZigzag[k]] = int16_t(br.receive_extend(s));
              ++k;
            }
          }
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
  unsigned hw = std::thread::hardware_concurrency();
  const int nthreads = B >= 8 ? int(hw ? (hw > 16 ? 16 : hw) : 1) : 1;
  std::vector<int> rcs(size_t(B), T2R_OK);
  std::vector<std::string> msgs(static_cast<size_t>(nthreads));
  auto work = [&](int t) {
    for (int b = t; b < B; b += nthreads) {
      Parsed ps;
      int rc = parse_headers(data[b], lens[b], &ps);
      if (rc == T2R_OK && ps.info.coef_count > coef_stride) {
        t2r::set_error("jpeg: image %d needs %lld coefficients, stride is %lld", b, (long long)ps.info.coef_count,
                       (long long)coef_stride);
        rc = T2R_ERR_INVALID_ARG;
      }
      if (rc == T2R_OK) rc = entropy_decode(data[b], lens[b], ps, coef + int64_t(b) * coef_stride);
      if (rc != T2R_OK) { msgs[size_t(t)] = t2r_last_error(); }
      infos[b] = ps.info;
      rcs[size_t(b)] = rc;
    }
  };
  if (nthreads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  for (int b = 0; b < B; ++b)
    if (rcs[size_t(b)] != T2R_OK) {
      for (auto& m : msgs)
        if (!m.empty()) { t2r::set_error("%s", m.c_str()); break; }
      return rcs[size_t(b)];
    }
  return T2R_OK;
}
