// host_io.cc — host-side record path: CRC-32C, TFRecord framing, tf.Example wire-format parsing.
//
// Replaces the tf.data C++ runtime pieces the reference reaches through
// utils/tfdata.py:174-210 (TFRecordDataset), :385 (tf.parse_example) for the per-replay-batch
// input path.  No protobuf dependency: the wire format is walked directly
// and values are copied bit-exactly into caller-owned (pinned) buffers.
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../include/t2r_b200.h"

namespace t2r {
void set_error(const char* fmt, ...);
}

namespace {

uint32_t g_crc_table[8][256];
std::atomic<bool> g_crc_ready{false};

void crc_init() {
  if (g_crc_ready.load(std::memory_order_acquire)) return;
  const uint32_t poly = 0x82F63B78u;  // reflected Castagnoli
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ poly : (c >> 1);
    g_crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t)
      g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xFF];
  g_crc_ready.store(true, std::memory_order_release);
}

inline uint32_t crc32c_sw(uint32_t crc, const uint8_t* p, uint64_t n) {
  // slicing-by-8
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= crc;
    crc = g_crc_table[7][lo & 0xFF] ^ g_crc_table[6][(lo >> 8) & 0xFF] ^ g_crc_table[5][(lo >> 16) & 0xFF] ^
          g_crc_table[4][lo >> 24] ^ g_crc_table[3][hi & 0xFF] ^ g_crc_table[2][(hi >> 8) & 0xFF] ^
          g_crc_table[1][(hi >> 16) & 0xFF] ^ g_crc_table[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) crc = (crc >> 8) ^ g_crc_table[0][(crc ^ *p++) & 0xFF];
  return crc;
}

#if defined(__x86_64__)
__attribute__((target("sse4.2"))) uint32_t crc32c_hw(uint32_t crc, const uint8_t* p, uint64_t n) {
  uint64_t c = crc;
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    c = __builtin_ia32_crc32di(c, v);
    p += 8;
    n -= 8;
  }
  uint32_t c32 = uint32_t(c);
  while (n--) c32 = __builtin_ia32_crc32qi(c32, *p++);
  return c32;
}
bool have_sse42() { return __builtin_cpu_supports("sse4.2"); }
#else
uint32_t crc32c_hw(uint32_t crc, const uint8_t* p, uint64_t n) { return crc32c_sw(crc, p, n); }
bool have_sse42() { return false; }
#endif

inline uint32_t mask_crc(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; }

// ---- protobuf wire helpers ----
struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
};

inline bool read_varint(Cursor& c, uint64_t* out) {
  uint64_t v = 0;
  int shift = 0;
  while (c.p < c.end && shift < 64) {
    const uint8_t b = *c.p++;
    v |= uint64_t(b & 0x7F) << shift;
    if (!(b & 0x80)) {
      *out = v;
      return true;
    }
    shift += 7;
  }
  return false;
}

inline bool read_len(Cursor& c, Cursor* sub) {
  uint64_t n;
  if (!read_varint(c, &n) || n > uint64_t(c.end - c.p)) return false;
  sub->p = c.p;
  sub->end = c.p + n;
  c.p += n;
  return true;
}

inline bool skip_field(Cursor& c, uint32_t wire) {
  uint64_t tmp;
  Cursor sub;
  switch (wire) {
    case 0: return read_varint(c, &tmp);
    case 1: if (c.end - c.p < 8) return false; c.p += 8; return true;
    case 2: return read_len(c, &sub);
    case 5: if (c.end - c.p < 4) return false; c.p += 4; return true;
    default: return false;
  }
}

struct ParseError {
  char msg[256];
  bool set = false;
  void fail(const char* fmt, ...) {
    if (set) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg, sizeof(msg), fmt, ap);
    va_end(ap);
    set = true;
  }
};

// Decode one Feature message into plan slot f for batch row b.  Returns element count or -1.
long decode_feature(Cursor feat, const T2RFeaturePlan& f, int b, ParseError& err) {
  const int cap = f.count >= 0 ? f.count : -f.count;
  long n = 0;
  while (feat.p < feat.end) {
    uint64_t tag;
    if (!read_varint(feat, &tag)) { err.fail("bad Feature tag in '%s'", f.key); return -1; }
    const uint32_t field = uint32_t(tag >> 3), wire = uint32_t(tag & 7);
    if (wire != 2 || field < 1 || field > 3) {
      if (!skip_field(feat, wire)) { err.fail("bad Feature field in '%s'", f.key); return -1; }
      continue;
    }
    Cursor list;
    if (!read_len(feat, &list)) { err.fail("truncated list in '%s'", f.key); return -1; }
    const int kind = field == 1 ? T2R_DT_BYTES : (field == 2 ? T2R_DT_FLOAT : T2R_DT_INT64);
    if (kind != f.dtype) {
      if (list.p == list.end) continue;  // empty list of another kind: treat as missing values
      err.fail("feature '%s': stored kind %d, expected %d", f.key, kind, f.dtype);
      return -1;
    }
    while (list.p < list.end) {
      uint64_t t2;
      if (!read_varint(list, &t2)) { err.fail("bad list tag in '%s'", f.key); return -1; }
      const uint32_t lf = uint32_t(t2 >> 3), lw = uint32_t(t2 & 7);
      if (lf != 1) {
        if (!skip_field(list, lw)) { err.fail("bad list field in '%s'", f.key); return -1; }
        continue;
      }
      if (kind == T2R_DT_BYTES) {
        Cursor s;
        if (lw != 2 || !read_len(list, &s)) { err.fail("bad bytes in '%s'", f.key); return -1; }
        if (n < cap) {
          static_cast<const uint8_t**>(f.dst)[(long long)b * cap + n] = s.p;
          f.dst_len[(long long)b * cap + n] = uint64_t(s.end - s.p);
        }
        ++n;
      } else if (kind == T2R_DT_FLOAT) {
        float* dst = static_cast<float*>(f.dst) + (long long)b * cap;
        if (lw == 2) {  // packed
          Cursor s;
          if (!read_len(list, &s) || ((s.end - s.p) % 4) != 0) { err.fail("bad packed floats in '%s'", f.key); return -1; }
          const long m = long((s.end - s.p) / 4);
          for (long i = 0; i < m; ++i, ++n)
            if (n < cap) memcpy(dst + n, s.p + 4 * i, 4);
        } else if (lw == 5) {
          if (list.end - list.p < 4) { err.fail("truncated float in '%s'", f.key); return -1; }
          if (n < cap) memcpy(dst + n, list.p, 4);
          list.p += 4;
          ++n;
        } else { err.fail("bad float wire type in '%s'", f.key); return -1; }
      } else {
        int64_t* dst = static_cast<int64_t*>(f.dst) + (long long)b * cap;
        if (lw == 2) {
          Cursor s;
          if (!read_len(list, &s)) { err.fail("bad packed int64 in '%s'", f.key); return -1; }
          while (s.p < s.end) {
            uint64_t v;
            if (!read_varint(s, &v)) { err.fail("bad varint in '%s'", f.key); return -1; }
            if (n < cap) dst[n] = int64_t(v);
            ++n;
          }
        } else if (lw == 0) {
          uint64_t v;
          if (!read_varint(list, &v)) { err.fail("bad varint in '%s'", f.key); return -1; }
          if (n < cap) dst[n] = int64_t(v);
          ++n;
        } else { err.fail("bad int64 wire type in '%s'", f.key); return -1; }
      }
    }
  }
  return n;
}

void fill_default(const T2RFeaturePlan& f, int b, long from) {
  const int cap = f.count >= 0 ? f.count : -f.count;
  for (long i = from; i < cap; ++i) {
    const long long o = (long long)b * cap + i;
    if (f.dtype == T2R_DT_FLOAT) static_cast<float*>(f.dst)[o] = f.pad_float;
    else if (f.dtype == T2R_DT_INT64) static_cast<int64_t*>(f.dst)[o] = f.pad_int64;
    else {
      static_cast<const uint8_t**>(f.dst)[o] = nullptr;
      f.dst_len[o] = 0;
    }
  }
}

bool parse_one(const uint8_t* rec, uint64_t len, int b, T2RFeaturePlan* plan, int nf, ParseError& err) {
  std::vector<char> seen(nf, 0);
  Cursor ex{rec, rec + len};
  while (ex.p < ex.end) {
    uint64_t tag;
    if (!read_varint(ex, &tag)) { err.fail("record %d: bad Example tag", b); return false; }
    if ((tag >> 3) != 1 || (tag & 7) != 2) {
      if (!skip_field(ex, uint32_t(tag & 7))) { err.fail("record %d: bad Example field", b); return false; }
      continue;
    }
    Cursor feats;
    if (!read_len(ex, &feats)) { err.fail("record %d: truncated Features", b); return false; }
    while (feats.p < feats.end) {
      uint64_t t2;
      if (!read_varint(feats, &t2)) { err.fail("record %d: bad Features tag", b); return false; }
      if ((t2 >> 3) != 1 || (t2 & 7) != 2) {
        if (!skip_field(feats, uint32_t(t2 & 7))) { err.fail("record %d: bad Features field", b); return false; }
        continue;
      }
      Cursor entry;
      if (!read_len(feats, &entry)) { err.fail("record %d: truncated map entry", b); return false; }
      Cursor key{nullptr, nullptr}, val{nullptr, nullptr};
      while (entry.p < entry.end) {
        uint64_t t3;
        if (!read_varint(entry, &t3)) { err.fail("record %d: bad entry tag", b); return false; }
        Cursor sub;
        if ((t3 & 7) != 2) {
          if (!skip_field(entry, uint32_t(t3 & 7))) { err.fail("record %d: bad entry field", b); return false; }
          continue;
        }
        if (!read_len(entry, &sub)) { err.fail("record %d: truncated entry", b); return false; }
        if ((t3 >> 3) == 1) key = sub;
        else if ((t3 >> 3) == 2) val = sub;
      }
      if (!key.p) continue;
      const size_t klen = size_t(key.end - key.p);
      for (int i = 0; i < nf; ++i) {
        const T2RFeaturePlan& f = plan[i];
        if (strlen(f.key) != klen || memcmp(f.key, key.p, klen) != 0) continue;
        Cursor v = val.p ? val : Cursor{key.end, key.end};
        const long n = decode_feature(v, f, b, err);
        if (n < 0) return false;
        if (f.count >= 0) {
          if (n == 0 && !f.required) break;  // present but empty: treated as missing
          if (n != f.count) {
            err.fail("record %d: feature '%s' has %ld values, expected %d", b, f.key, n, f.count);
            return false;
          }
        } else {
          const long cap = -f.count;
          if (f.dst_len && f.dtype != T2R_DT_BYTES) f.dst_len[b] = uint64_t(n < cap ? n : cap);
          fill_default(f, b, n < cap ? n : cap);
        }
        seen[i] = 1;
        // the same key may feed several plan entries (same `name`, different spec paths)
      }
    }
  }
  for (int i = 0; i < nf; ++i) {
    if (seen[i]) continue;
    const T2RFeaturePlan& f = plan[i];
    if (f.required && f.count >= 0) {
      err.fail("record %d: required feature '%s' missing", b, f.key);
      return false;
    }
    if (f.count < 0 && f.dst_len && f.dtype != T2R_DT_BYTES) f.dst_len[b] = 0;
    fill_default(f, b, 0);
  }
  return true;
}

// tf.SequenceExample { Features context = 1; FeatureLists feature_lists = 2; }: walks feature_lists
// and, for every planned key, decodes step t's Feature into dst[(b * max_steps + t) * count ...].
bool parse_one_sequence(const uint8_t* rec, uint64_t len, int b, int B, const T2RFeaturePlan* plan, int nf,
                        int max_steps, int64_t* steps, ParseError& err) {
  for (int i = 0; i < nf; ++i) steps[(long long)i * B + b] = 0;
  Cursor ex{rec, rec + len};
  while (ex.p < ex.end) {
    uint64_t tag;
    if (!read_varint(ex, &tag)) { err.fail("record %d: bad SequenceExample tag", b); return false; }
    if ((tag >> 3) != 2 || (tag & 7) != 2) {
      if (!skip_field(ex, uint32_t(tag & 7))) { err.fail("record %d: bad SequenceExample field", b); return false; }
      continue;
    }
    Cursor lists;
    if (!read_len(ex, &lists)) { err.fail("record %d: truncated FeatureLists", b); return false; }
    while (lists.p < lists.end) {
      uint64_t t2;
      if (!read_varint(lists, &t2)) { err.fail("record %d: bad FeatureLists tag", b); return false; }
      if ((t2 >> 3) != 1 || (t2 & 7) != 2) {
        if (!skip_field(lists, uint32_t(t2 & 7))) { err.fail("record %d: bad FeatureLists field", b); return false; }
        continue;
      }
      Cursor entry;
      if (!read_len(lists, &entry)) { err.fail("record %d: truncated feature_list entry", b); return false; }
      Cursor key{nullptr, nullptr}, val{nullptr, nullptr};
      while (entry.p < entry.end) {
        uint64_t t3;
        if (!read_varint(entry, &t3)) { err.fail("record %d: bad entry tag", b); return false; }
        Cursor sub;
        if ((t3 & 7) != 2) {
          if (!skip_field(entry, uint32_t(t3 & 7))) { err.fail("record %d: bad entry field", b); return false; }
          continue;
        }
        if (!read_len(entry, &sub)) { err.fail("record %d: truncated entry", b); return false; }
        if ((t3 >> 3) == 1) key = sub;
        else if ((t3 >> 3) == 2) val = sub;
      }
      if (!key.p || !val.p) continue;
      const size_t klen = size_t(key.end - key.p);
      for (int i = 0; i < nf; ++i) {
        const T2RFeaturePlan& f = plan[i];
        if (strlen(f.key) != klen || memcmp(f.key, key.p, klen) != 0) continue;
        Cursor fl = val;   // FeatureList { repeated Feature feature = 1; }
        int64_t t = 0;
        while (fl.p < fl.end) {
          uint64_t t4;
          if (!read_varint(fl, &t4)) { err.fail("record %d: bad FeatureList tag in '%s'", b, f.key); return false; }
          if ((t4 >> 3) != 1 || (t4 & 7) != 2) {
            if (!skip_field(fl, uint32_t(t4 & 7))) { err.fail("record %d: bad FeatureList field", b); return false; }
            continue;
          }
          Cursor feat;
          if (!read_len(fl, &feat)) { err.fail("record %d: truncated step of '%s'", b, f.key); return false; }
          if (max_steps > 0) {
            if (t >= max_steps) { err.fail("record %d: '%s' has more than %d steps", b, f.key, max_steps); return false; }
            T2RFeaturePlan step = f;
            const long long off = ((long long)b * max_steps + t) * f.count;
            if (f.dtype == T2R_DT_FLOAT) step.dst = static_cast<float*>(f.dst) + off;
            else if (f.dtype == T2R_DT_INT64) step.dst = static_cast<int64_t*>(f.dst) + off;
            else { step.dst = static_cast<const uint8_t**>(f.dst) + off; step.dst_len = f.dst_len + off; }
            const long n = decode_feature(feat, step, 0, err);
            if (n < 0) return false;
            if (n != f.count) {
              err.fail("record %d: step %lld of '%s' has %ld values, expected %d", b, (long long)t, f.key, n, f.count);
              return false;
            }
          }
          ++t;
        }
        steps[(long long)i * B + b] = t;
      }
    }
  }
  return true;
}

}  // namespace

extern "C" int32_t t2r_sequence_example_parse_batch(const uint8_t* const* records, const uint64_t* lengths,
                                                    int32_t B, const T2RFeaturePlan* plan, int32_t n_features,
                                                    int32_t max_steps, int64_t* steps) {
  if (!records || !lengths || !plan || !steps || B <= 0 || n_features <= 0 || max_steps < 0) {
    t2r::set_error("sequence_example_parse_batch: bad args");
    return T2R_ERR_INVALID_ARG;
  }
  for (int i = 0; i < n_features; ++i)
    if (!plan[i].key || plan[i].count <= 0 || (max_steps > 0 && (!plan[i].dst || (plan[i].dtype == T2R_DT_BYTES && !plan[i].dst_len)))) {
      t2r::set_error("sequence_example_parse_batch: bad plan entry %d", i);
      return T2R_ERR_INVALID_ARG;
    }
  ParseError err;
  for (int b = 0; b < B; ++b)
    if (!parse_one_sequence(records[b], lengths[b], b, B, plan, n_features, max_steps, steps, err)) {
      t2r::set_error("%s", err.msg);
      return T2R_ERR_PARSE;
    }
  return T2R_OK;
}

extern "C" uint32_t t2r_crc32c(const uint8_t* data, uint64_t n) {
  static const bool hw = have_sse42();
  if (!hw) crc_init();
  const uint32_t c = hw ? crc32c_hw(0xFFFFFFFFu, data, n) : crc32c_sw(0xFFFFFFFFu, data, n);
  return c ^ 0xFFFFFFFFu;
}

extern "C" uint32_t t2r_masked_crc32c(const uint8_t* data, uint64_t n) { return mask_crc(t2r_crc32c(data, n)); }

extern "C" int64_t t2r_tfrecord_index(const uint8_t* file, uint64_t file_len, uint64_t* offsets,
                                      uint64_t* lengths, int64_t max_records, int32_t verify_crc) {
  if (!file && file_len) {
    t2r::set_error("tfrecord_index: null file");
    return T2R_ERR_INVALID_ARG;
  }
  uint64_t pos = 0;
  int64_t n = 0;
  while (pos < file_len) {
    if (file_len - pos < 12) {
      t2r::set_error("tfrecord_index: truncated header at byte %llu", (unsigned long long)pos);
      return T2R_ERR_PARSE;
    }
    uint64_t len;
    uint32_t len_crc;
    memcpy(&len, file + pos, 8);
    memcpy(&len_crc, file + pos + 8, 4);
    if (verify_crc && t2r_masked_crc32c(file + pos, 8) != len_crc) {
      t2r::set_error("tfrecord_index: length CRC mismatch at record %lld", (long long)n);
      return T2R_ERR_PARSE;
    }
    if (len > file_len - pos - 12 || file_len - pos - 12 - len < 4) {
      t2r::set_error("tfrecord_index: truncated record %lld", (long long)n);
      return T2R_ERR_PARSE;
    }
    const uint64_t data_off = pos + 12;
    if (verify_crc) {
      uint32_t data_crc;
      memcpy(&data_crc, file + data_off + len, 4);
      if (t2r_masked_crc32c(file + data_off, len) != data_crc) {
        t2r::set_error("tfrecord_index: data CRC mismatch at record %lld", (long long)n);
        return T2R_ERR_PARSE;
      }
    }
    if (n < max_records) {
      if (offsets) offsets[n] = data_off;
      if (lengths) lengths[n] = len;
    }
    ++n;
    pos = data_off + len + 4;
  }
  return n;
}

extern "C" int32_t t2r_example_parse_batch(const uint8_t* const* records, const uint64_t* lengths,
                                           int32_t B, T2RFeaturePlan* plan, int32_t n_features) {
  if (!records || !lengths || !plan || B <= 0 || n_features <= 0) {
    t2r::set_error("example_parse_batch: bad args");
    return T2R_ERR_INVALID_ARG;
  }
  for (int i = 0; i < n_features; ++i) {
    if (!plan[i].key || !plan[i].dst || plan[i].count == 0 ||
        (plan[i].dtype == T2R_DT_BYTES && !plan[i].dst_len)) {
      t2r::set_error("example_parse_batch: bad plan entry %d", i);
      return T2R_ERR_INVALID_ARG;
    }
  }
  unsigned hw = std::thread::hardware_concurrency();
  int nthreads = B >= 64 ? int(hw ? (hw > 8 ? 8 : hw) : 1) : 1;
  std::vector<ParseError> errs(nthreads);
  auto work = [&](int t) {
    for (int b = t; b < B; b += nthreads)
      if (!parse_one(records[b], lengths[b], b, plan, n_features, errs[t])) return;
  };
  if (nthreads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  for (int t = 0; t < nthreads; ++t)
    if (errs[t].set) {
      t2r::set_error("%s", errs[t].msg);
      return T2R_ERR_PARSE;
    }
  return T2R_OK;
}
