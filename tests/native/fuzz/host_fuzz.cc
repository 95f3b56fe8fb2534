// AddressSanitizer fuzz harness for the host-side parsers of untrusted bytes (csrc/jpeg_host.cc, csrc/host_io.cc):
// JPEG header parsing + Huffman entropy decoding, TFRecord indexing, tf.Example / SequenceExample parsing.
// Mutates valid seed inputs (byte flips, truncation, insertion, marker injection) and calls the C-ABI; the only
// acceptable outcomes are T2R_OK or an error status - any out-of-bounds access aborts under ASan.
//
//   g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fwrapv -std=c++17 -pthread \
//       tests/native/fuzz/host_fuzz.cc tensor2robot_b200/csrc/jpeg_host.cc tensor2robot_b200/csrc/host_io.cc \
//       -o /tmp/host_fuzz && /tmp/host_fuzz <seed dir> <iterations>
// (scripts/host_fuzz.sh writes the seed files and runs it.)
#include <dirent.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <random>
#include <string>
#include <vector>

#include "../../../include/t2r_b200.h"

namespace t2r {
static thread_local char g_err[512];
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace t2r
extern "C" const char* t2r_last_error(void) { return t2r::g_err; }

static std::vector<uint8_t> read_file(const std::string& path) {
  std::vector<uint8_t> out;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return out;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(size_t(n));
  if (n && fread(out.data(), 1, size_t(n), f) != size_t(n)) out.clear();
  fclose(f);
  return out;
}

static void mutate(std::vector<uint8_t>& b, std::mt19937& rng) {
  if (b.empty()) return;
  switch (rng() % 5) {
    case 0:
      for (unsigned i = 0, n = 1 + rng() % 6; i < n; ++i) b[rng() % b.size()] = uint8_t(rng());
      break;
    case 1:
      b.resize(1 + rng() % b.size());
      break;
    case 2: {
      const size_t at = rng() % b.size();
      std::vector<uint8_t> ins(1 + rng() % 20);
      for (auto& x : ins) x = uint8_t(rng());
      b.insert(b.begin() + long(at), ins.begin(), ins.end());
      break;
    }
    case 3:
      b[rng() % std::min<size_t>(b.size(), 700)] = 0xFF;
      break;
    default: {   // overwrite a length-like field with an extreme value
      const size_t at = rng() % b.size();
      const uint8_t v = (rng() & 1) ? 0xFF : 0x00;
      for (size_t i = at; i < std::min(b.size(), at + 1 + rng() % 4); ++i) b[i] = v;
    }
  }
}

static long fuzz_jpeg(const std::vector<uint8_t>& in) {
  T2RJpegInfo info;
  if (t2r_jpeg_parse(in.data(), in.size(), &info) != 0) return 0;
  std::vector<int16_t> coef(static_cast<size_t>(info.coef_count), 0);
  const uint8_t* ptr = in.data();
  const uint64_t len = in.size();
  T2RJpegInfo out;
  long ok = t2r_jpeg_entropy_decode_batch(&ptr, &len, 1, &out, coef.data(), info.coef_count) == 0;
  // the complete host decoder, three copies at once so that the worker pool runs too
  if (int64_t(info.width) * info.height <= (1 << 20)) {
    const uint8_t* ptrs[3] = {ptr, ptr, ptr};
    const uint64_t lens[3] = {len, len, len};
    std::vector<uint8_t> rgb(size_t(3) * info.width * info.height * 3);
    ok += t2r_jpeg_decode_host_batch(ptrs, lens, 3, info.height, info.width, 3, rgb.data()) == 0;
    ok += t2r_jpeg_decode_host_batch(ptrs, lens, 2, info.height, info.width, 1, rgb.data()) == 0;
  }
  return ok;
}

static long fuzz_records(const std::vector<uint8_t>& file, bool sequence) {
  const int64_t n = t2r_tfrecord_index(file.data(), file.size(), nullptr, nullptr, 0, 0);   // no CRC: reach the parser
  if (n <= 0) return 0;
  std::vector<uint64_t> off(static_cast<size_t>(n), 0), len(static_cast<size_t>(n), 0);
  if (t2r_tfrecord_index(file.data(), file.size(), off.data(), len.data(), n, 0) != n) return 0;
  const int B = int(std::min<int64_t>(n, 4));
  std::vector<const uint8_t*> recs;
  std::vector<uint64_t> lens;
  for (int i = 0; i < B; ++i) {
    recs.push_back(file.data() + off[size_t(i)]);
    lens.push_back(len[size_t(i)]);
  }
  std::vector<float> pose(static_cast<size_t>(B) * 2, 0.f), reward(static_cast<size_t>(B), 0.f);
  std::vector<int64_t> ids(static_cast<size_t>(B) * 3, 0);
  std::vector<const uint8_t*> img(static_cast<size_t>(B), nullptr);
  std::vector<uint64_t> img_len(static_cast<size_t>(B), 0);
  T2RFeaturePlan plan[4];
  memset(plan, 0, sizeof(plan));
  plan[0].key = "pose"; plan[0].dtype = T2R_DT_FLOAT; plan[0].count = 2; plan[0].required = 0; plan[0].dst = pose.data();
  plan[1].key = "reward"; plan[1].dtype = T2R_DT_FLOAT; plan[1].count = 1; plan[1].required = 0; plan[1].dst = reward.data();
  plan[2].key = "ids"; plan[2].dtype = T2R_DT_INT64; plan[2].count = 3; plan[2].required = 0; plan[2].dst = ids.data();
  plan[3].key = "state/image"; plan[3].dtype = T2R_DT_BYTES; plan[3].count = 1; plan[3].required = 0;
  plan[3].dst = img.data(); plan[3].dst_len = img_len.data();
  if (!sequence) {
    long ok = t2r_example_parse_batch(recs.data(), lens.data(), B, plan, 4) == 0;
    // variable-length features (count < 0: at most |count| values per row, padded) and a required key
    std::vector<float> vpose(static_cast<size_t>(B) * 3, 0.f);
    std::vector<int64_t> vids(static_cast<size_t>(B) * 2, 0);
    std::vector<uint64_t> vpose_n(static_cast<size_t>(B), 0), vids_n(static_cast<size_t>(B), 0);
    T2RFeaturePlan vplan[3];
    memset(vplan, 0, sizeof(vplan));
    vplan[0].key = "pose"; vplan[0].dtype = T2R_DT_FLOAT; vplan[0].count = -3; vplan[0].dst = vpose.data();
    vplan[0].dst_len = vpose_n.data(); vplan[0].pad_float = -1.f;
    vplan[1].key = "ids"; vplan[1].dtype = T2R_DT_INT64; vplan[1].count = -2; vplan[1].dst = vids.data();
    vplan[1].dst_len = vids_n.data(); vplan[1].pad_int64 = -7;
    vplan[2].key = "reward"; vplan[2].dtype = T2R_DT_FLOAT; vplan[2].count = 1; vplan[2].required = 1; vplan[2].dst = reward.data();
    ok += t2r_example_parse_batch(recs.data(), lens.data(), B, vplan, 3) == 0;
    return ok;
  }
  const int T = 5;
  std::vector<float> spose(static_cast<size_t>(B) * T * 2, 0.f);
  std::vector<int64_t> seq_len(static_cast<size_t>(B), 0);
  T2RFeaturePlan splan[1];
  memset(splan, 0, sizeof(splan));
  splan[0].key = "pose"; splan[0].dtype = T2R_DT_FLOAT; splan[0].count = 2; splan[0].required = 0; splan[0].dst = spose.data();
  return t2r_sequence_example_parse_batch(recs.data(), lens.data(), B, splan, 1, T, seq_len.data()) == 0;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: host_fuzz <seed dir> <iterations>\n");
    return 2;
  }
  std::vector<std::pair<std::string, std::vector<uint8_t>>> seeds;
  DIR* d = opendir(argv[1]);
  if (!d) return 2;
  while (dirent* e = readdir(d)) {
    const std::string name = e->d_name;
    if (name.size() < 4) continue;
    auto data = read_file(std::string(argv[1]) + "/" + name);
    if (!data.empty()) seeds.emplace_back(name, std::move(data));
  }
  closedir(d);
  if (seeds.empty()) return 2;
  const long iters = atol(argv[2]);
  std::mt19937 rng(12345);
  long accepted = 0;
  for (long it = 0; it < iters; ++it) {
    const auto& seed = seeds[size_t(it) % seeds.size()];
    std::vector<uint8_t> in = seed.second;
    for (unsigned m = 0, n = 1 + rng() % 3; m < n; ++m) mutate(in, rng);
    const bool is_jpeg = seed.first.find(".jpg") != std::string::npos;
    if (is_jpeg) accepted += fuzz_jpeg(in);
    else accepted += fuzz_records(in, seed.first.find("seq") != std::string::npos);
  }
  printf("host_fuzz: %ld inputs over %zu seeds, %ld accepted, no memory error\n", iters, seeds.size(), accepted);
  return 0;
}
