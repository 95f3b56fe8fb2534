// jpeg.cu — device half of the split JPEG decoder (see jpeg_host.cc): dequantisation + ISLOW inverse DCT
// into per-component planes, then fancy chroma upsampling + YCbCr -> RGB.  Integer arithmetic identical
// to libjpeg's jidctint.c / jdsample.c / jdcolor.c (restated in oracle/jpeg.py, which is pinned
// bit-exactly against libjpeg-turbo), so the output equals tf.image.decode_image's
// (utils/tfdata.py:426-484) for the supported streams.
#include <algorithm>

#include "common.cuh"

namespace t2r {

namespace {
constexpr int CONST_BITS = 13, PASS1_BITS = 2;
constexpr int F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270;
constexpr int F_0_899976223 = 7373, F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137;
constexpr int F_1_961570560 = 16069, F_2_053119869 = 16819, F_2_562915447 = 20995, F_3_072711026 = 25172;

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// One 1-D pass of jpeg_idct_islow: in[8] -> out[8], outputs descaled by `shift`.
__device__ __forceinline__ void idct8(const int (&d)[8], int (&o)[8], int shift) {
  int z2 = d[2], z3 = d[6];
  int z1 = (z2 + z3) * F_0_541196100;
  const int t2e = z1 + z3 * (-F_1_847759065);
  const int t3e = z1 + z2 * F_0_765366865;
  const int t0e = (d[0] + d[4]) << CONST_BITS;
  const int t1e = (d[0] - d[4]) << CONST_BITS;
  const int tmp10 = t0e + t3e, tmp13 = t0e - t3e, tmp11 = t1e + t2e, tmp12 = t1e - t2e;
  int tmp0 = d[7], tmp1 = d[5], tmp2 = d[3], tmp3 = d[1];
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
  o[0] = descale(tmp10 + tmp3, shift); o[7] = descale(tmp10 - tmp3, shift);
  o[1] = descale(tmp11 + tmp2, shift); o[6] = descale(tmp11 - tmp2, shift);
  o[2] = descale(tmp12 + tmp1, shift); o[5] = descale(tmp12 - tmp1, shift);
  o[3] = descale(tmp13 + tmp0, shift); o[4] = descale(tmp13 - tmp0, shift);
}
}  // namespace

struct JpegGeom {
  int width, height, ncomp, hmax, vmax;
  int h[3], v[3], tq[3];
  int bw[3], bh[3];            // blocks per row / column of each component plane
  long long coef_offset[3];
  long long coef_stride, coef_count;
};

// One thread per 8x8 block: dequantise, column pass, row pass, store 64 samples into the plane.
__global__ void __launch_bounds__(128) jpeg_idct_kernel(const int16_t* __restrict__ coef, const uint16_t* __restrict__ qt,
                                                        uint8_t* __restrict__ planes, JpegGeom g, int comp,
                                                        long long blocks_per_image, long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long img = i / blocks_per_image;
  const long long blk = i - img * blocks_per_image;
  const int by = int(blk / g.bw[comp]), bx = int(blk - (long long)by * g.bw[comp]);
  const int16_t* c = coef + img * g.coef_stride + g.coef_offset[comp] + blk * 64;
  const uint16_t* q = qt + (img * 4 + g.tq[comp]) * 64;
  int ws[8][8];
#pragma unroll
  for (int col = 0; col < 8; ++col) {
    int d[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) d[r] = int(c[r * 8 + col]) * int(q[r * 8 + col]);
    idct8(d, o, CONST_BITS - PASS1_BITS);
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[r][col] = o[r];
  }
  const int pw = g.bw[comp] * 8;
  uint8_t* dst = planes + img * g.coef_count + g.coef_offset[comp] + ((long long)by * 8) * pw + bx * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int o[8];
    idct8(ws[r], o, CONST_BITS + PASS1_BITS + 3);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo |= uint32_t(min(max(o[k] + 128, 0), 255)) << (8 * k);
      hi |= uint32_t(min(max(o[4 + k] + 128, 0), 255)) << (8 * k);
    }
    *reinterpret_cast<uint2*>(dst + (long long)r * pw) = make_uint2(lo, hi);
  }
}

// Fancy upsampling of one chroma sample position (jdsample.c), sampled at full-resolution pixel (y, x).
__device__ __forceinline__ int chroma_at(const uint8_t* __restrict__ pl, int pw, int cw, int ch, int hs, int vs, int y,
                                         int x) {
  if (hs == 1) return pl[(long long)y * pw + x];           // 4:4:4 (vs is 1 too)
  const int cx = x >> 1;
  if (vs == 1) {                                            // h2v1_fancy_upsample
    const int v0 = pl[(long long)y * pw + cx];
    if (x & 1) {
      if (cx == cw - 1) return v0;
      return (3 * v0 + pl[(long long)y * pw + cx + 1] + 2) >> 2;
    }
    if (cx == 0) return v0;
    return (3 * v0 + pl[(long long)y * pw + cx - 1] + 1) >> 2;
  }
  // h2v2_fancy_upsample: vertical 3:1 with the nearer row (edge rows replicate), then horizontal
  const int cy = y >> 1;
  const int oy = (y & 1) ? min(cy + 1, ch - 1) : max(cy - 1, 0);
  const uint8_t* r0 = pl + (long long)cy * pw;
  const uint8_t* r1 = pl + (long long)oy * pw;
  const int col = 3 * r0[cx] + r1[cx];
  if (x & 1) {
    if (cx == cw - 1) return (4 * col + 7) >> 4;
    return (3 * col + 3 * r0[cx + 1] + r1[cx + 1] + 7) >> 4;
  }
  if (cx == 0) return (4 * col + 8) >> 4;
  return (3 * col + 3 * r0[cx - 1] + r1[cx - 1] + 8) >> 4;
}

__global__ void __launch_bounds__(256) jpeg_color_kernel(const uint8_t* __restrict__ planes, uint8_t* __restrict__ out,
                                                         JpegGeom g, int channels, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = int(i % g.width);
    const long long r = i / g.width;
    const int y = int(r % g.height);
    const long long img = r / g.height;
    const uint8_t* base = planes + img * g.coef_count;
    const int yy = base[g.coef_offset[0] + (long long)y * (g.bw[0] * 8) + x];
    uint8_t* o = out + i * channels;
    if (g.ncomp == 1 || channels == 1) {
      if (channels == 1) { o[0] = uint8_t(yy); } else { o[0] = o[1] = o[2] = uint8_t(yy); }
      continue;
    }
    const int hs = g.hmax / g.h[1], vs = g.vmax / g.v[1];
    const int cw = (g.width * g.h[1] + g.hmax - 1) / g.hmax;     // downsampled_width / height
    const int ch = (g.height * g.v[1] + g.vmax - 1) / g.vmax;
    const int cb = chroma_at(base + g.coef_offset[1], g.bw[1] * 8, cw, ch, hs, vs, y, x) - 128;
    const int cr = chroma_at(base + g.coef_offset[2], g.bw[2] * 8, cw, ch, hs, vs, y, x) - 128;
    // jdcolor.c tables, SCALEBITS = 16
    const int rr = yy + ((91881 * cr + 32768) >> 16);
    const int gg = yy + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
    const int bb = yy + ((116130 * cb + 32768) >> 16);
    o[0] = uint8_t(min(max(rr, 0), 255));
    o[1] = uint8_t(min(max(gg, 0), 255));
    o[2] = uint8_t(min(max(bb, 0), 255));
  }
}

}  // namespace t2r

using namespace t2r;

extern "C" int32_t t2r_jpeg_idct_color(const int16_t* coef, const uint16_t* qt, const T2RJpegInfo* geom, uint8_t* planes,
                                       uint8_t* out, int32_t B, int64_t coef_stride, int32_t channels, void* stream) {
  T2R_CHECK_ARG(coef && qt && geom && planes && out && B > 0 && geom->struct_size == sizeof(T2RJpegInfo),
                "jpeg_idct_color: bad args");
  T2R_CHECK_ARG((channels == 1 || channels == 3) && (geom->ncomp == 1 || geom->ncomp == 3) &&
                    coef_stride >= geom->coef_count, "jpeg_idct_color: bad geometry");
  JpegGeom g;
  g.width = geom->width; g.height = geom->height; g.ncomp = geom->ncomp; g.hmax = geom->hmax; g.vmax = geom->vmax;
  g.coef_stride = coef_stride; g.coef_count = geom->coef_count;
  for (int c = 0; c < 3; ++c) {
    g.h[c] = geom->h[c]; g.v[c] = geom->v[c]; g.tq[c] = geom->tq[c];
    g.bw[c] = geom->mcux * geom->h[c]; g.bh[c] = geom->mcuy * geom->v[c];
    g.coef_offset[c] = geom->coef_offset[c];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ncomp_needed = (channels == 1) ? 1 : g.ncomp;   // luma only for grey output
  for (int c = 0; c < ncomp_needed; ++c) {
    const long long bpi = (long long)g.bw[c] * g.bh[c];
    const long long total = bpi * B;
    jpeg_idct_kernel<<<unsigned((total + 127) / 128), 128, 0, st>>>(coef, qt, planes, g, c, bpi, total);
    T2R_LAUNCH_OK();
  }
  const long long total = (long long)B * g.width * g.height;
  const int blocks = int(std::min<long long>((total + 255) / 256, 148LL * 32));
  jpeg_color_kernel<<<blocks, 256, 0, st>>>(planes, out, g, channels, total);
  T2R_LAUNCH_OK();
  return T2R_OK;
}
