// image.cu — fused crop + uint8->float + photometric distortion + clip, and legacy bilinear resize.
//
// Reference call sites: research/qtopt/t2r_models.py:297-308 (crop -> convert_image_dtype ->
// ApplyPhotometricImageDistortions), preprocessors/image_transformations.py:176-264,
// preprocessors/distortion.py:56-107.  TF op semantics restated (SURVEY §8c-3,8,9):
//   convert_image_dtype(u8->f32): x * (1/255)
//   adjust_brightness: x + delta
//   adjust_saturation: RGB->HSV, s = clamp(s*f, 0, 1), HSV->RGB
//   adjust_hue:        RGB->HSV, h = frac(h + delta), HSV->RGB
//   adjust_contrast:   (x - mean_c) * f + mean_c, mean per image and channel
//   noise:             x + N(0, sigma)
//   clip_by_value(0, 1)
// HBM-bound: one read of the cropped u8 window, one write of the bf16/f32 result.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "philox.cuh"

namespace t2r {

__device__ __forceinline__ void rgb_to_hsv(float r, float g, float b, float* h, float* s, float* v) {
  const float mx = fmaxf(r, fmaxf(g, b)), mn = fminf(r, fminf(g, b));
  const float range = mx - mn;
  *v = mx;
  *s = mx > 0.f ? range / mx : 0.f;
  float hh = 0.f;
  if (range > 0.f) {
    const float norm = 1.0f / (6.0f * range);
    if (r == mx) hh = norm * (g - b);
    else if (g == mx) hh = norm * (b - r) + 2.0f / 6.0f;
    else hh = norm * (r - g) + 4.0f / 6.0f;
    if (hh < 0.f) hh += 1.0f;
  }
  *h = hh;
}

__device__ __forceinline__ void hsv_to_rgb(float h, float s, float v, float* r, float* g, float* b) {
  const float c = s * v, m = v - c;
  const float dh = h * 6.0f;
  const int cat = int(dh);
  float fm = dh;
  while (fm <= 0.f) fm += 2.0f;
  while (fm >= 2.0f) fm -= 2.0f;
  const float x = c * (1.0f - fabsf(fm - 1.0f));
  float rr = 0.f, gg = 0.f, bb = 0.f;
  switch (cat) {
    case 0: rr = c; gg = x; break;
    case 1: rr = x; gg = c; break;
    case 2: gg = c; bb = x; break;
    case 3: gg = x; bb = c; break;
    case 4: rr = x; bb = c; break;
    case 5: rr = c; bb = x; break;
    default: break;
  }
  *r = rr + m; *g = gg + m; *b = bb + m;
}

// Everything up to (not including) contrast.
__device__ __forceinline__ void pre_contrast(const T2RDistortParams& pr, float& r, float& g, float& b) {
  if (pr.brightness_delta != 0.f) {
    r += pr.brightness_delta; g += pr.brightness_delta; b += pr.brightness_delta;
  }
  if (pr.saturation_scale != 1.f) {
    float h, s, v;
    rgb_to_hsv(r, g, b, &h, &s, &v);
    s = fminf(fmaxf(s * pr.saturation_scale, 0.f), 1.f);
    hsv_to_rgb(h, s, v, &r, &g, &b);
  }
  if (pr.hue_delta != 0.f) {
    float h, s, v;
    rgb_to_hsv(r, g, b, &h, &s, &v);
    h += pr.hue_delta;
    h -= floorf(h);
    hsv_to_rgb(h, s, v, &r, &g, &b);
  }
}

// Pixel loaders: uint8 frames are converted like tf.image.convert_image_dtype (x * 1/255); float frames
// (already converted, e.g. after the BC-Z resize) are taken as they are.
__device__ __forceinline__ void load_rgb(const uint8_t* px, float& r, float& g, float& b) {
  r = float(px[0]) * (1.0f / 255.0f); g = float(px[1]) * (1.0f / 255.0f); b = float(px[2]) * (1.0f / 255.0f);
}
__device__ __forceinline__ void load_rgb(const float* px, float& r, float& g, float& b) {
  r = px[0]; g = px[1]; b = px[2];
}

template <typename SrcT>
__global__ void __launch_bounds__(256) image_mean_kernel(const SrcT* __restrict__ src,
                                                         const T2RDistortParams* __restrict__ params,
                                                         float* chan_mean, int H, int W, int h, int w) {
  const int n = blockIdx.y;
  const T2RDistortParams pr = params[n];
  const SrcT* img = src + (size_t)n * H * W * 3;
  float acc[3] = {0.f, 0.f, 0.f};
  const int total = h * w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / w, x = i - y * w;
    const SrcT* px = img + ((size_t)(y + pr.crop_y) * W + (x + pr.crop_x)) * 3;
    float r, g, b;
    load_rgb(px, r, g, b);
    pre_contrast(pr, r, g, b);
    acc[0] += r; acc[1] += g; acc[2] += b;
  }
  __shared__ float sm[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = acc[c];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[c][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int j = 0; j < 8; ++j) s += sm[threadIdx.x][j];
    atomicAdd(chan_mean + n * 3 + threadIdx.x, s / float(total));
  }
}

template <typename SrcT, bool OUT_F32>
__global__ void __launch_bounds__(256) crop_convert_distort_kernel(
    const SrcT* __restrict__ src, void* __restrict__ dst, const T2RDistortParams* __restrict__ params,
    const float* __restrict__ chan_mean, int H, int W, int h, int w, int use_contrast, uint64_t seed,
    uint64_t offset) {
  const int n = blockIdx.y;
  const T2RDistortParams pr = params[n];
  const SrcT* img = src + (size_t)n * H * W * 3;
  const int total = h * w;
  float m[3] = {0.f, 0.f, 0.f};
  if (use_contrast) { m[0] = chan_mean[n * 3]; m[1] = chan_mean[n * 3 + 1]; m[2] = chan_mean[n * 3 + 2]; }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / w, x = i - y * w;
    const SrcT* px = img + ((size_t)(y + pr.crop_y) * W + (x + pr.crop_x)) * 3;
    float r, g, b;
    load_rgb(px, r, g, b);
    pre_contrast(pr, r, g, b);
    if (use_contrast && pr.contrast_scale != 1.f) {
      r = (r - m[0]) * pr.contrast_scale + m[0];
      g = (g - m[1]) * pr.contrast_scale + m[1];
      b = (b - m[2]) * pr.contrast_scale + m[2];
    }
    if (pr.noise_stddev != 0.f) {
      const Philox4 rnd = philox4x32_10(seed, (uint64_t)n * total + i, offset);
      float z0, z1, z2, z3;
      box_muller(rnd.v[0], rnd.v[1], &z0, &z1);
      box_muller(rnd.v[2], rnd.v[3], &z2, &z3);
      r += pr.noise_stddev * z0; g += pr.noise_stddev * z1; b += pr.noise_stddev * z2;
    }
    r = fminf(fmaxf(r, 0.f), 1.f); g = fminf(fmaxf(g, 0.f), 1.f); b = fminf(fmaxf(b, 0.f), 1.f);
    const size_t o = ((size_t)n * total + i) * 3;
    if (OUT_F32) {
      float* d = static_cast<float*>(dst) + o;
      d[0] = r; d[1] = g; d[2] = b;
    } else {
      __nv_bfloat16* d = static_cast<__nv_bfloat16*>(dst) + o;
      d[0] = __float2bfloat16_rn(r); d[1] = __float2bfloat16_rn(g); d[2] = __float2bfloat16_rn(b);
    }
  }
}


// uint8 frames, vectorised: a block converts kVecRows crop rows of one image.
//   1. the 3*w source bytes of every row (arbitrary byte alignment: crop_x * 3) are fetched with aligned 16-byte loads
//      into shared memory (one extra vector per row covers the misalignment);
//   2. a thread converts 8 consecutive pixels: 24 bytes out of shared memory through 7 aligned words + funnel shifts,
//      the same per-pixel arithmetic as the scalar kernel, results staged in shared memory;
//   3. the staged rows leave as consecutive 16-byte stores (w % 8 == 0 makes every output row a multiple of 48 B).
// HBM traffic is the algorithmic 3 B read + 6 B (bf16) / 12 B (fp32) written per pixel; the scalar kernel issued three
// byte loads and three 2-byte stores per pixel and reached 0.25 of the copy bandwidth.
constexpr int kVecRows = 4;
constexpr int kVecRowsDirect = 8;

// DIRECT: no output staging - every thread stores its 8 pixels (48 / 96 contiguous bytes) straight to global memory as
// 16-byte vectors (a warp covers one contiguous 1.5 / 3 KB span), one barrier per block, twice the rows per block.
template <bool OUT_F32, bool DIRECT>
__global__ void __launch_bounds__(256) crop_convert_distort_vec_kernel(
    const uint8_t* __restrict__ src, void* __restrict__ dst, const T2RDistortParams* __restrict__ params,
    const float* __restrict__ chan_mean, int H, int W, int h, int w, int use_contrast, uint64_t seed, uint64_t offset,
    int row_pitch /* shared bytes per source row, multiple of 16 */) {
  extern __shared__ uint4 smem_v[];
  uint8_t* in_sm = reinterpret_cast<uint8_t*>(smem_v);
  constexpr int kOutBytes = OUT_F32 ? 12 : 6;   // per pixel
  constexpr int kRows = DIRECT ? kVecRowsDirect : kVecRows;
  uint8_t* out_sm = in_sm + kRows * row_pitch;
  const int n = blockIdx.y;
  const int y0 = blockIdx.x * kRows;
  const int rows = min(kRows, h - y0);
  const T2RDistortParams pr = params[n];
  const uint8_t* img = src + (size_t)n * H * W * 3;
  const uint8_t* img_end = src + (size_t)gridDim.y * H * W * 3;
  const int total = h * w;
  float m[3] = {0.f, 0.f, 0.f};
  if (use_contrast) { m[0] = chan_mean[n * 3]; m[1] = chan_mean[n * 3 + 1]; m[2] = chan_mean[n * 3 + 2]; }

  // ---- 1. source rows -> shared memory (aligned 16-byte vectors) ----
  const int vecs_per_row = row_pitch / 16;
  for (int t = threadIdx.x; t < rows * vecs_per_row; t += blockDim.x) {
    const int r = t / vecs_per_row, v = t - r * vecs_per_row;
    const uint8_t* start = img + ((size_t)(y0 + r + pr.crop_y) * W + pr.crop_x) * 3;
    const uint8_t* a0 = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(start) & ~uintptr_t(15));
    const uint8_t* p = a0 + (size_t)v * 16;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (p >= start + 3 * w) {
      // past the row: nothing to fetch
    } else if (p + 16 <= img_end && p >= src) {
      q = __ldg(reinterpret_cast<const uint4*>(p));
    } else {   // the vector straddles the end (or the start) of the buffer: byte loads of what exists
      uint8_t b[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) b[k] = (p + k >= src && p + k < img_end) ? p[k] : uint8_t(0);
      memcpy(&q, b, 16);
    }
    reinterpret_cast<uint4*>(in_sm + r * row_pitch)[v] = q;
  }
  __syncthreads();

  // ---- 2. 8 pixels per thread ----
  const int groups_per_row = w / 8;
  for (int t = threadIdx.x; t < rows * groups_per_row; t += blockDim.x) {
    const int r = t / groups_per_row, gi = t - r * groups_per_row;
    const uint8_t* start = img + ((size_t)(y0 + r + pr.crop_y) * W + pr.crop_x) * 3;
    const int off = int(reinterpret_cast<uintptr_t>(start) & 15) + gi * 24;
    const uint32_t* words = reinterpret_cast<const uint32_t*>(in_sm + r * row_pitch) + (off >> 2);
    const uint32_t sh = uint32_t(off & 3) * 8u;
    uint32_t wv[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) wv[k] = words[k];
    uint32_t px[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) px[k] = __funnelshift_r(wv[k], wv[k + 1], sh);
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(px);
    float outv[24];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float r_ = float(bytes[3 * j]) * (1.0f / 255.0f), g_ = float(bytes[3 * j + 1]) * (1.0f / 255.0f),
            b_ = float(bytes[3 * j + 2]) * (1.0f / 255.0f);
      pre_contrast(pr, r_, g_, b_);
      if (use_contrast && pr.contrast_scale != 1.f) {
        r_ = (r_ - m[0]) * pr.contrast_scale + m[0];
        g_ = (g_ - m[1]) * pr.contrast_scale + m[1];
        b_ = (b_ - m[2]) * pr.contrast_scale + m[2];
      }
      if (pr.noise_stddev != 0.f) {
        const int i = (y0 + r) * w + gi * 8 + j;
        const Philox4 rnd = philox4x32_10(seed, (uint64_t)n * total + i, offset);
        float z0, z1, z2, z3;
        box_muller(rnd.v[0], rnd.v[1], &z0, &z1);
        box_muller(rnd.v[2], rnd.v[3], &z2, &z3);
        r_ += pr.noise_stddev * z0; g_ += pr.noise_stddev * z1; b_ += pr.noise_stddev * z2;
      }
      outv[3 * j] = fminf(fmaxf(r_, 0.f), 1.f);
      outv[3 * j + 1] = fminf(fmaxf(g_, 0.f), 1.f);
      outv[3 * j + 2] = fminf(fmaxf(b_, 0.f), 1.f);
    }
    uint8_t* o = DIRECT ? static_cast<uint8_t*>(dst) + ((size_t)n * total + (size_t)(y0 + r) * w + gi * 8) * kOutBytes
                        : out_sm + ((size_t)r * w + gi * 8) * kOutBytes;
    if (OUT_F32) {
#pragma unroll
      for (int k = 0; k < 6; ++k)
        reinterpret_cast<float4*>(o)[k] = make_float4(outv[4 * k], outv[4 * k + 1], outv[4 * k + 2], outv[4 * k + 3]);
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        reinterpret_cast<uint4*>(o)[k] = make_uint4(pack_bf16(outv[8 * k], outv[8 * k + 1]), pack_bf16(outv[8 * k + 2], outv[8 * k + 3]),
                                                    pack_bf16(outv[8 * k + 4], outv[8 * k + 5]), pack_bf16(outv[8 * k + 6], outv[8 * k + 7]));
    }
  }
  if (DIRECT) return;
  __syncthreads();

  // ---- 3. staged rows -> global, consecutive 16-byte stores ----
  const int out_vecs = rows * w * kOutBytes / 16;
  uint4* gdst = reinterpret_cast<uint4*>(static_cast<uint8_t*>(dst) + ((size_t)n * total + (size_t)y0 * w) * kOutBytes);
  for (int t = threadIdx.x; t < out_vecs; t += blockDim.x) gdst[t] = reinterpret_cast<const uint4*>(out_sm)[t];
}

// Persistent, double-buffered variant: a block walks over (image, 8-row group) items; the aligned 16-byte source vectors
// of the NEXT item are in flight (cp.async into the other shared buffer) while the current one is converted and stored,
// so the load latency that bounded the one-shot kernels (load phase, barrier, convert / store phase) is hidden.
constexpr int kPipeRows = 8;

template <bool OUT_F32>
__device__ __forceinline__ void crop_convert_group(const uint8_t* in_row, int off, const T2RDistortParams& pr, const float* m,
                                                   int use_contrast, uint64_t seed, uint64_t offset, uint64_t pixel_index0,
                                                   uint8_t* o) {
  const uint32_t* words = reinterpret_cast<const uint32_t*>(in_row) + (off >> 2);
  const uint32_t sh = uint32_t(off & 3) * 8u;
  uint32_t wv[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) wv[k] = words[k];
  uint32_t px[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) px[k] = __funnelshift_r(wv[k], wv[k + 1], sh);
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(px);
  float outv[24];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float r_ = float(bytes[3 * j]) * (1.0f / 255.0f), g_ = float(bytes[3 * j + 1]) * (1.0f / 255.0f),
          b_ = float(bytes[3 * j + 2]) * (1.0f / 255.0f);
    pre_contrast(pr, r_, g_, b_);
    if (use_contrast && pr.contrast_scale != 1.f) {
      r_ = (r_ - m[0]) * pr.contrast_scale + m[0];
      g_ = (g_ - m[1]) * pr.contrast_scale + m[1];
      b_ = (b_ - m[2]) * pr.contrast_scale + m[2];
    }
    if (pr.noise_stddev != 0.f) {
      const Philox4 rnd = philox4x32_10(seed, pixel_index0 + j, offset);
      float z0, z1, z2, z3;
      box_muller(rnd.v[0], rnd.v[1], &z0, &z1);
      box_muller(rnd.v[2], rnd.v[3], &z2, &z3);
      r_ += pr.noise_stddev * z0; g_ += pr.noise_stddev * z1; b_ += pr.noise_stddev * z2;
    }
    outv[3 * j] = fminf(fmaxf(r_, 0.f), 1.f);
    outv[3 * j + 1] = fminf(fmaxf(g_, 0.f), 1.f);
    outv[3 * j + 2] = fminf(fmaxf(b_, 0.f), 1.f);
  }
  if (OUT_F32) {
#pragma unroll
    for (int k = 0; k < 6; ++k)
      reinterpret_cast<float4*>(o)[k] = make_float4(outv[4 * k], outv[4 * k + 1], outv[4 * k + 2], outv[4 * k + 3]);
  } else {
#pragma unroll
    for (int k = 0; k < 3; ++k)
      reinterpret_cast<uint4*>(o)[k] = make_uint4(pack_bf16(outv[8 * k], outv[8 * k + 1]), pack_bf16(outv[8 * k + 2], outv[8 * k + 3]),
                                                  pack_bf16(outv[8 * k + 4], outv[8 * k + 5]), pack_bf16(outv[8 * k + 6], outv[8 * k + 7]));
  }
}

template <bool OUT_F32>
__global__ void __launch_bounds__(256) crop_convert_distort_pipe_kernel(
    const uint8_t* __restrict__ src, void* __restrict__ dst, const T2RDistortParams* __restrict__ params,
    const float* __restrict__ chan_mean, int N, int H, int W, int h, int w, int use_contrast, uint64_t seed, uint64_t offset,
    int row_pitch) {
  extern __shared__ uint4 smem_v[];
  uint8_t* bufs[2] = {reinterpret_cast<uint8_t*>(smem_v), reinterpret_cast<uint8_t*>(smem_v) + kPipeRows * row_pitch};
  constexpr int kOutBytes = OUT_F32 ? 12 : 6;
  const int groups_y = (h + kPipeRows - 1) / kPipeRows;
  const int items = N * groups_y;
  const uint8_t* src_end = src + (size_t)N * H * W * 3;
  const int vecs_per_row = row_pitch / 16;
  const int total = h * w;

  auto issue = [&](int item, uint8_t* buf) {      // the source vectors of one item -> shared memory, asynchronously
    const int n = item / groups_y, y0 = (item - n * groups_y) * kPipeRows;
    const int rows = min(kPipeRows, h - y0);
    const int cy = params[n].crop_y, cx = params[n].crop_x;
    const uint8_t* img = src + (size_t)n * H * W * 3;
    for (int t = threadIdx.x; t < rows * vecs_per_row; t += blockDim.x) {
      const int r = t / vecs_per_row, v = t - r * vecs_per_row;
      const uint8_t* start = img + ((size_t)(y0 + r + cy) * W + cx) * 3;
      const uint8_t* a0 = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(start) & ~uintptr_t(15));
      const uint8_t* p = a0 + (size_t)v * 16;
      uint4* d = reinterpret_cast<uint4*>(buf + r * row_pitch) + v;
      if (p >= start + 3 * w) {
        *d = make_uint4(0, 0, 0, 0);
      } else if (p + 16 <= src_end && p >= src) {
        const uint32_t daddr = static_cast<uint32_t>(__cvta_generic_to_shared(d));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(daddr), "l"(p) : "memory");
      } else {   // the vector straddles the end (or the start) of the buffer: byte loads of what exists
        uint8_t b[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) b[k] = (p + k >= src && p + k < src_end) ? p[k] : uint8_t(0);
        uint4 q;
        memcpy(&q, b, 16);
        *d = q;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int item = blockIdx.x;
  if (item >= items) return;
  issue(item, bufs[0]);
  for (int it = 0; item < items; item += gridDim.x, ++it) {
    uint8_t* cur = bufs[it & 1];
    const int next = item + gridDim.x;
    if (next < items) {
      issue(next, bufs[(it + 1) & 1]);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();                               // every thread's vectors of `cur` have landed
    const int n = item / groups_y, y0 = (item - n * groups_y) * kPipeRows;
    const int rows = min(kPipeRows, h - y0);
    const T2RDistortParams pr = params[n];
    float m[3] = {0.f, 0.f, 0.f};
    if (use_contrast) { m[0] = chan_mean[n * 3]; m[1] = chan_mean[n * 3 + 1]; m[2] = chan_mean[n * 3 + 2]; }
    const uint8_t* img = src + (size_t)n * H * W * 3;
    const int groups_per_row = w / 8;
    for (int t = threadIdx.x; t < rows * groups_per_row; t += blockDim.x) {
      const int r = t / groups_per_row, gi = t - r * groups_per_row;
      const uint8_t* start = img + ((size_t)(y0 + r + pr.crop_y) * W + pr.crop_x) * 3;
      const int off = int(reinterpret_cast<uintptr_t>(start) & 15) + gi * 24;
      const size_t pix = (size_t)(y0 + r) * w + gi * 8;
      uint8_t* o = static_cast<uint8_t*>(dst) + ((size_t)n * total + pix) * kOutBytes;
      crop_convert_group<OUT_F32>(cur + r * row_pitch, off, pr, m, use_contrast, seed, offset, (uint64_t)n * total + pix, o);
    }
    __syncthreads();                               // `cur` is free for the loads of the item after next
  }
}

// TF1 tf.image.resize_images(BILINEAR), align_corners=False, legacy sampling: src = dst * scale.
__global__ void __launch_bounds__(256) resize_bilinear_legacy_kernel(const float* __restrict__ src,
                                                                     float* __restrict__ dst, int N, int H,
                                                                     int W, int C, int h, int w) {
  const long long total = (long long)N * h * w * C;
  const float sy = float(H) / float(h), sx = float(W) / float(w);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % C);
    long long r = i / C;
    const int x = int(r % w); r /= w;
    const int y = int(r % h);
    const int n = int(r / h);
    const float fy = y * sy, fx = x * sx;
    const int y0 = int(floorf(fy)), x0 = int(floorf(fx));
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = fy - y0, lx = fx - x0;
    const float* base = src + (size_t)n * H * W * C;
    const float tl = base[((size_t)y0 * W + x0) * C + c], tr = base[((size_t)y0 * W + x1) * C + c];
    const float bl = base[((size_t)y1 * W + x0) * C + c], br = base[((size_t)y1 * W + x1) * C + c];
    const float top = tl + (tr - tl) * lx, bot = bl + (br - bl) * lx;
    dst[i] = top + (bot - top) * ly;
  }
}

// ApplyPhotometricImageDistortionsCheap (image_transformations.py:365-384): per-channel gamma, c ** g_c.
__global__ void __launch_bounds__(256) channel_gamma_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                            long long n, int C, float g0, float g1, float g2, float g3) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    const int c = int(i % C);
    const float g = c == 0 ? g0 : (c == 1 ? g1 : (c == 2 ? g2 : g3));
    dst[i] = powf(src[i], g);
  }
}

// ApplyDepthImageDistortions (image_transformations.py:403-459) for one tensor of the list:
// clip(alpha * x + N(0, sigma), lo, hi); alpha = 1, sigma = 0 is the "noise not applied" branch.
__global__ void __launch_bounds__(256) depth_distort_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                            long long n, float alpha, float sigma, float lo, float hi,
                                                            uint64_t seed, uint64_t offset) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < (n + 3) / 4; i += gridDim.x * 256LL) {
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (sigma != 0.f) {
      const Philox4 rnd = philox4x32_10(seed, (uint64_t)i, offset);
      box_muller(rnd.v[0], rnd.v[1], &z[0], &z[1]);
      box_muller(rnd.v[2], rnd.v[3], &z[2], &z[3]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long j = i * 4 + k;
      if (j < n) dst[j] = fminf(fmaxf(fmaf(alpha, src[j], sigma * z[k]), lo), hi);
    }
  }
}

// Mixup regularisation (research/bcz/model.py:164-172): y[b] = lambda * x[b] + (1 - lambda) * x[B - 1 - b]
// (tf.reverse along the batch axis), fp32 [B, inner].
__global__ void __launch_bounds__(256) mixup_reverse_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int B,
                                                                long long inner, float lambda) {
  const long long total = (long long)B * inner;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const long long b = i / inner, j = i - b * inner;
    y[i] = lambda * x[i] + (1.0f - lambda) * x[(B - 1 - b) * inner + j];
  }
}

}  // namespace t2r

using namespace t2r;

extern "C" int32_t t2r_mixup_reverse_f32(const float* x, float* y, int32_t B, int64_t inner, float lambda, void* stream) {
  T2R_CHECK_ARG(x && y && x != y && B > 0 && inner > 0, "mixup_reverse_f32: bad args (out of place only)");
  const long long total = (long long)B * inner;
  const int grid = int(std::min<long long>((total + 255) / 256, 148LL * 16));
  mixup_reverse_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, B, inner, lambda);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_crop_convert_distort(const uint8_t* src, void* dst, const T2RDistortParams* params,
                                            float* chan_mean, int32_t N, int32_t H, int32_t W, int32_t h,
                                            int32_t w, int32_t out_f32, int32_t use_contrast,
                                            uint64_t seed, uint64_t offset, void* stream) {
  T2R_CHECK_ARG(src && dst && params && N > 0 && h > 0 && w > 0 && h <= H && w <= W,
                "crop_convert_distort: bad args");
  T2R_CHECK_ARG(!use_contrast || chan_mean, "crop_convert_distort: contrast needs chan_mean workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int bx = std::max(1, std::min((h * w + 255) / 256, std::max(1, 148 * 8 / N)));
  if (use_contrast) {
    T2R_CUDA_OK(cudaMemsetAsync(chan_mean, 0, sizeof(float) * 3 * N, st));
    image_mean_kernel<uint8_t><<<dim3(bx, N), 256, 0, st>>>(src, params, chan_mean, H, W, h, w);
    T2R_LAUNCH_OK();
  }
  const int row_pitch = ((3 * w + 15) / 16 + 1) * 16;
  const int out_bytes = out_f32 ? 12 : 6;
  const size_t vec_smem = size_t(kVecRows) * (row_pitch + size_t(w) * out_bytes);
  static const bool no_vec = std::getenv("T2R_DISABLE_VEC_CROP") != nullptr;
  if (!no_vec && w % 8 == 0 && vec_smem <= 48 * 1024 && reinterpret_cast<uintptr_t>(dst) % 16 == 0 &&
      (size_t(h) * w * out_bytes) % 16 == 0) {
    // measured on B200 (472 x 472 crops, batch 512, CUDA events): direct stores 0.390 ms, staged stores 0.515 ms
    static const bool pipe = getenv("T2R_CROP_PIPE") != nullptr && getenv("T2R_CROP_PIPE")[0] == '1';
    if (pipe) {
      const int items = N * ((h + kPipeRows - 1) / kPipeRows);
      const int grid = std::min(items, 148 * 6);
      const size_t smem = size_t(2) * kPipeRows * row_pitch;
      if (out_f32)
        crop_convert_distort_pipe_kernel<true><<<grid, 256, smem, st>>>(src, dst, params, chan_mean, N, H, W, h, w,
                                                                        use_contrast, seed, offset, row_pitch);
      else
        crop_convert_distort_pipe_kernel<false><<<grid, 256, smem, st>>>(src, dst, params, chan_mean, N, H, W, h, w,
                                                                         use_contrast, seed, offset, row_pitch);
      T2R_LAUNCH_OK();
      return T2R_OK;
    }
    static const bool direct = !(getenv("T2R_CROP_DIRECT") != nullptr && getenv("T2R_CROP_DIRECT")[0] == '0');
    if (direct) {
      const dim3 grid((h + kVecRowsDirect - 1) / kVecRowsDirect, N);
      const size_t smem = size_t(kVecRowsDirect) * row_pitch;
      if (out_f32)
        crop_convert_distort_vec_kernel<true, true><<<grid, 256, smem, st>>>(src, dst, params, chan_mean, H, W, h, w,
                                                                             use_contrast, seed, offset, row_pitch);
      else
        crop_convert_distort_vec_kernel<false, true><<<grid, 256, smem, st>>>(src, dst, params, chan_mean, H, W, h, w,
                                                                              use_contrast, seed, offset, row_pitch);
      T2R_LAUNCH_OK();
      return T2R_OK;
    }
    const dim3 grid((h + kVecRows - 1) / kVecRows, N);
    if (out_f32)
      crop_convert_distort_vec_kernel<true, false><<<grid, 256, vec_smem, st>>>(src, dst, params, chan_mean, H, W, h, w,
                                                                                use_contrast, seed, offset, row_pitch);
    else
      crop_convert_distort_vec_kernel<false, false><<<grid, 256, vec_smem, st>>>(src, dst, params, chan_mean, H, W, h, w,
                                                                                 use_contrast, seed, offset, row_pitch);
    T2R_LAUNCH_OK();
    return T2R_OK;
  }
  if (out_f32)
    crop_convert_distort_kernel<uint8_t, true><<<dim3(bx, N), 256, 0, st>>>(src, dst, params, chan_mean, H, W, h, w,
                                                                            use_contrast, seed, offset);
  else
    crop_convert_distort_kernel<uint8_t, false><<<dim3(bx, N), 256, 0, st>>>(src, dst, params, chan_mean, H, W, h, w,
                                                                             use_contrast, seed, offset);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_distort_f32(const float* src, float* dst, const T2RDistortParams* params, float* chan_mean,
                                   int32_t N, int32_t H, int32_t W, int32_t use_contrast, uint64_t seed,
                                   uint64_t offset, void* stream) {
  T2R_CHECK_ARG(src && dst && params && N > 0 && H > 0 && W > 0, "distort_f32: bad args");
  T2R_CHECK_ARG(!use_contrast || chan_mean, "distort_f32: contrast needs chan_mean workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int bx = std::max(1, std::min((H * W + 255) / 256, std::max(1, 148 * 8 / N)));
  if (use_contrast) {
    T2R_CUDA_OK(cudaMemsetAsync(chan_mean, 0, sizeof(float) * 3 * N, st));
    image_mean_kernel<float><<<dim3(bx, N), 256, 0, st>>>(src, params, chan_mean, H, W, H, W);
    T2R_LAUNCH_OK();
  }
  crop_convert_distort_kernel<float, true><<<dim3(bx, N), 256, 0, st>>>(src, dst, params, chan_mean, H, W, H, W,
                                                                        use_contrast, seed, offset);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_resize_bilinear_legacy(const float* src, float* dst, int32_t N, int32_t H, int32_t W,
                                              int32_t C, int32_t h, int32_t w, void* stream) {
  T2R_CHECK_ARG(src && dst && N > 0 && H > 0 && W > 0 && C > 0 && h > 0 && w > 0, "resize: bad args");
  const long long total = (long long)N * h * w * C;
  const int grid = int(std::min<long long>((total + 255) / 256, 148LL * 16));
  resize_bilinear_legacy_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, N, H, W, C, h, w);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_channel_gamma_f32(const float* src, float* dst, int64_t n, int32_t C, float g0, float g1, float g2,
                                         float g3, void* stream) {
  T2R_CHECK_ARG(src && dst && n > 0 && C >= 1 && C <= 4 && n % C == 0, "channel_gamma_f32: bad args (1..4 channels)");
  const int grid = int(std::min<long long>((n + 255) / 256, 148LL * 16));
  channel_gamma_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n, C, g0, g1, g2, g3);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_depth_distort_f32(const float* src, float* dst, int64_t n, float alpha, float noise_stddev, float min_depth,
                                         float max_depth, uint64_t seed, uint64_t offset, void* stream) {
  T2R_CHECK_ARG(src && dst && n > 0 && min_depth <= max_depth && noise_stddev >= 0.f, "depth_distort_f32: bad args");
  const int grid = int(std::min<long long>(((n + 3) / 4 + 255) / 256, 148LL * 16));
  depth_distort_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n, alpha, noise_stddev, min_depth,
                                                                            max_depth, seed, offset);
  T2R_LAUNCH_OK();
  return T2R_OK;
}
