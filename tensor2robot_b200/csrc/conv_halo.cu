// conv_halo.cu — stride-1 KxK convolution (fprop / dgrad) with the im2col done *inside* shared
// memory: one TMA load of a (16+KH-1) x (8+KW-1) pixel halo serves every filter tap.
//
// Why: the tap-table kernel (conv_igemm.cu) re-loads the 128-pixel activation tile once per tap,
// so a 64-channel 3x3 / 5x5 layer moves 9x / 25x its input through L2 -> SMEM and is L2-bound at
// ~0.35 of the tensor peak.  Here each output tile is TW=8 x TH=16 pixels; its halo (64 channels =
// 128 B per pixel, 128B-swizzled, pixel-linear) is loaded once, and tap (dh, dw) is just a
// *shifted window* of the same bytes:
//
//     descriptor start = halo + (dh * HW + dw) * 128 B,   SBO = HW * 128 B (one 8-pixel row group
//     per tile row),   base_offset = 0
//
// The tcgen05 shared-memory descriptor applies the 128B swizzle on absolute address bits, so any
// 128 B-granular shift and any SBO read back exactly what TMA wrote (measured:
// profiles/r01_umma_shifted_window_experiment.txt, tests/native/exp_desc.cu).
//
// A CTA works on "super tiles" of two independent 8x16 tiles (M = 2 x 128) that share every
// weight tile (N = 64), which halves the weight traffic per pixel; when all taps' weights fit
// (<= 9 slots of 8 KB, e.g. 3x3 64->64) they stay resident in shared memory for the whole kernel.
//
// Warp roles as in conv_igemm.cu: warps 0 / 2 / 3 TMA producers, warp 1 MMA issuer, warp 2 also TMEM allocator,
// warps 4..11 epilogue (warpgroup h handles tile h of the pair).
//
// Replaces: slim.conv2d 3x3 / 5x5 stride 1 (research/qtopt/networks.py:443-591) and
// conv2d_fixed_padding 3x3 stride 1 (layers/film_resnet_model.py:89-105), forward and data gradient.
#include <algorithm>
#include <cstdlib>

#include "conv_common.cuh"

namespace t2r {

constexpr int kHaloTW = 8, kHaloTH = 16;
constexpr int kHaloWSlotsStream = 8;
constexpr int kHaloWSlotsMax = 9;

struct HaloParams {
  CUtensorMap tmap_a;  // box 64 x HW x HH x 1
  CUtensorMap tmap_b;  // box 64 x 64
  uint8_t tap_dh[kMaxTaps], tap_dw[kMaxTaps];  // tap offset inside the halo
  int32_t tap_kchunk0[kMaxTaps];
  int n_taps, chunks;
  int HW, HH;
  int org_dh, org_dw;  // halo origin relative to the tile's first output pixel
  int halo_bytes;      // HW*HH*128 rounded up to 1024
  int w_slots;
  int tiles_w, tiles_h;
  int N, Ho, Wo, Cout, n_tiles_n;
  int total_halves, total_super;
  long long os_n, os_h, os_w;
  void* out;
  const void* residual;
  const float* bias;
  int flags;
  double* stats;
};

template <bool kResident>
__global__ void __launch_bounds__(384, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // layout: halo[2 stages][2 tiles] | weights[w_slots][8 KB] | store staging 16 KB | barriers
  const uint32_t w_base = smem_base + 4u * p.halo_bytes;
  const uint32_t store_stage_base = w_base + uint32_t(p.w_slots) * 8192u;
  const uint32_t bar_base = store_stage_base + 8u * 2048u;
  auto halo_addr = [&](int s, int h) { return smem_base + uint32_t(s * 2 + h) * p.halo_bytes; };
  auto hfull_bar = [&](int s) { return bar_base + 8u * s; };
  auto hempty_bar = [&](int s) { return bar_base + 16u + 8u * s; };
  auto tfull_bar = [&](int s) { return bar_base + 32u + 8u * s; };
  auto tempty_bar = [&](int s) { return bar_base + 48u + 8u * s; };
  auto wfull_bar = [&](int s) { return bar_base + 64u + 8u * s; };
  auto wempty_bar = [&](int s) { return bar_base + 64u + 8u * kHaloWSlotsMax + 8u * s; };
  const uint32_t tmem_ptr_addr = bar_base + 64u + 16u * kHaloWSlotsMax;
  float* stat_acc = reinterpret_cast<float*>(smem_raw + (bar_base + 512u - smem_u32(smem_raw)));  // [2][Cout <= 128]
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
  constexpr int kTmemCols = 256;  // 2 accumulator stages x 2 tiles x 64 columns

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a);
    tma_prefetch_desc(&p.tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(hfull_bar(s), 2);  // one arrival (+tx bytes) per halo producer
      mbar_init(hempty_bar(s), 1);
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);
    }
    for (int s = 0; s < p.w_slots; ++s) {
      mbar_init(wfull_bar(s), 1);
      mbar_init(wempty_bar(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_addr, kTmemCols);
    tmem_relinquish();
  }
  if (p.stats != nullptr)
    for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) stat_acc[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const uint32_t halo_tx = uint32_t(p.HW) * p.HH * 128u;

  if (warp == 0 || warp == 2 || warp == 3) {
    // ===================== TMA producers =====================
    // A TMA load costs its whole latency per issuing warp (profiles/r01_ncu_summary.md, 3.3), so the
    // loads of a super tile - two halos and, when streaming, one 8 KB weight tile per tap - are dealt
    // round-robin to three warps.  (With 25 taps a single weight warp was the bottleneck: 25 x ~1000
    // cycles per super tile against 13.8 k cycles of MMA.)  Resident mode: warps 0 / 2 load the halos,
    // warp 3 loads the weights once.
    if (lane == 0) {
      const int pid = warp == 0 ? 0 : warp - 1;
      int hs = 0, ws = 0, turn = 0;
      uint32_t hph = 0, wph = 0;
      bool first = true;
      for (int s = blockIdx.x; s < p.total_super; s += gridDim.x) {
        const int nt = s % p.n_tiles_n;
        const int pair = s / p.n_tiles_n;
        for (int c = 0; c < p.chunks; ++c) {
          for (int h = 0; h < 2; ++h) {
            const bool mine = kResident ? (pid == h) : (turn == pid);
            if (mine) {
              mbar_wait(hempty_bar(hs), hph ^ 1u);
              mbar_expect_tx(hfull_bar(hs), halo_tx);
              const int m = pair * 2 + h;
              int img = p.N, oh0 = 0, ow0 = 0;  // image index N is out of bounds: zero fill
              if (m < p.total_halves) {
                img = m / tiles_per_img;
                const int rem = m - img * tiles_per_img;
                oh0 = (rem / p.tiles_w) * kHaloTH;
                ow0 = (rem % p.tiles_w) * kHaloTW;
              }
              tma_load_4d(halo_addr(hs, h), &p.tmap_a, hfull_bar(hs), c * 64, ow0 + p.org_dw, oh0 + p.org_dh, img);
            }
            if (!kResident && ++turn == 3) turn = 0;
          }
          if (++hs == 2) {
            hs = 0;
            hph ^= 1u;
          }
          if (kResident) {
            if (pid == 2 && first)
              for (int t = 0; t < p.n_taps; ++t) {
                const int slot = c * p.n_taps + t;
                mbar_expect_tx(wfull_bar(slot), 8192u);
                tma_load_2d(w_base + uint32_t(slot) * 8192u, &p.tmap_b, wfull_bar(slot), (p.tap_kchunk0[t] + c) * 64,
                            nt * 64);
              }
          } else {
            for (int t = 0; t < p.n_taps; ++t) {
              if (turn == pid) {
                mbar_wait(wempty_bar(ws), wph ^ 1u);
                mbar_expect_tx(wfull_bar(ws), 8192u);
                tma_load_2d(w_base + uint32_t(ws) * 8192u, &p.tmap_b, wfull_bar(ws), (p.tap_kchunk0[t] + c) * 64,
                            nt * 64);
              }
              if (++turn == 3) turn = 0;
              if (++ws == p.w_slots) {
                ws = 0;
                wph ^= 1u;
              }
            }
          }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this loop converged and an elected lane issues (umma_bf16_elect): descriptors
    // are a precomputed 64-bit base plus a 16-byte-unit offset, all warp-uniform.
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      const uint32_t sbo = uint32_t(p.HW) * 128u;
      const uint64_t a_base = make_smem_desc_sw128(smem_base, 16, sbo, 0);   // + halo offset >> 4
      const uint64_t b_base = make_smem_desc_sw128(w_base, 16, 1024, 0);     // + slot * 512
      const uint32_t halo16 = uint32_t(p.halo_bytes) >> 4;
      int hs = 0, ws = 0, as = 0;
      uint32_t hph = 0, wph = 0, aphase = 0;
      for (int s = blockIdx.x; s < p.total_super; s += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 128;
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(hfull_bar(hs), hph);
          tc_fence_after();
          const uint64_t a_stage = a_base + uint64_t(uint32_t(hs * 2) * halo16);
          for (int t = 0; t < p.n_taps; ++t) {
            const int slot = kResident ? (c * p.n_taps + t) : ws;
            mbar_wait(wfull_bar(slot), kResident ? 0u : wph);
            tc_fence_after();
            const uint64_t bt = b_base + uint64_t(uint32_t(slot) * 512u);
            const uint64_t at = a_stage + uint64_t((uint32_t(p.tap_dh[t]) * p.HW + p.tap_dw[t]) * 8u);
            const uint32_t acc = (c > 0 || t > 0) ? 1u : 0u;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma_bf16_elect(d_tmem, at + uint64_t(kk * 2), bt + uint64_t(kk * 2), idesc, acc | uint32_t(kk > 0));
              umma_bf16_elect(d_tmem + 64, at + uint64_t(halo16 + kk * 2), bt + uint64_t(kk * 2), idesc,
                              acc | uint32_t(kk > 0));
            }
            if (!kResident) {
              umma_commit_elect(wempty_bar(slot));
              if (++ws == p.w_slots) {
                ws = 0;
                wph ^= 1u;
              }
            }
          }
          umma_commit_elect(hempty_bar(hs));
          if (++hs == 2) {
            hs = 0;
            hph ^= 1u;
          }
        }
        umma_commit_elect(tfull_bar(as));
        if (++as == 2) {
          as = 0;
          aphase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: warpgroup h owns tile h of the pair =====================
    const int ew = warp - 4;
    const int quad = ew & 3;
    const int half = ew >> 2;
    const int row = quad * 32 + lane;
    const int th = row >> 3, tw = row & 7;
    const bool out_f32 = (p.flags & T2R_EPI_OUT_F32) != 0;
    const bool has_res = (p.flags & T2R_EPI_RESIDUAL) != 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int s = blockIdx.x; s < p.total_super; s += gridDim.x) {
      const int nt = s % p.n_tiles_n;
      const int m = (s / p.n_tiles_n) * 2 + half;
      const bool live = m < p.total_halves;
      const int img = live ? m / tiles_per_img : 0;
      const int rem = live ? m - img * tiles_per_img : 0;
      const int oh0 = (rem / p.tiles_w) * kHaloTH, ow0 = (rem % p.tiles_w) * kHaloTW;
      const int oh = oh0 + th, ow = ow0 + tw;
      const bool valid = live && (oh < p.Ho) && (ow < p.Wo);
      const long long pix_off = img * p.os_n + oh * p.os_h + ow * p.os_w;
      const int ch0 = nt * 64;
      long long roff[4];
      bool rvalid[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = quad * 32 + 8 * i + (lane >> 2);
        const int roh = oh0 + (r >> 3), row_w = ow0 + (r & 7);
        rvalid[i] = live && (roh < p.Ho) && (row_w < p.Wo);
        roff[i] = img * p.os_n + roh * p.os_h + row_w * p.os_w;
      }
      uint4 rnext[4];
      if (has_res && valid) {
        const uint4* r = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.residual) + pix_off + ch0);
#pragma unroll
        for (int j = 0; j < 4; ++j) rnext[j] = r[j];
      }
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int ch = ch0 + c * 32;
        uint4 rcur[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rcur[j] = rnext[j];
        if (c == 0 && has_res && valid) {
          const uint4* r = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.residual) + pix_off + ch + 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) rnext[j] = r[j];
        }
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + as * 128 + half * 64 + c * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.flags & T2R_EPI_BIAS) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(p.bias + ch + j);
            f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
          }
        }
        if (has_res && valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 q = rcur[j];
            f[8 * j + 0] += bf16_lo(q.x); f[8 * j + 1] += bf16_hi(q.x);
            f[8 * j + 2] += bf16_lo(q.y); f[8 * j + 3] += bf16_hi(q.y);
            f[8 * j + 4] += bf16_lo(q.z); f[8 * j + 5] += bf16_hi(q.z);
            f[8 * j + 6] += bf16_lo(q.w); f[8 * j + 7] += bf16_hi(q.w);
          }
        }
        if (p.flags & T2R_EPI_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (out_f32) {
          if (valid) {
            float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) + pix_off + ch);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
        } else {
          // swizzled per-warp staging so that each store instruction writes 8 pixels x 64 B
          const uint32_t wb = store_stage_base + uint32_t(ew) * 2048u;
          const uint32_t sw = (uint32_t(lane) >> 1) & 3u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t q0 = pack_bf16(f[8 * j + 0], f[8 * j + 1]), q1 = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
            uint32_t q2 = pack_bf16(f[8 * j + 4], f[8 * j + 5]), q3 = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
            if (!valid) q0 = q1 = q2 = q3 = 0u;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wb + uint32_t(lane) * 64u + ((uint32_t(j) ^ sw) << 4)),
                         "r"(q0), "r"(q1), "r"(q2), "r"(q3)
                         : "memory");
          }
          __syncwarp();
          if (p.stats != nullptr) {  // fused bn_stats, see conv_igemm.cu
            float s1 = 0.f, s2 = 0.f;
            const uint32_t cj = uint32_t(lane) >> 3, cb = (uint32_t(lane) & 7u) * 2u;
#pragma unroll
            for (uint32_t r = 0; r < 32; ++r) {
              uint16_t h;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(wb + r * 64u + ((cj ^ ((r >> 1) & 3u)) << 4) + cb) : "memory");
              const float v = __uint_as_float(uint32_t(h) << 16);
              s1 += v;
              s2 = fmaf(v, v, s2);
            }
            atomicAdd(stat_acc + ch + lane, s1);
            atomicAdd(stat_acc + p.Cout + ch + lane, s2);
          }
          const uint32_t jj = uint32_t(lane) & 3u;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t r = 8u * i + (uint32_t(lane) >> 2);
            uint4 q;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                         : "r"(wb + r * 64u + ((jj ^ ((r >> 1) & 3u)) << 4))
                         : "memory");
            if (rvalid[i]) *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + roff[i] + ch + jj * 8) = q;
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (++as == 2) {
        as = 0;
        aphase ^= 1u;
      }
    }
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int i = threadIdx.x - 128; i < 2 * p.Cout; i += 256) {
        const float v = stat_acc[i];
        if (v != 0.f) atomicAdd(p.stats + i, double(v));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

bool conv_halo_eligible(int stride, int n_taps, int k_channels, int n_channels) {
  static const bool disabled = std::getenv("T2R_DISABLE_HALO") != nullptr;
  return !disabled && stride == 1 && n_taps >= 4 && n_channels == 64 && k_channels % 64 == 0 && k_channels <= 128;
}

int conv_halo_launch(const HaloRequest& r, cudaStream_t stream) {
  HaloParams p;
  memset(&p, 0, sizeof(p));
  int dh_min = 127, dh_max = -127, dw_min = 127, dw_max = -127;
  for (int t = 0; t < r.n_taps; ++t) {
    dh_min = std::min<int>(dh_min, r.taps[t].dh); dh_max = std::max<int>(dh_max, r.taps[t].dh);
    dw_min = std::min<int>(dw_min, r.taps[t].dw); dw_max = std::max<int>(dw_max, r.taps[t].dw);
  }
  p.HH = kHaloTH + dh_max - dh_min;
  p.HW = kHaloTW + dw_max - dw_min;
  T2R_CHECK_ARG(p.HH <= 256 && p.HW <= 256 && p.HW * 8 < (1 << 14), "halo too large");
  p.org_dh = dh_min;
  p.org_dw = dw_min;
  for (int t = 0; t < r.n_taps; ++t) {
    p.tap_dh[t] = uint8_t(r.taps[t].dh - dh_min);
    p.tap_dw[t] = uint8_t(r.taps[t].dw - dw_min);
    p.tap_kchunk0[t] = r.taps[t].kchunk0;
  }
  p.n_taps = r.n_taps;
  p.chunks = r.C / 64;
  p.halo_bytes = (p.HW * p.HH * 128 + 1023) & ~1023;
  p.n_tiles_n = r.Cout / 64;
  const bool resident = p.n_tiles_n == 1 && p.n_taps * p.chunks <= kHaloWSlotsMax;
  p.w_slots = resident ? p.n_taps * p.chunks : kHaloWSlotsStream;
  {
    uint64_t dims[4] = {uint64_t(r.C), uint64_t(r.W), uint64_t(r.H), uint64_t(r.N)};
    uint64_t strides[3] = {uint64_t(r.C) * 2, uint64_t(r.W) * r.C * 2, uint64_t(r.H) * r.W * r.C * 2};
    uint32_t box[4] = {64, uint32_t(p.HW), uint32_t(p.HH), 1};
    if (encode_tmap_bf16(&p.tmap_a, r.x, 4, dims, strides, box) != 0) return T2R_ERR_CUDA;
  }
  {
    uint64_t dims[2] = {r.Ktot, uint64_t(r.Cout)};
    uint64_t strides[1] = {r.Ktot * 2};
    uint32_t box[2] = {64, 64};
    if (encode_tmap_bf16(&p.tmap_b, r.w, 2, dims, strides, box) != 0) return T2R_ERR_CUDA;
  }
  p.tiles_w = int(ceil_div(r.Wo, kHaloTW));
  p.tiles_h = int(ceil_div(r.Ho, kHaloTH));
  p.N = r.N; p.Ho = r.Ho; p.Wo = r.Wo; p.Cout = r.Cout;
  p.total_halves = r.N * p.tiles_w * p.tiles_h;
  p.total_super = int(ceil_div(p.total_halves, 2)) * p.n_tiles_n;
  p.os_n = r.os_n; p.os_h = r.os_h; p.os_w = r.os_w;
  p.out = r.out; p.residual = r.residual; p.bias = r.bias; p.flags = r.flags; p.stats = r.stats;
  if (p.total_super <= 0) return T2R_OK;
  const int smem = 4 * p.halo_bytes + p.w_slots * 8192 + 8 * 2048 + 512 /*barriers*/ + 1024 /*stats*/ + 1024 /*align*/;
  T2R_CHECK_ARG(smem <= 227 * 1024, "halo conv needs %d B of shared memory", smem);
  const int grid = std::min(p.total_super, num_sms());
  static bool configured = false;
  if (!configured) {
    T2R_CUDA_OK(cudaFuncSetAttribute(conv_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    T2R_CUDA_OK(cudaFuncSetAttribute(conv_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  if (resident) {
    conv_halo_kernel<true><<<grid, 384, smem, stream>>>(p);
  } else {
    conv_halo_kernel<false><<<grid, 384, smem, stream>>>(p);
  }
  T2R_LAUNCH_OK();
  return T2R_OK;
}

}  // namespace t2r
