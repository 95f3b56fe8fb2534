// conv_wgrad_halo.cu — weight gradient of stride-1 KxK convolutions with 64 input channels, with the
// im2col done inside shared memory (the wgrad counterpart of conv_halo.cu).
//
//   dW[co, tap, ci] = sum over pixels  dY[pix, co] * X[pix + tap, ci]
//
// conv_wgrad.cu re-loads a 64-pixel X tile for every pair of taps.  Here a pixel tile is 8 x 8 output
// pixels; its (8+KH-1) x (8+KW-1) pixel halo of X (128 B = 64 channels per pixel, 128B-swizzled,
// pixel-linear) is loaded ONCE and every tap pair reads it through a shifted MN-major descriptor:
//
//   A (M = 2 taps x 64 channels, K = 16 pixels = 2 tile rows):
//       start = halo + ((dh1 + 2*kk) * HW + dw1) * 128 B     tap 1, tile rows 2kk, 2kk+1
//       SBO   = HW * 128 B                                   next tile row (8-pixel K group)
//       LBO   = ((dh2 - dh1) * HW + (dw2 - dw1)) * 128 B     tap 2's window, as the second M block
//   B (dY tile, 64 pixels x BLOCK_N channels, MN-major): as in conv_wgrad.cu.
//
// MN-major SWIZZLE_128B descriptors with base_offset = 0 accept any 128 B-aligned start, SBO and LBO
// (measured: tests/native/exp_desc_mn.cu).  L2 -> SMEM traffic per pixel tile drops from
// (taps/2) * 16 KB + 8 KB to one 13-19 KB halo + 8 KB, which makes the kernel MMA-bound (N = 64 rate).
//
// Replaces the filter gradients of slim.conv2d 3x3 / 5x5 (research/qtopt/networks.py:443-591) and of
// conv2d_fixed_padding 3x3 stride 1 with 64 channels (layers/film_resnet_model.py:89-105).
#include <algorithm>
#include <cstdlib>

#include "conv_common.cuh"

namespace t2r {

struct WgradHaloParams {
  CUtensorMap tmap_x;   // box 64 x HW x HH x 1 over X
  CUtensorMap tmap_dy;  // box 64 x 8 x 8 x 1 over dY
  uint8_t tap_dh[kMaxTaps], tap_dw[kMaxTaps];  // tap offsets inside the halo
  int n_taps, n_groups, groups_per_cta, n_gsets, n_chunks_n, ksplits;
  int HW, HH, org_dh, org_dw, halo_bytes;
  int tiles_w, tiles_h, total_ptiles;
  int Ktot, Cout;
  float* dw;
};

constexpr int kWhTile = 8;  // 8 x 8 output pixels per tile
constexpr int kWhStages = 6;

template <int BLOCK_N>
__global__ void __launch_bounds__(256, 1) conv_wgrad_halo_kernel(const __grid_constant__ WgradHaloParams p) {
  constexpr int kNB = BLOCK_N / 64;
  constexpr int kDyBytes = kNB * 8192;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = uint32_t(p.halo_bytes) + kDyBytes;
  const uint32_t bar_base = smem_base + kWhStages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWhStages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * kWhStages);
  const uint32_t tmem_ptr_addr = tfull_bar + 8u;
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_x);
    tma_prefetch_desc(&p.tmap_dy);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWhStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_addr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // work item: (group set, output-channel chunk, pixel-tile range)
  int item = blockIdx.x;
  const int ks = item % p.ksplits; item /= p.ksplits;
  const int nc = item % p.n_chunks_n; item /= p.n_chunks_n;
  const int g0 = item * p.groups_per_cta;
  const int g1 = min(g0 + p.groups_per_cta, p.n_groups);
  const int pt0 = int((long long)p.total_ptiles * ks / p.ksplits);
  const int pt1 = int((long long)p.total_ptiles * (ks + 1) / p.ksplits);
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int n0 = nc * BLOCK_N;
  const uint32_t halo_tx = uint32_t(p.HW) * p.HH * 128u;

  if (warp == 0 || warp == 2 || warp == 3) {
    // three producer warps take the pixel tiles round-robin (see conv_igemm.cu)
    if (lane == 0) {
      const int pid = warp == 0 ? 0 : warp - 1;
      int stage = 0, turn = 0;
      uint32_t phase = 0;
      for (int pt = pt0; pt < pt1; ++pt) {
        if (turn == pid) {
          const int img = pt / tiles_per_img;
          const int rem = pt - img * tiles_per_img;
          const int oh0 = (rem / p.tiles_w) * kWhTile;
          const int ow0 = (rem % p.tiles_w) * kWhTile;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sx = smem_base + stage * stage_bytes;
          mbar_expect_tx(full_bar(stage), halo_tx + kDyBytes);
          tma_load_4d(sx, &p.tmap_x, full_bar(stage), 0, ow0 + p.org_dw, oh0 + p.org_dh, img);
#pragma unroll
          for (int j = 0; j < kNB; ++j)
            tma_load_4d(sx + p.halo_bytes + j * 8192, &p.tmap_dy, full_bar(stage), n0 + j * 64, ow0, oh0, img);
        }
        if (++turn == 3) turn = 0;
        if (++stage == kWhStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // whole warp converged, elected lane issues (see conv_igemm.cu)
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 1, 1);
      const uint32_t sbo = uint32_t(p.HW) * 128u;
      const uint64_t b_base = make_smem_desc_sw128(smem_base + p.halo_bytes, 8192, 1024, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = pt0; pt < pt1; ++pt) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sx = smem_base + stage * stage_bytes;
        const uint64_t bs = b_base + uint64_t(uint32_t(stage) * (stage_bytes >> 4));
        for (int g = g0; g < g1; ++g) {
          const int t1 = 2 * g, t2 = min(2 * g + 1, p.n_taps - 1);
          const uint32_t a1 = (uint32_t(p.tap_dh[t1]) * p.HW + p.tap_dw[t1]) * 128u;
          const uint32_t a2 = (uint32_t(p.tap_dh[t2]) * p.HW + p.tap_dw[t2]) * 128u;
          const uint32_t lbo = a2 > a1 ? a2 - a1 : 128u;   // odd tap count: the spare M block is never stored
          const uint64_t ad = make_smem_desc_sw128(sx + a1, lbo, sbo, 0);
          const uint32_t d_tmem = tmem_base + (g - g0) * BLOCK_N;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // 64 pixels = 4 x (K = 16 = two 8-pixel tile rows)
            umma_bf16_elect(d_tmem, ad + uint64_t(uint32_t(2 * kk) * (sbo >> 4)), bs + uint64_t(kk * 128), idesc,
                            (pt > pt0 || kk > 0) ? 1u : 0u);
        }
        umma_commit_elect(empty_bar(stage));
        if (++stage == kWhStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit_elect(tfull_bar);
    }
  } else if (warp >= 4) {
    const int quad = warp - 4;
    const int row = quad * 32 + lane;
    if (pt1 > pt0) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      for (int g = g0; g < g1; ++g) {
        const int k = g * 128 + row;  // column of the OHWI weight row: tap * 64 + ci
        const bool kvalid = k < p.Ktot;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + (g - g0) * BLOCK_N + c0, v);
          tmem_ld_wait();
          if (kvalid) {
            float* dst = p.dw + (long long)(n0 + c0) * p.Ktot + k;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c0 + j < p.Cout) atomicAdd(dst + (long long)j * p.Ktot, __uint_as_float(v[j]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

bool conv_wgrad_halo_eligible(int stride, int n_taps, int Cin, int Cout) {
  static const bool disabled = std::getenv("T2R_DISABLE_WGRAD_HALO") != nullptr;
  return !disabled && stride == 1 && n_taps >= 4 && Cin == 64 && Cout % 64 == 0 && Cout <= 128;
}

// taps: dh/dw of every filter tap relative to the output pixel (ConvTap), in OHWI tap order.
int conv_wgrad_halo_launch(const void* x, const void* dy, float* dw, int N, int H, int W, int Ho, int Wo, int Cout,
                           const ConvTap* taps, int n_taps, cudaStream_t stream) {
  WgradHaloParams p;
  memset(&p, 0, sizeof(p));
  int dh_min = 127, dh_max = -127, dw_min = 127, dw_max = -127;
  for (int t = 0; t < n_taps; ++t) {
    dh_min = std::min<int>(dh_min, taps[t].dh); dh_max = std::max<int>(dh_max, taps[t].dh);
    dw_min = std::min<int>(dw_min, taps[t].dw); dw_max = std::max<int>(dw_max, taps[t].dw);
  }
  p.HH = kWhTile + dh_max - dh_min;
  p.HW = kWhTile + dw_max - dw_min;
  p.org_dh = dh_min; p.org_dw = dw_min;
  for (int t = 0; t < n_taps; ++t) {
    p.tap_dh[t] = uint8_t(taps[t].dh - dh_min);
    p.tap_dw[t] = uint8_t(taps[t].dw - dw_min);
  }
  p.n_taps = n_taps;
  // one spare pixel row keeps the (unused) second M block of an odd last group inside the stage
  p.halo_bytes = ((p.HW * p.HH + 1) * 128 + 1023) & ~1023;
  const int block_n = Cout % 128 == 0 ? 128 : 64;
  {
    uint64_t dims[4] = {64, uint64_t(W), uint64_t(H), uint64_t(N)};
    uint64_t strides[3] = {uint64_t(64) * 2, uint64_t(W) * 64 * 2, uint64_t(H) * W * 64 * 2};
    uint32_t box[4] = {64, uint32_t(p.HW), uint32_t(p.HH), 1};
    if (encode_tmap_bf16(&p.tmap_x, x, 4, dims, strides, box) != 0) return T2R_ERR_CUDA;
  }
  {
    uint64_t dims[4] = {uint64_t(Cout), uint64_t(Wo), uint64_t(Ho), uint64_t(N)};
    uint64_t strides[3] = {uint64_t(Cout) * 2, uint64_t(Wo) * Cout * 2, uint64_t(Ho) * Wo * Cout * 2};
    uint32_t box[4] = {64, kWhTile, kWhTile, 1};
    if (encode_tmap_bf16(&p.tmap_dy, dy, 4, dims, strides, box) != 0) return T2R_ERR_CUDA;
  }
  p.n_groups = (n_taps + 1) / 2;
  const int max_groups = 512 / block_n;
  p.n_gsets = int(ceil_div(p.n_groups, max_groups));
  p.groups_per_cta = int(ceil_div(p.n_groups, p.n_gsets));   // balanced group sets
  p.n_chunks_n = Cout / block_n;
  p.tiles_w = int(ceil_div(Wo, kWhTile));
  p.tiles_h = int(ceil_div(Ho, kWhTile));
  p.total_ptiles = N * p.tiles_w * p.tiles_h;
  p.Ktot = n_taps * 64;
  p.Cout = Cout;
  p.dw = dw;
  const int base_items = p.n_gsets * p.n_chunks_n;
  int ks = std::max(1, num_sms() / base_items);
  ks = std::max(1, std::min(ks, std::max(1, p.total_ptiles / 4)));
  p.ksplits = ks;
  const int smem = kWhStages * (p.halo_bytes + (block_n / 64) * 8192) + 1024 + 256;
  T2R_CHECK_ARG(smem <= 227 * 1024, "wgrad halo needs %d B of shared memory", smem);
  static bool configured = false;
  if (!configured) {
    T2R_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_halo_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    T2R_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_halo_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int grid = base_items * ks;
  if (block_n == 128) conv_wgrad_halo_kernel<128><<<grid, 256, smem, stream>>>(p);
  else conv_wgrad_halo_kernel<64><<<grid, 256, smem, stream>>>(p);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

}  // namespace t2r
