// exp_desc_mn.cu — hardware experiment: MN-major SWIZZLE_128B UMMA descriptors over a TMA-written
// "halo" of pixel rows (128 B = 64 channels per pixel).  The wgrad A operand is MN-major: M = channels,
// K = pixels.  Question: with base_offset = 0, can the descriptor (a) start at any 128 B-aligned pixel,
// (b) step between 8-pixel K groups with an SBO that is the halo row pitch, and (c) place the second
// 64-channel M block at an arbitrary LBO (another tap's window, e.g. +128 B)?
//
//   A[k][m]: k = 16 pixels = 2 groups of 8 consecutive halo pixels, group g at start + g*SBO;
//            m = 128 = 2 blocks of 64 channels, block j at + j*LBO.
//   B[k][n] = delta(k, n) (MN-major, 16 x 64) so that D[m][n] = A[n][m] for n < 16.
#include <vector>
#include "../../tensor2robot_b200/csrc/common.cuh"
using namespace t2r;

struct P { CUtensorMap ta, tb; int start_px, sbo, lbo; float* out; };

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sa = base, sb = base + 32768, bar = sb + 8192, bar2 = bar + 8, tptr = bar + 16;
  volatile uint32_t* tptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(tptr, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tptr_gen;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 256 * 128 + 64 * 128);
    tma_load_2d(sa, &p.ta, bar, 0, 0);
    tma_load_2d(sb, &p.tb, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
    const uint64_t ad = make_smem_desc_sw128(sa + p.start_px * 128, p.lbo, p.sbo, 0);
    const uint64_t bd = make_smem_desc_sw128(sb, 8192, 1024, 0);
    umma_bf16(tmem, ad, bd, idesc, 0u);
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

int main() {
  const int R = 256;
  // halo value: pixel r, channel c -> r + c/64 (c/64 is exactly representable only coarsely; use two runs)
  std::vector<__nv_bfloat16> hp(R * 64), hc(R * 64), hb(64 * 64);
  for (int r = 0; r < R; ++r) for (int c = 0; c < 64; ++c) { hp[r * 64 + c] = __float2bfloat16(float(r)); hc[r * 64 + c] = __float2bfloat16(float(c)); }
  for (int kk = 0; kk < 64; ++kk) for (int n = 0; n < 64; ++n) hb[kk * 64 + n] = __float2bfloat16((kk == n && kk < 16) ? 1.f : 0.f);
  __nv_bfloat16 *dp, *dc, *db; float* dout;
  cudaMalloc(&dp, hp.size() * 2); cudaMalloc(&dc, hc.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(dp, hp.data(), hp.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dc, hc.data(), hc.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  printf("start_px sbo lbo | pixels_ok channels_ok\n");
  const int starts[] = {0, 1, 3, 8, 13};
  const int sbos[] = {1024, 1280, 1536};
  const int lbos[] = {8192, 128, 256, 1280 + 128, 1024};
  for (int sbo : sbos) for (int lbo : lbos) for (int st : starts) {
    bool ok[2];
    for (int which = 0; which < 2; ++which) {
      P p; uint64_t dims[2] = {64, uint64_t(R)}, strides[1] = {128}; uint32_t box[2] = {64, 256};
      if (encode_tmap_bf16(&p.ta, which ? dc : dp, 2, dims, strides, box)) return 1;
      uint64_t dimsb[2] = {64, 64}; uint32_t boxb[2] = {64, 64};
      if (encode_tmap_bf16(&p.tb, db, 2, dimsb, strides, boxb)) return 1;
      p.start_px = st; p.sbo = sbo; p.lbo = lbo; p.out = dout;
      cudaMemset(dout, 0xff, 128 * 64 * 4);
      k<<<1, 128, 48 * 1024>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 2; }
      std::vector<float> h(128 * 64); cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost);
      bool good = true;
      for (int m = 0; m < 128 && good; ++m) for (int n = 0; n < 16; ++n) {
        const int blk = m / 64, ch = m % 64;
        const int px = st + blk * (lbo / 128) + (n / 8) * (sbo / 128) + (n % 8);
        const float want = which ? float(ch) : float(px);
        if (h[m * 64 + n] != want) { good = false; break; }
      }
      ok[which] = good;
    }
    printf("%8d %4d %5d | %9s %11s\n", st, sbo, lbo, ok[0] ? "yes" : "NO", ok[1] ? "yes" : "NO");
  }
  return 0;
}
