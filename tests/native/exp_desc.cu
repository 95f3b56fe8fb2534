// exp_desc.cu — hardware experiment: can a UMMA shared-memory descriptor (SWIZZLE_128B, K-major)
// address a SHIFTED window of a larger TMA-written tile?  (prerequisite for im2col reuse in shared
// memory: one halo tile serving all filter taps.)
//
// A "halo" of R rows x 64 bf16 (128 B per row) is written by ONE TMA load (128B swizzle, address
// based).  The MMA then reads M = 128 rows as 16 groups of 8 rows: group g starts at
// start + g * SBO, with start = base + off * 128.  We sweep off (window shift in rows), SBO (row
// pitch of a halo wider than the tile) and the descriptor's base_offset field, and report which
// combinations return the right rows AND the right columns.
#include <vector>

#include "../../tensor2robot_b200/csrc/common.cuh"

using namespace t2r;

struct ExpParams {
  CUtensorMap tmap_a, tmap_b;
  int off_rows, sbo_bytes, base_offset_mode;  // mode 0: 0, 1: (start >> 7) & 7
  float* out;                                 // [128][64]
};

__global__ void __launch_bounds__(128, 1) exp_kernel(const __grid_constant__ ExpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sa = base;               // up to 256 rows * 128 B = 32 KB
  const uint32_t sb = base + 32768;       // 64 rows * 128 B
  const uint32_t bar = base + 32768 + 8192;
  const uint32_t bar2 = bar + 8;
  const uint32_t tptr = bar + 16;
  volatile uint32_t* tptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr_gen;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 256 * 128 + 64 * 128);
    tma_load_2d(sa, &p.tmap_a, bar, 0, 0);
    tma_load_2d(sb, &p.tmap_b, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t start = sa + p.off_rows * 128;
    const uint32_t bo = p.base_offset_mode ? ((start >> 7) & 7u) : 0u;
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t ad = make_smem_desc_sw128(start + kk * 32, 16, p.sbo_bytes, bo);
      const uint64_t bd = make_smem_desc_sw128(sb + kk * 32, 16, 1024, 0);
      umma_bf16(tmem, ad, bd, idesc, kk > 0);
    }
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
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

int main() {
  const int R = 256;
  std::vector<__nv_bfloat16> hrow(R * 64), hcol(R * 64), hb(64 * 64);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < 64; ++c) {
      hrow[r * 64 + c] = __float2bfloat16(float(r));
      hcol[r * 64 + c] = __float2bfloat16(float(c));
    }
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) hb[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *drow, *dcol, *db;
  float* dout;
  cudaMalloc(&drow, hrow.size() * 2); cudaMalloc(&dcol, hcol.size() * 2); cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(drow, hrow.data(), hrow.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dcol, hcol.data(), hcol.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  const int offs[] = {0, 1, 3, 8, 10, 13};
  const int sbos[] = {1024, 1280, 1536, 2048};
  printf("off_rows sbo base_offset_mode | rows_ok cols_ok\n");
  for (int mode = 0; mode < 2; ++mode)
    for (int sbo : sbos)
      for (int off : offs) {
        if (off + 15 * (sbo / 128) + 8 > R) continue;
        bool ok[2];
        for (int which = 0; which < 2; ++which) {
          ExpParams p;
          uint64_t dims[2] = {64, uint64_t(R)};
          uint64_t strides[1] = {128};
          uint32_t box[2] = {64, 256};
          if (encode_tmap_bf16(&p.tmap_a, which ? dcol : drow, 2, dims, strides, box)) { printf("tmap fail\n"); return 1; }
          uint64_t dimsb[2] = {64, 64};
          uint32_t boxb[2] = {64, 64};
          if (encode_tmap_bf16(&p.tmap_b, db, 2, dimsb, strides, boxb)) { printf("tmap fail\n"); return 1; }
          p.off_rows = off; p.sbo_bytes = sbo; p.base_offset_mode = mode; p.out = dout;
          cudaMemset(dout, 0xff, 128 * 64 * 4);
          exp_kernel<<<1, 128, 48 * 1024>>>(p);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 2; }
          std::vector<float> h(128 * 64);
          cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost);
          bool good = true;
          for (int m = 0; m < 128 && good; ++m)
            for (int n = 0; n < 64; ++n) {
              const float want = which ? float(n) : float(off + (m / 8) * (sbo / 128) + (m % 8));
              if (h[m * 64 + n] != want) { good = false; break; }
            }
          ok[which] = good;
        }
        printf("%8d %4d %16d | %7s %7s\n", off, sbo, mode, ok[0] ? "yes" : "NO", ok[1] ? "yes" : "NO");
      }
  return 0;
}
