// exp_mma_rate.cu — hardware experiment: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128,
// K=16, both operands in shared memory) as a function of N, of the A-descriptor geometry (aligned
// tile vs shifted halo window with SBO != 1024) and of how many distinct accumulators are in flight.
// Every SM runs one CTA; one thread issues `iters` MMAs back to back and waits for the final commit.
#include <vector>

#include "../../tensor2robot_b200/csrc/common.cuh"

using namespace t2r;

struct RateParams {
  int n;          // 64 / 128 / 256
  int iters;
  int a_off;      // byte offset of the A window start (multiple of 128)
  int sbo;        // A stride between 8-row groups
  int n_acc;      // accumulators cycled through (1, 2 or 4)
  int a_tiles;    // distinct A tiles cycled through (bank / reuse effects)
  long long* cycles;  // [gridDim.x]
};

template <int N, int AMN, int BMN>
__global__ void __launch_bounds__(128, 1) rate_kernel(const RateParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sa = base;                 // 4 x 40 KB of "halo"
  const uint32_t sb = base + 160 * 1024;    // 256 rows x 128 B
  const uint32_t bar = sb + 32768;
  const uint32_t tptr = bar + 16;
  volatile uint32_t* tptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5;
  // initialise operands with finite values
  for (uint32_t i = threadIdx.x; i < (192 * 1024) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;  // tiny bf16 pairs
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr_gen;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, AMN, BMN);
    const long long t0 = clock64();
    int acc = 0, at = 0;
    for (int i = 0; i < p.iters; ++i) {
      const int kk = i & 3;
      // K-major: K advance = 32 B inside the swizzled row; MN-major: 16 pixels = 2 x 8-row atoms = 2048 B,
      // 64-wide M/N blocks 8 KB apart (the wgrad layout).
      const uint64_t ad = AMN ? make_smem_desc_sw128(sa + at * 40960 + kk * 2048, 8192, 1024, 0)
                              : make_smem_desc_sw128(sa + at * 40960 + p.a_off + kk * 32, 16, p.sbo, 0);
      const uint64_t bd = BMN ? make_smem_desc_sw128(sb + kk * 2048, 8192, 1024, 0)
                              : make_smem_desc_sw128(sb + kk * 32, 16, 1024, 0);
      umma_bf16(tmem + acc * N, ad, bd, idesc, 1u);
      if (++acc == p.n_acc) acc = 0;
      if (kk == 3 && ++at == p.a_tiles) at = 0;
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    p.cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N, int AMN = 0, int BMN = 0>
static double run(int iters, int a_off, int sbo, int n_acc, int a_tiles, int grid) {
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  RateParams p{N, iters, a_off, sbo, n_acc, a_tiles, d};
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(rate_kernel<N, AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<N, AMN, BMN><<<grid, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); exit(2); }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  long long mx = 0;
  for (auto v : h) mx = std::max(mx, v);
  return double(mx) / iters;
}

int main() {
  const int iters = 4096;
  printf("grid N a_off sbo n_acc a_tiles | cycles_per_mma  TFLOP/s_at_1.9GHz_148SM\n");
  for (int grid : {148}) {
    for (int n : {64, 128, 256}) {
      struct { int a_off, sbo, n_acc, a_tiles; } cfgs[] = {
          {0, 1024, 1, 1}, {0, 1024, 2, 1}, {0, 1024, 2, 4}, {1408, 1280, 2, 1}, {1408, 1536, 2, 4}, {0, 1024, 4, 4}};
      for (auto c : cfgs) {
        if (c.n_acc * n > 512) continue;
        double cyc = n == 64    ? run<64>(iters, c.a_off, c.sbo, c.n_acc, c.a_tiles, grid)
                     : n == 128 ? run<128>(iters, c.a_off, c.sbo, c.n_acc, c.a_tiles, grid)
                                : run<256>(iters, c.a_off, c.sbo, c.n_acc, c.a_tiles, grid);
        printf("%4d %3d %5d %4d %5d %7d | %8.1f  %8.1f\n", grid, n, c.a_off, c.sbo, c.n_acc, c.a_tiles, cyc,
               2.0 * 128 * n * 16 / cyc * 1.9e9 * 148 / 1e12);
      }
    }
  }
  printf("\nMN-major operands (grid 148, 2 accumulators): N  A_mn B_mn | cycles_per_mma\n");
  printf("%3d %d %d | %.1f\n", 64, 1, 1, run<64, 1, 1>(iters, 0, 1024, 2, 1, 148));
  printf("%3d %d %d | %.1f\n", 128, 1, 1, run<128, 1, 1>(iters, 0, 1024, 2, 1, 148));
  printf("%3d %d %d | %.1f\n", 256, 1, 1, run<256, 1, 1>(iters, 0, 1024, 2, 1, 148));
  printf("%3d %d %d | %.1f\n", 256, 1, 0, run<256, 1, 0>(iters, 0, 1024, 2, 1, 148));
  printf("%3d %d %d | %.1f\n", 256, 0, 1, run<256, 0, 1>(iters, 0, 1024, 2, 1, 148));
  printf("%3d %d %d | %.1f\n", 128, 1, 0, run<128, 1, 0>(iters, 0, 1024, 2, 1, 148));
  printf("%3d %d %d | %.1f\n", 128, 0, 1, run<128, 0, 1>(iters, 0, 1024, 2, 1, 148));
  return 0;
}
