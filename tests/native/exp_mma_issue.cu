// exp_mma_issue.cu — hardware experiment: is a stream of N = 64 / 128 tcgen05.mma instructions bound by
// the tensor pipe or by the issuing thread?  Three issue styles for the same MMA stream (M = 128, K = 16,
// K-major SW128 operands, 2 accumulators, 4 k-steps per "tile"):
//   mode 0: one elected lane inside `if (lane == 0)`, descriptors rebuilt with make_smem_desc_sw128
//   mode 1: same branch, descriptors = precomputed 64-bit base + (byte offset >> 4)
//   mode 2: the whole warp runs the loop converged; the MMA is predicated with elect.sync, so that every
//           operand is warp-uniform (uniform registers / uniform datapath)
#include <vector>
#include "../../tensor2robot_b200/csrc/common.cuh"
using namespace t2r;

struct P { int iters; long long* cycles; };

__device__ __forceinline__ void umma_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) k(const P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sa = base, sb = base + 64 * 1024, bar = sb + 32768, tptr = bar + 16;
  volatile uint32_t* tptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (100 * 1024) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tptr_gen;
  constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
  if (warp == 0) {
    const long long t0 = clock64();
    if (MODE == 0) {
      if (lane == 0) {
        for (int i = 0; i < p.iters; ++i) {
          const int kk = i & 3, tile = (i >> 2) & 3;
          const uint64_t ad = make_smem_desc_sw128(sa + tile * 16384 + kk * 32, 16, 1024, 0);
          const uint64_t bd = make_smem_desc_sw128(sb + kk * 32, 16, 1024, 0);
          umma_bf16(tmem + (i & 1) * N, ad, bd, idesc, 1u);
        }
        umma_commit(bar);
      }
    } else if (MODE == 1) {
      if (lane == 0) {
        const uint64_t a0 = make_smem_desc_sw128(sa, 16, 1024, 0), b0 = make_smem_desc_sw128(sb, 16, 1024, 0);
        for (int i = 0; i < p.iters; ++i) {
          const int kk = i & 3, tile = (i >> 2) & 3;
          umma_bf16(tmem + (i & 1) * N, a0 + uint64_t((tile * 16384 + kk * 32) >> 4), b0 + uint64_t((kk * 32) >> 4), idesc, 1u);
        }
        umma_commit(bar);
      }
    } else {
      const uint64_t a0 = make_smem_desc_sw128(sa, 16, 1024, 0), b0 = make_smem_desc_sw128(sb, 16, 1024, 0);
      for (int i = 0; i < p.iters; ++i) {
        const int kk = i & 3, tile = (i >> 2) & 3;
        umma_elect(tmem + (i & 1) * N, a0 + uint64_t((tile * 16384 + kk * 32) >> 4), b0 + uint64_t((kk * 32) >> 4), idesc, 1u);
      }
      if (lane == 0) umma_commit(bar);
    }
    if (lane == 0) {
      mbar_wait(bar, 0);
      p.cycles[blockIdx.x] = clock64() - t0;
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int MODE>
static double run(int iters) {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  P p{iters, d};
  cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  k<N, MODE><<<148, 128, 120 * 1024>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); exit(2); }
  std::vector<long long> h(148); cudaMemcpy(h.data(), d, sizeof(long long) * 148, cudaMemcpyDeviceToHost); cudaFree(d);
  long long mx = 0; for (auto v : h) mx = std::max(mx, v);
  return double(mx) / iters;
}

int main() {
  const int iters = 8192;
  printf("N mode | cycles per MMA\n");
  printf(" 64 0 | %.1f\n", run<64, 0>(iters));
  printf(" 64 1 | %.1f\n", run<64, 1>(iters));
  printf(" 64 2 | %.1f\n", run<64, 2>(iters));
  printf("128 0 | %.1f\n", run<128, 0>(iters));
  printf("128 1 | %.1f\n", run<128, 1>(iters));
  printf("128 2 | %.1f\n", run<128, 2>(iters));
  printf("256 0 | %.1f\n", run<256, 0>(iters));
  printf("256 2 | %.1f\n", run<256, 2>(iters));
  return 0;
}
