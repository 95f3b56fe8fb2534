// exp_tma_rate.cu — hardware experiment: sustained TMA tile-load rate per SM (bytes / SM-cycle) for
// the activation boxes the conv kernels use, as a function of box height, ring depth, re-read factor
// (1 = every tile comes from DRAM, 9 = a 3x3 filter's tap re-reads, mostly L2 hits) and the number
// of producer threads issuing the loads.  One CTA per SM; a consumer thread only waits + releases.
#include <vector>

#include "../../tensor2robot_b200/csrc/common.cuh"

using namespace t2r;

struct TmaParams {
  CUtensorMap map;     // [N][H][W][64] bf16, box 64 x 8 x TH x 1
  int stages, stage_bytes, tiles_w, tiles_h, TH, n_img, rereads, tiles_per_cta, producers;
  long long* cycles;
};

__global__ void __launch_bounds__(256, 1) tma_kernel(const __grid_constant__ TmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int me = warp & 3;  // stream index: producer warp `me`, consumer warp 4 + `me`
  const uint32_t bar0 = base + p.producers * p.stages * p.stage_bytes;
  const uint32_t bar = bar0 + me * 512;
  const uint32_t mybase = base + me * p.stages * p.stage_bytes;
  auto full = [&](int s) { return bar + 8u * s; };
  auto empty = [&](int s) { return bar + 8u * (32 + s); };
  if (warp < p.producers && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int tpi = p.tiles_w * p.tiles_h;
  const long long t0 = clock64();
  if (warp < p.producers && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < p.tiles_per_cta; ++i) {
      const int tile = ((blockIdx.x * p.producers + me) * p.tiles_per_cta + i) % (p.n_img * tpi);
      const int img = tile / tpi, rem = tile % tpi;
      const int oh0 = (rem / p.tiles_w) * p.TH, ow0 = (rem % p.tiles_w) * 8;
      for (int r = 0; r < p.rereads; ++r) {
        mbar_wait(empty(s), ph ^ 1u);
        mbar_expect_tx(full(s), p.stage_bytes);
        tma_load_4d(mybase + s * p.stage_bytes, &p.map, full(s), 0, ow0 + (r % 3) - 1, oh0 + (r / 3) - 1, img);
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp >= 4 && warp < 4 + p.producers && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < p.tiles_per_cta * p.rereads; ++i) {
      mbar_wait(full(s), ph);
      mbar_arrive(empty(s));
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
    if (me == 0) p.cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const int N = 512, H = 118, W = 118, C = 64;
  const size_t bytes = size_t(N) * H * W * C * 2;
  void* x;
  cudaMalloc(&x, bytes);
  cudaMemset(x, 0x11, bytes);
  long long* dcyc;
  cudaMalloc(&dcyc, 148 * sizeof(long long));
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("TH stage_KB stages rereads | B/clk/SM   TB/s(all SMs, measured wall)\n");
  for (int TH : {8, 16})
    for (int producers : {1, 2, 4})
    for (int stages : {4})
      for (int rereads : {1, 9}) {
        TmaParams p;
        p.producers = producers;
        p.TH = TH; p.stage_bytes = 8 * TH * 128; p.stages = stages;
        if (p.producers * p.stages * p.stage_bytes > 200 * 1024) continue;
        uint64_t dims[4] = {uint64_t(C), uint64_t(W), uint64_t(H), uint64_t(N)};
        uint64_t strides[3] = {uint64_t(C) * 2, uint64_t(W) * C * 2, uint64_t(H) * W * C * 2};
        uint32_t box[4] = {64, 8, uint32_t(TH), 1};
        if (encode_tmap_bf16(&p.map, x, 4, dims, strides, box)) { printf("tmap fail\n"); return 1; }
        p.tiles_w = (W + 7) / 8; p.tiles_h = (H + TH - 1) / TH; p.n_img = N; p.rereads = rereads;
        p.tiles_per_cta = (rereads == 1 ? 2048 : 512) * 8 / TH / producers;
        p.cycles = dcyc;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        tma_kernel<<<148, 256, p.producers * p.stages * p.stage_bytes + 4096>>>(p);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 2; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> h(148);
        cudaMemcpy(h.data(), dcyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0; for (auto v : h) mx = std::max(mx, v);
        const double per_cta = double(p.tiles_per_cta) * rereads * p.stage_bytes * producers;
        printf("%2d %7d %6d*%d %7d | %8.1f   %6.2f\n", TH, p.stage_bytes / 1024, stages, producers, rereads, per_cta / mx,
               per_cta * 148 / (ms * 1e-3) / 1e12);
      }
  return 0;
}
