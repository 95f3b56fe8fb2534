// exp_tma5d.cu — where does a 5-D TMA box with a 64-byte inner extent and SWIZZLE_128B land in shared memory?
#include <vector>
#include "../../tensor2robot_b200/csrc/common.cuh"
using namespace t2r;
struct P { CUtensorMap map; uint4* out; int c2, c3, bytes; };
__global__ void k(const __grid_constant__ P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 16384;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, p.bytes);
    tma_load_5d(base, &p.map, bar, 0, 0, p.c2, p.c3, 0);
    mbar_wait(bar, 0);
  }
  __syncthreads();
  const uint4* s = reinterpret_cast<const uint4*>(smem_raw + (base - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) p.out[i] = s[i];
}
int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int Hp = 40, Wp = 48, N = 1, s = 2, TW = 8, TH = 16;
  std::vector<__nv_bfloat16> h(size_t(N) * Hp * Wp * 4);
  // value = row*64 + col (exact in bf16 only up to 256 -> use row in ch0, col in ch1, 0, 0)
  for (int r = 0; r < Hp; ++r) for (int c = 0; c < Wp; ++c) {
    h[(size_t(r) * Wp + c) * 4 + 0] = __float2bfloat16(float(r));
    h[(size_t(r) * Wp + c) * 4 + 1] = __float2bfloat16(float(c));
    h[(size_t(r) * Wp + c) * 4 + 2] = __float2bfloat16(0.f);
    h[(size_t(r) * Wp + c) * 4 + 3] = __float2bfloat16(0.f);
  }
  __nv_bfloat16* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  uint4* out; cudaMalloc(&out, 16384);
  P p; p.out = out; p.c2 = 1; p.c3 = 2;
  const int Wo = 16, rows = (Hp - 2) / s + 1;
  uint64_t dims[5] = {32, 2, uint64_t(Wo), uint64_t(rows), uint64_t(N)};
  uint64_t strides[4] = {uint64_t(Wp) * 8, uint64_t(s) * 8, uint64_t(s) * Wp * 8, uint64_t(Hp) * Wp * 8};
  uint32_t box[5] = {32, 2, TW, TH, 1};
  if (mode == 1) { dims[1] = 1; box[1] = 1; }                       // degenerate pair dimension
  if (mode == 2) { dims[0] = 64; box[0] = 64; dims[1] = 1; box[1] = 1; }  // 128-byte windows, 5-D
  if (mode == 3) { strides[0] = 16; dims[1] = 2; }                  // monotonic strides (pair = +16 B)
  printf("mode %d\n", mode);
  p.bytes = int(box[0] * box[1] * box[2] * box[3] * 2);
  if (encode_tmap_bf16(&p.map, d, 5, dims, strides, box)) { printf("encode failed\n"); return 1; }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480);
  k<<<1, 128, 20480>>>(p);
  { cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 2; } }
  std::vector<__nv_bfloat16> o(8192); cudaMemcpy(o.data(), out, 16384, cudaMemcpyDeviceToHost);
  // un-swizzle assuming address-based SW128 and print, for the first few 128-byte rows, (row,col) of each pixel slot
  for (int m = 0; m < 12; ++m) {
    printf("smem row %2d:", m);
    for (int chunk = 0; chunk < 8; ++chunk) {
      const int phys = chunk ^ (m & 7);
      const __nv_bfloat16* q = &o[(m * 128 + phys * 16) / 2];
      printf(" [%g,%g|%g,%g]", __bfloat162float(q[0]), __bfloat162float(q[1]), __bfloat162float(q[4]), __bfloat162float(q[5]));
    }
    printf("\n");
  }
  printf("expected row m=(ow=m%%8+c2, oh'=m/8+c3): chunks 0-3 = image row 2*oh', pixels 2*ow..2*ow+7; chunks 4-7 = image row 2*oh'+1\n");
  return 0;
}
