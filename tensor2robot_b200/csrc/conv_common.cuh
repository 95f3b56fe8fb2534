// conv_common.cuh — declarations shared by the tcgen05 convolution kernels.
#pragma once
#include "common.cuh"

namespace t2r {

// One filter tap of an implicit-GEMM convolution.
struct ConvTap {
  int8_t map;       // which activation tensor map (stride phase)
  int8_t dh, dw;    // coordinate offset inside that map
  int8_t pad;
  int32_t kchunk0;  // first 64-wide K chunk of this tap in the weight matrix
};
constexpr int kMaxTaps = 32;

// Pick the TW x TH = `pixels` rectangle that wastes the fewest padded pixels.
void pick_tile(int Ho, int Wo, int pixels, int* TW, int* TH);
// Tensor maps (box 64 x TW x TH x 1, 128B swizzle) for the stride*stride phases of a
// [N,H,W,C] bf16 tensor; unused slots are filled with map 0.
int make_phase_maps(CUtensorMap* maps, const void* x, int N, int H, int W, int C, int stride,
                    int TW, int TH);

// Stride-1 KxK convolution through the shared-memory halo kernel (conv_halo.cu).
struct HaloRequest {
  const void* x;       // [N,H,W,C] bf16, the tensor the taps slide over
  int N, H, W, C;
  const void* w;       // [Cout][Ktot] bf16, K-major
  uint64_t Ktot;
  ConvTap taps[kMaxTaps];
  int n_taps;
  int Ho, Wo, Cout;    // output iteration space and its channel count
  long long os_n, os_h, os_w;
  void* out;
  const void* residual;
  const float* bias;
  int flags;
  double* stats;       // optional fused bn_stats output, fp64 [2*Cout]
};
bool conv_halo_eligible(int stride, int n_taps, int k_channels, int n_channels);
int conv_halo_launch(const HaloRequest& r, cudaStream_t stream);

}  // namespace t2r
