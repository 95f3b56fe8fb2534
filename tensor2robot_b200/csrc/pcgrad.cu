// PCGrad gradient surgery (research/qtopt/pcgrad.py:123-154, the per-variable implementation) on the flat
// gradient buffers of the engine.
//
// The reference projects, per variable v and task t, grad = g_t sequentially against every task gradient g_k:
//   grad -= min(<grad, g_k> / (|g_k|^2 + 1e-5), 0) * g_k,   and sums the T results.
// Every intermediate `grad` stays in span{g_0..g_{T-1}}, so all inner products follow from the T x T Gram matrix
// of the variable: instead of T^2 dependent reductions over the variable, ONE pass computes the Gram matrices of all
// variables (HBM-bound: T reads of the buffer), a tiny kernel runs the sequential projections on coefficients, and one
// more pass writes out = sum_j a[v][j] * g_j.  Variables outside the allow / deny lists get a = 1 (the plain sum of
// the task gradients, pcgrad.py:108-116).
#include "common.cuh"

namespace t2r {

constexpr int kMaxTasks = 8;

// grid (V, slices): block (v, s) strides over segment v.  gram[v][j][k] (j <= k filled, mirrored by the coefficient
// kernel) accumulates through atomics; zeroed by the caller.
__global__ void __launch_bounds__(256) pcgrad_gram_kernel(const float* __restrict__ grads, int T, long long stride,
                                                          const long long* __restrict__ seg_off,
                                                          const long long* __restrict__ seg_len,
                                                          float* __restrict__ gram) {
  __shared__ float sm[8];
  const int v = blockIdx.x;
  const long long off = seg_off[v], len = seg_len[v];
  float acc[kMaxTasks * (kMaxTasks + 1) / 2];
#pragma unroll
  for (int i = 0; i < kMaxTasks * (kMaxTasks + 1) / 2; ++i) acc[i] = 0.f;
  for (long long i = (long long)blockIdx.y * 256 + threadIdx.x; i < len; i += (long long)gridDim.y * 256) {
    float g[kMaxTasks];
#pragma unroll
    for (int j = 0; j < kMaxTasks; ++j) g[j] = j < T ? grads[(long long)j * stride + off + i] : 0.f;
    int q = 0;
#pragma unroll
    for (int j = 0; j < kMaxTasks; ++j)
#pragma unroll
      for (int k = j; k < kMaxTasks; ++k, ++q) acc[q] = fmaf(g[j], g[k], acc[q]);
  }
  int q = 0;
#pragma unroll
  for (int j = 0; j < kMaxTasks; ++j)
#pragma unroll
    for (int k = j; k < kMaxTasks; ++k, ++q) {
      if (k >= T) continue;                     // warp-uniform
      float s = acc[q];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      __syncthreads();
      if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
      __syncthreads();
      if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += sm[w];
        atomicAdd(gram + ((long long)v * T + j) * T + k, t);
      }
    }
}

// One thread per variable: the reference's nested task loops on coefficient vectors.
__global__ void pcgrad_coef_kernel(const float* __restrict__ gram, const unsigned char* __restrict__ use_pcgrad, int T,
                                   int V, float eps, float* __restrict__ coef) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  float a[kMaxTasks];
  if (use_pcgrad != nullptr && !use_pcgrad[v]) {
    for (int j = 0; j < T; ++j) coef[(long long)v * T + j] = 1.f;
    return;
  }
  const float* G = gram + (long long)v * T * T;
  for (int j = 0; j < T; ++j) a[j] = 0.f;
  for (int t = 0; t < T; ++t) {
    float c[kMaxTasks];
    for (int j = 0; j < T; ++j) c[j] = j == t ? 1.f : 0.f;
    for (int k = 0; k < T; ++k) {
      float ip = 0.f;
      for (int j = 0; j < T; ++j) ip = fmaf(c[j], j <= k ? G[j * T + k] : G[k * T + j], ip);
      const float pd = ip / (G[k * T + k] + eps);
      if (pd < 0.f) c[k] -= pd;
    }
    for (int j = 0; j < T; ++j) a[j] += c[j];
  }
  for (int j = 0; j < T; ++j) coef[(long long)v * T + j] = a[j];
}

__global__ void __launch_bounds__(256) pcgrad_combine_kernel(const float* __restrict__ grads, int T, long long stride,
                                                             const long long* __restrict__ seg_off,
                                                             const long long* __restrict__ seg_len,
                                                             const float* __restrict__ coef, float* __restrict__ out) {
  const int v = blockIdx.x;
  const long long off = seg_off[v], len = seg_len[v];
  float a[kMaxTasks];
#pragma unroll
  for (int j = 0; j < kMaxTasks; ++j) a[j] = j < T ? coef[(long long)v * T + j] : 0.f;
  for (long long i = (long long)blockIdx.y * 256 + threadIdx.x; i < len; i += (long long)gridDim.y * 256) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxTasks; ++j)
      if (j < T) s = fmaf(a[j], grads[(long long)j * stride + off + i], s);
    out[off + i] = s;
  }
}

}  // namespace t2r

using namespace t2r;

extern "C" int32_t t2r_pcgrad_project(const float* grads, int32_t T, int64_t stride, const int64_t* seg_off,
                                      const int64_t* seg_len, const uint8_t* use_pcgrad, int32_t V, float eps, float* gram,
                                      float* coef, float* out, void* stream) {
  T2R_CHECK_ARG(grads && seg_off && seg_len && gram && coef && out && T >= 1 && T <= kMaxTasks && V >= 1 && V <= 65535 * 32 &&
                    stride > 0,
                "pcgrad_project: bad args (1 <= tasks <= 8)");
  static_assert(sizeof(long long) == sizeof(int64_t), "segment tables are int64");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  T2R_CUDA_OK(cudaMemsetAsync(gram, 0, sizeof(float) * V * T * T, st));
  const dim3 grid(V, 32);
  pcgrad_gram_kernel<<<grid, 256, 0, st>>>(grads, T, stride, reinterpret_cast<const long long*>(seg_off),
                                            reinterpret_cast<const long long*>(seg_len), gram);
  T2R_LAUNCH_OK();
  pcgrad_coef_kernel<<<(V + 127) / 128, 128, 0, st>>>(gram, use_pcgrad, T, V, eps, coef);
  T2R_LAUNCH_OK();
  pcgrad_combine_kernel<<<grid, 256, 0, st>>>(grads, T, stride, reinterpret_cast<const long long*>(seg_off),
                                               reinterpret_cast<const long long*>(seg_len), coef, out);
  T2R_LAUNCH_OK();
  return T2R_OK;
}
