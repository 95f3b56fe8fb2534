// metric.cu — n-pairs metric-learning loss (Grasp2Vec, research/grasp2vec/losses.py:152-181), forward and
// backward in one call.  The loss itself lives in a third-party dependency that is absent from
// /root/reference (tf.contrib.losses.metric_learning.npairs_loss, TF 1.15 / tf_slim); its published
// algorithm, for labels = range(B) as the reference passes them:
//
//   sim      = anchor @ positive^T                                   [B, B]
//   xent     = mean_i( logsumexp_j sim[i, j] - sim[i, i] )
//   l2       = 0.25 * reg_lambda * ( mean_i |anchor_i|^2 + mean_i |positive_i|^2 )
//   loss     = xent + l2
//
// Small fp32 problem (B x D = 256 x 1024): CUDA-core GEMMs (t2r_sgemm) + one row-softmax kernel.
#include <algorithm>

#include "common.cuh"

namespace t2r {

// One block per row i: stable log-sum-exp, the row's cross entropy, and dsim[i, :] = (softmax - onehot_i) / B
// written in place.  row_loss[i] = lse_i - sim[i, i].
__global__ void __launch_bounds__(256) npairs_rows_kernel(float* __restrict__ sim, float* __restrict__ row_loss,
                                                          int B) {
  __shared__ float sm[8];
  __shared__ float bcast;
  const int i = blockIdx.x;
  float* row = sim + (long long)i * B;
  float m = -INFINITY;
  for (int j = threadIdx.x; j < B; j += 256) m = fmaxf(m, row[j]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = sm[0];
    for (int k = 1; k < 8; ++k) v = fmaxf(v, sm[k]);
    bcast = v;
  }
  __syncthreads();
  m = bcast;
  float s = 0.f;
  for (int j = threadIdx.x; j < B; j += 256) s += expf(row[j] - m);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int k = 0; k < 8; ++k) v += sm[k];
    bcast = v;
    row_loss[i] = m + logf(v) - row[i];
  }
  __syncthreads();
  s = bcast;
  const float inv_b = 1.f / float(B);
  for (int j = threadIdx.x; j < B; j += 256) {
    const float p = expf(row[j] - m) / s;
    row[j] = (p - (j == i ? 1.f : 0.f)) * inv_b;
  }
}

// loss = mean(row_loss) + 0.25 * lambda * (sum(a^2) + sum(p^2)) / B        (single block)
__global__ void __launch_bounds__(256) npairs_finalize_kernel(const float* __restrict__ row_loss,
                                                              const float* __restrict__ a, const float* __restrict__ p,
                                                              int B, int D, float reg_lambda, float* loss) {
  __shared__ float sm[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < B; i += 256) acc += row_loss[i];
  float sq = 0.f;
  for (long long i = threadIdx.x; i < (long long)B * D; i += 256) sq += a[i] * a[i] + p[i] * p[i];
  acc += 0.25f * reg_lambda * sq;
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int k = 0; k < 8; ++k) v += sm[k];
    loss[0] = v / float(B);
  }
}

__global__ void relu_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = (__bfloat162float(x[i]) > 0.f) ? x[i] : __float2bfloat16_rn(0.f);
}

}  // namespace t2r

using namespace t2r;

extern "C" int32_t t2r_npairs_loss(const float* anchor, const float* positive, int32_t B, int32_t D, float reg_lambda,
                                   float* sim_ws, float* row_ws, float* loss, float* d_anchor, float* d_positive,
                                   void* stream) {
  T2R_CHECK_ARG(anchor && positive && sim_ws && row_ws && loss && d_anchor && d_positive && B > 0 && D > 0,
                "npairs_loss: bad args");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // sim = anchor @ positive^T
  if (int rc = t2r_sgemm(0, 1, B, B, D, 1.f, anchor, D, positive, D, 0.f, sim_ws, B, stream)) return rc;
  npairs_rows_kernel<<<B, 256, 0, st>>>(sim_ws, row_ws, B);
  T2R_LAUNCH_OK();
  npairs_finalize_kernel<<<1, 256, 0, st>>>(row_ws, anchor, positive, B, D, reg_lambda, loss);
  T2R_LAUNCH_OK();
  // d_anchor = dsim @ positive + (0.5 * lambda / B) * anchor ; d_positive = dsim^T @ anchor + (0.5 * lambda / B) * positive
  const float reg = 0.5f * reg_lambda / float(B);
  T2R_CUDA_OK(cudaMemcpyAsync(d_anchor, anchor, sizeof(float) * size_t(B) * D, cudaMemcpyDeviceToDevice, st));
  T2R_CUDA_OK(cudaMemcpyAsync(d_positive, positive, sizeof(float) * size_t(B) * D, cudaMemcpyDeviceToDevice, st));
  if (int rc = t2r_sgemm(0, 0, B, D, B, 1.f, sim_ws, B, positive, D, reg, d_anchor, D, stream)) return rc;
  if (int rc = t2r_sgemm(1, 0, B, D, B, 1.f, sim_ws, B, anchor, D, reg, d_positive, D, stream)) return rc;
  return T2R_OK;
}

extern "C" int32_t t2r_relu_fwd_bf16(const void* x, void* y, int64_t n, void* stream) {
  T2R_CHECK_ARG(x && y && n > 0, "relu_fwd: bad args");
  const int grid = int(std::min<long long>((n + 255) / 256, 148LL * 16));
  relu_fwd_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x),
                                                                            static_cast<__nv_bfloat16*>(y), n);
  T2R_LAUNCH_OK();
  return T2R_OK;
}
