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

// ---------------------------------------------------------------------------------------------
// Triplet loss with semi-hard negative mining (Grasp2Vec TripletLoss, research/grasp2vec/losses.py:51-71 ->
// tf.contrib.losses.metric_learning.triplet_semihard_loss, absent third-party code whose mining the
// reference restates in layers/tec.py:322-383):
//   D[i,j]   = max(|e_i|^2 + |e_j|^2 - 2 e_i.e_j, 0), D[i,i] = 0            (pairwise_distance, squared)
//   for every anchor a and positive p (label[p] == label[a], p != a):
//     outside = min{ D[a,n] : label[n] != label[a], D[a,n] > D[a,p] }          (if that set is not empty)
//     inside  = max{ D[a,n] : label[n] != label[a] }
//     loss   += max(margin + D[a,p] - (outside if it exists else inside), 0)
//   loss /= number of (a, p) pairs
// The backward pass is dD (the +-1/num_pos coefficients of the active pairs) pushed through D.
// ---------------------------------------------------------------------------------------------
namespace t2r {

__global__ void pairwise_sqdist_kernel(const float* __restrict__ gram, float* __restrict__ dist, int M) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= M) return;
  const float d = gram[(long long)i * M + i] + gram[(long long)j * M + j] - 2.f * gram[(long long)i * M + j];
  dist[(long long)i * M + j] = (i == j) ? 0.f : fmaxf(d, 0.f);
}

// One block per anchor.  coef[a, :] receives d loss / d D[a, :] (before the 1 / num_pos factor); stats[0]
// accumulates the loss sum, stats[1] the number of positive pairs.
__global__ void __launch_bounds__(256) triplet_semihard_kernel(const float* __restrict__ dist, const int* __restrict__ labels,
                                                               float margin, int M, float* __restrict__ coef,
                                                               float* __restrict__ stats) {
  __shared__ float s_val[256];
  __shared__ int s_idx[256];
  const int a = blockIdx.x;
  const float* row = dist + (long long)a * M;
  const int la = labels[a];
  for (int j = threadIdx.x; j < M; j += 256) coef[(long long)a * M + j] = 0.f;
  __syncthreads();
  // inside: the largest negative distance (and where it is)
  float best = -INFINITY;
  int besti = -1;
  for (int n = threadIdx.x; n < M; n += 256)
    if (labels[n] != la && row[n] > best) { best = row[n]; besti = n; }
  s_val[threadIdx.x] = best; s_idx[threadIdx.x] = besti;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const float v = s_val[threadIdx.x + s];
      const int vi = s_idx[threadIdx.x + s];
      if (vi >= 0 && (s_idx[threadIdx.x] < 0 || v > s_val[threadIdx.x] ||
                      (v == s_val[threadIdx.x] && vi < s_idx[threadIdx.x]))) {
        s_val[threadIdx.x] = v; s_idx[threadIdx.x] = vi;
      }
    }
    __syncthreads();
  }
  const float inside = s_val[0];
  const int inside_i = s_idx[0];
  __syncthreads();
  for (int p = 0; p < M; ++p) {
    if (p == a || labels[p] != la) continue;          // block-uniform
    const float dap = row[p];
    float o = INFINITY;
    int oi = -1;
    for (int n = threadIdx.x; n < M; n += 256)
      if (labels[n] != la && row[n] > dap && row[n] < o) { o = row[n]; oi = n; }
    s_val[threadIdx.x] = o; s_idx[threadIdx.x] = oi;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) {
        const float v = s_val[threadIdx.x + s];
        const int vi = s_idx[threadIdx.x + s];
        if (vi >= 0 && (s_idx[threadIdx.x] < 0 || v < s_val[threadIdx.x] ||
                        (v == s_val[threadIdx.x] && vi < s_idx[threadIdx.x]))) {
          s_val[threadIdx.x] = v; s_idx[threadIdx.x] = vi;
        }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const int ni = s_idx[0] >= 0 ? s_idx[0] : inside_i;
      const float neg = s_idx[0] >= 0 ? s_val[0] : inside;
      atomicAdd(stats + 1, 1.f);
      if (ni >= 0) {
        const float l = margin + dap - neg;
        if (l > 0.f) {
          atomicAdd(stats, l);
          coef[(long long)a * M + p] += 1.f;
          coef[(long long)a * M + ni] -= 1.f;
        }
      }
    }
    __syncthreads();
  }
}

// S = (coef + coef^T) masked where the distance was clamped; also rowsum[i] = sum_j S[i, j]; in place into coef2.
__global__ void triplet_sym_kernel(const float* __restrict__ coef, const float* __restrict__ dist, float* __restrict__ sym,
                                   float* __restrict__ rowsum, int M) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= M) return;
  float v = coef[(long long)i * M + j] + coef[(long long)j * M + i];
  if (i == j || !(dist[(long long)i * M + j] > 0.f)) v = 0.f;      // clamped / diagonal entries carry no gradient
  sym[(long long)i * M + j] = v;
  atomicAdd(rowsum + i, v);
}

// dE[i, :] = scale * (rowsum[i] * E[i, :] - (S @ E)[i, :]),  scale = 2 / num_pos; loss = stats[0] / stats[1]
__global__ void triplet_finish_kernel(const float* __restrict__ emb, const float* __restrict__ rowsum,
                                      const float* __restrict__ stats, float* __restrict__ d_emb, float* __restrict__ loss,
                                      long long total, int D) {
  const float np_ = fmaxf(stats[1], 1.f);
  const float scale = 2.f / np_;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    d_emb[i] = scale * (rowsum[i / D] * emb[i] - d_emb[i]);
  if (blockIdx.x == 0 && threadIdx.x == 0) loss[0] = stats[0] / np_;
}

}  // namespace t2r

extern "C" int32_t t2r_triplet_semihard_loss(const float* emb, const int32_t* labels, int32_t M, int32_t D, float margin,
                                             float* ws /* 3*M*M + M + 2 floats */, float* loss, float* d_emb,
                                             void* stream) {
  T2R_CHECK_ARG(emb && labels && ws && loss && d_emb && M > 1 && D > 0, "triplet_semihard_loss: bad args");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* gram = ws;                          // [M, M], reused for coef
  float* dist = ws + (size_t)M * M;          // [M, M]
  float* sym = ws + 2 * (size_t)M * M;       // [M, M]
  float* rowsum = ws + 3 * (size_t)M * M;    // [M]
  float* stats = rowsum + M;                 // [2]
  if (int rc = t2r_sgemm(0, 1, M, M, D, 1.f, emb, D, emb, D, 0.f, gram, M, stream)) return rc;
  const dim3 grid2((M + 127) / 128, M);
  pairwise_sqdist_kernel<<<grid2, 128, 0, st>>>(gram, dist, M);
  T2R_LAUNCH_OK();
  T2R_CUDA_OK(cudaMemsetAsync(rowsum, 0, sizeof(float) * (size_t(M) + 2), st));
  triplet_semihard_kernel<<<M, 256, 0, st>>>(dist, labels, margin, M, gram /* coef */, stats);
  T2R_LAUNCH_OK();
  triplet_sym_kernel<<<grid2, 128, 0, st>>>(gram, dist, sym, rowsum, M);
  T2R_LAUNCH_OK();
  // d_emb <- S @ E, then finished in place
  if (int rc = t2r_sgemm(0, 0, M, D, M, 1.f, sym, M, emb, D, 0.f, d_emb, D, stream)) return rc;
  const long long total = (long long)M * D;
  triplet_finish_kernel<<<int(std::min<long long>((total + 255) / 256, 148LL * 8)), 256, 0, st>>>(emb, rowsum, stats,
                                                                                                   d_emb, loss, total, D);
  T2R_LAUNCH_OK();
  return T2R_OK;
}
