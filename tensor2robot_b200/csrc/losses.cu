// losses.cu — the label-sized loss tail of the BC-Z model in ONE launch: several weighted regression / log losses with
// tf.losses semantics, forward value and gradient together.
//
// Reference: research/bcz/model.py:476-585 (training_outputs): per action component a tf.losses.huber_loss /
// mean_squared_error (xyz, quaternion / axis-angle, ...) or tf.losses.log_loss on sigmoid(logit) (target_close,
// stop_token) with weights = component weight * (1 - stop_token), the QuaterNet norm penalty huber(1, |q|) and the
// "first waypoint" diagnostics; every one reduced as Reduction.SUM_BY_NONZERO_WEIGHTS:
//     loss = sum(l_i * w_i) / max(#{w_i != 0}, 1).
// TensorFlow runs ~15 small ops per component; here one block walks the (at most 16) segments, each a flat fp32 array of
// a few thousand elements, and writes loss[s], d loss[s] / d prediction and (log loss) the sigmoid.
#include "common.cuh"

namespace t2r {

__device__ __forceinline__ float block_sum(float v, float* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < int(blockDim.x >> 5); ++i) s += sm[i];
  return s;
}

struct LossSegments {
  T2RLossSegment seg[T2R_MAX_LOSS_SEGMENTS];
  int n;
};

__global__ void __launch_bounds__(256) weighted_losses_kernel(const LossSegments p, float* __restrict__ losses) {
  __shared__ float sm[8];
  float total = 0.f;
  for (int s = 0; s < p.n; ++s) {
    const T2RLossSegment g = p.seg[s];
    float acc = 0.f, cnt = 0.f;
    for (int64_t i = threadIdx.x; i < g.n; i += blockDim.x) {
      const int64_t row = i / g.cols;
      float w = g.weight;
      if (g.row_mask != nullptr) w *= g.row_mask_is_complement ? (1.0f - g.row_mask[row]) : g.row_mask[row];
      if (g.row_mod > 0 && row % g.row_mod != 0) w = 0.f;
      const float y = g.labels != nullptr ? g.labels[i] : g.label_const;
      const float x = g.predictions[i];
      float l, d;
      if (g.kind == T2R_LOSS_HUBER) {
        const float e = x - y, a = fabsf(e), q = fminf(a, g.delta);
        l = 0.5f * q * q + g.delta * (a - q);
        d = fminf(fmaxf(e, -g.delta), g.delta);
      } else if (g.kind == T2R_LOSS_MSE) {
        const float e = x - y;
        l = e * e;
        d = 2.f * e;
      } else {  // T2R_LOSS_SIGMOID_LOG: x is the logit; tf.losses.log_loss(y, sigmoid(x)), epsilon 1e-7
        const float q = 1.0f / (1.0f + __expf(-x));
        const float eps = 1e-7f;
        l = -y * __logf(q + eps) - (1.0f - y) * __logf(1.0f - q + eps);
        d = (-y / (q + eps) + (1.0f - y) / (1.0f - q + eps)) * q * (1.0f - q);
        if (g.sigmoid_out != nullptr) g.sigmoid_out[i] = q;
      }
      acc = fmaf(l, w, acc);
      cnt += (w != 0.f) ? 1.f : 0.f;
      if (g.dpredictions != nullptr) g.dpredictions[i] = d * w;
    }
    const float sum = block_sum(acc, sm);
    const float nonzero = block_sum(cnt, sm);
    const float inv = 1.0f / fmaxf(nonzero, 1.0f);
    if (g.dpredictions != nullptr) {
      __syncthreads();
      for (int64_t i = threadIdx.x; i < g.n; i += blockDim.x) g.dpredictions[i] *= inv;
    }
    if (threadIdx.x == 0) losses[s] = sum * inv;
    if (g.in_total) total += sum * inv;
  }
  if (threadIdx.x == 0) losses[p.n] = total;
}

}  // namespace t2r

extern "C" int32_t t2r_weighted_losses(const T2RLossSegment* segments, int32_t n_segments, float* losses, void* stream) {
  T2R_CHECK_ARG(segments && losses && n_segments >= 1 && n_segments <= T2R_MAX_LOSS_SEGMENTS,
                "weighted_losses: 1..%d segments", T2R_MAX_LOSS_SEGMENTS);
  t2r::LossSegments p;
  p.n = n_segments;
  for (int i = 0; i < n_segments; ++i) {
    const T2RLossSegment& g = segments[i];
    T2R_CHECK_ARG(g.struct_size == sizeof(T2RLossSegment), "weighted_losses: bad T2RLossSegment size");
    T2R_CHECK_ARG(g.predictions && g.n > 0 && g.cols > 0 && g.n % g.cols == 0 && g.kind >= T2R_LOSS_HUBER &&
                      g.kind <= T2R_LOSS_SIGMOID_LOG, "weighted_losses: segment %d is malformed", i);
    p.seg[i] = g;
  }
  t2r::weighted_losses_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, losses);
  T2R_LAUNCH_OK();
  return T2R_OK;
}
