// loss_cem_optim.cu — sigmoid log-loss, on-device CEM (sample / elite refit), Bellman target and
// the fused optimizer + EMA + bf16-cast update.
//
// Reference call sites:
//   log loss    : tf.losses.log_loss, research/qtopt/t2r_models.py:229-239, models/critic_model.py:171-192
//   CEM         : utils/cross_entropy.py:30-154, policies/policies.py:133-169
//   optimizers  : models/optimizers.py:61-146, research/qtopt/optimizer_builder.py:25-96
#include <algorithm>

#include "common.cuh"
#include "philox.cuh"

namespace t2r {

// loss = mean_i -(y log(q+eps) + (1-y) log(1-q+eps)), q = sigmoid(z).
// d loss / d z_i = (1/n) * q(1-q) * ( -(y/(q+eps)) + (1-y)/(1-q+eps) )
__global__ void __launch_bounds__(256) sigmoid_logloss_kernel(const float* __restrict__ logit,
                                                              const float* __restrict__ label,
                                                              float* __restrict__ q_out, float* loss,
                                                              float* __restrict__ dlogit, long long n) {
  const float eps = 1e-7f;
  const float inv_n = 1.0f / float(n);
  float local = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float z = logit[i], y = label[i];
    const float q = 1.0f / (1.0f + expf(-z));
    local += -(y * logf(q + eps) + (1.0f - y) * logf(1.0f - q + eps));
    if (q_out) q_out[i] = q;
    if (dlogit) dlogit[i] = inv_n * q * (1.0f - q) * (-(y / (q + eps)) + (1.0f - y) / (1.0f - q + eps));
  }
  __shared__ float sm[8];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int j = 0; j < 8; ++j) s += sm[j];
    atomicAdd(loss, s * inv_n);
  }
}

__global__ void sigmoid_kernel(const float* __restrict__ z, float* __restrict__ q, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    q[i] = 1.0f / (1.0f + expf(-z[i]));
}

// samples[b,a,d] = mean[b,d] + std[b,d] * z,  z ~ N(0,1) from Philox(seed, index=(b*A+a)*D+d pair)
__global__ void __launch_bounds__(256) cem_sample_kernel(const float* __restrict__ mean,
                                                         const float* __restrict__ stddev,
                                                         float* __restrict__ samples, int B, int A, int D,
                                                         uint64_t seed, uint64_t offset) {
  const long long total = (long long)B * A * D;
  const long long quads = (total + 3) / 4;
  for (long long qd = blockIdx.x * (long long)blockDim.x + threadIdx.x; qd < quads;
       qd += (long long)gridDim.x * blockDim.x) {
    const Philox4 r = philox4x32_10(seed, uint64_t(qd), offset);
    float z[4];
    box_muller(r.v[0], r.v[1], &z[0], &z[1]);
    box_muller(r.v[2], r.v[3], &z[2], &z[3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = qd * 4 + j;
      if (i < total) {
        const int d = int(i % D);
        const long long b = i / ((long long)A * D);
        samples[i] = mean[b * D + d] + stddev[b * D + d] * z[j];
      }
    }
  }
}

// One block per row b.  Rank of sample a under a stable ascending sort =
// #{j : v_j < v_a} + #{j < a : v_j == v_a}; elites are ranks >= A - num_elites
// (utils/cross_entropy.py:90-98: sorted(...)[-num_elites:]).
__global__ void __launch_bounds__(128) cem_refit_kernel(const float* __restrict__ samples,
                                                        const float* __restrict__ values,
                                                        float* __restrict__ mean, float* __restrict__ stddev,
                                                        float* __restrict__ best_value,
                                                        int* __restrict__ best_index, int A, int D,
                                                        int num_elites) {
  extern __shared__ float sm[];  // values[A], elite flag[A] (as float)
  float* v = sm;
  float* elite = sm + A;
  const int b = blockIdx.x;
  for (int a = threadIdx.x; a < A; a += blockDim.x) v[a] = values[(long long)b * A + a];
  __syncthreads();
  for (int a = threadIdx.x; a < A; a += blockDim.x) {
    const float va = v[a];
    int rank = 0;
    for (int j = 0; j < A; ++j) rank += (v[j] < va) || (v[j] == va && j < a);
    elite[a] = (rank >= A - num_elites) ? 1.f : 0.f;
  }
  __syncthreads();
  // np.mean / np.std(ddof=1) over the elites, two-pass for accuracy
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
    for (int a = 0; a < A; ++a)
      if (elite[a] != 0.f) s += samples[((long long)b * A + a) * D + d];
    const float m = s / float(num_elites);
    float ss = 0.f;
    for (int a = 0; a < A; ++a)
      if (elite[a] != 0.f) {
        const float t = samples[((long long)b * A + a) * D + d] - m;
        ss += t * t;
      }
    mean[(long long)b * D + d] = m;
    stddev[(long long)b * D + d] = sqrtf(ss / float(num_elites - 1));
  }
  if (threadIdx.x == 0) {
    int bi = 0;
    float bv = v[0];
    for (int a = 1; a < A; ++a)
      if (v[a] > bv) { bv = v[a]; bi = a; }  // first maximum, like np.argmax
    if (best_value) best_value[b] = bv;
    if (best_index) best_index[b] = bi;
  }
}

__global__ void bellman_target_kernel(const float* __restrict__ reward, const float* __restrict__ done,
                                      const float* __restrict__ max_q, float gamma,
                                      float* __restrict__ target, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    target[i] = reward[i] + gamma * (1.0f - done[i]) * max_q[i];
}

// Fused optimizer step over the flat parameter buffer: gradient unscale (1 / world size) + slim l2 gradient on the
// regularised prefix + the optimizer rule + EMA shadow + bf16 compute copy, ONE pass.  HBM-bound (22 - 34 B per
// parameter), so everything moves as 16-byte vectors: the buffers are 256-byte aligned and every variable is padded
// to 64 elements (nn.VariableStore.finalize), hence n and n_decay are multiples of 4 and a float4 never straddles
// the l2 boundary.  `Rule` holds the slot pointers and maps (w, g) -> new w for one element.
struct MomentumRule {   // TF MomentumOptimizer (use_nesterov=False): accum = momentum*accum + g ; w -= lr*accum
  float* accum;
  float lr, momentum;
  __device__ __forceinline__ void load(long long i4, float4 (&s)[2]) const { s[0] = reinterpret_cast<const float4*>(accum)[i4]; }
  __device__ __forceinline__ void store(long long i4, const float4 (&s)[2]) const { reinterpret_cast<float4*>(accum)[i4] = s[0]; }
  __device__ __forceinline__ float apply(float w, float g, float& s0, float&) const {
    s0 = momentum * s0 + g;
    return w - lr * s0;
  }
};
struct AdamRule {       // TF AdamOptimizer: m, v moments; w -= lr_t * m / (sqrt(v) + eps), lr_t carries the bias corrections
  float *m, *v;
  float lr_t, beta1, beta2, eps;
  __device__ __forceinline__ void load(long long i4, float4 (&s)[2]) const {
    s[0] = reinterpret_cast<const float4*>(m)[i4];
    s[1] = reinterpret_cast<const float4*>(v)[i4];
  }
  __device__ __forceinline__ void store(long long i4, const float4 (&s)[2]) const {
    reinterpret_cast<float4*>(m)[i4] = s[0];
    reinterpret_cast<float4*>(v)[i4] = s[1];
  }
  __device__ __forceinline__ float apply(float w, float g, float& mi, float& vi) const {
    mi = beta1 * mi + (1.0f - beta1) * g;
    vi = beta2 * vi + (1.0f - beta2) * g * g;
    return w - lr_t * mi / (sqrtf(vi) + eps);
  }
};
struct RmsPropRule {    // TF RMSPropOptimizer (not centered): ms = rho*ms + (1-rho)*g^2; mom = momentum*mom + lr*g/sqrt(ms+eps); w -= mom
  float *ms, *mom;
  float lr, rho, momentum, eps;
  __device__ __forceinline__ void load(long long i4, float4 (&s)[2]) const {
    s[0] = reinterpret_cast<const float4*>(ms)[i4];
    s[1] = reinterpret_cast<const float4*>(mom)[i4];
  }
  __device__ __forceinline__ void store(long long i4, const float4 (&s)[2]) const {
    reinterpret_cast<float4*>(ms)[i4] = s[0];
    reinterpret_cast<float4*>(mom)[i4] = s[1];
  }
  __device__ __forceinline__ float apply(float w, float g, float& msi, float& momi) const {
    msi = rho * msi + (1.0f - rho) * g * g;
    momi = momentum * momi + lr * g / sqrtf(msi + eps);
    return w - momi;
  }
};

template <class Rule>
__global__ void __launch_bounds__(256) optimizer_kernel(float* __restrict__ w, const float* __restrict__ g,
                                                        float* __restrict__ ema, __nv_bfloat16* __restrict__ wb,
                                                        long long n4, long long n_decay4, float l2, float grad_scale,
                                                        float ema_decay, Rule rule) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 w4 = reinterpret_cast<const float4*>(w)[i];
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 s[2];
    rule.load(i, s);
    float wv[4] = {w4.x, w4.y, w4.z, w4.w};
    const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
    float s0[4] = {s[0].x, s[0].y, s[0].z, s[0].w}, s1[4] = {s[1].x, s[1].y, s[1].z, s[1].w};
    const float decay = i < n_decay4 ? l2 : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) wv[j] = rule.apply(wv[j], fmaf(decay, wv[j], gv[j] * grad_scale), s0[j], s1[j]);
    s[0] = make_float4(s0[0], s0[1], s0[2], s0[3]);
    s[1] = make_float4(s1[0], s1[1], s1[2], s1[3]);
    rule.store(i, s);
    reinterpret_cast<float4*>(w)[i] = make_float4(wv[0], wv[1], wv[2], wv[3]);
    if (ema) {
      float4 e = reinterpret_cast<const float4*>(ema)[i];
      e.x = ema_decay * e.x + (1.0f - ema_decay) * wv[0];
      e.y = ema_decay * e.y + (1.0f - ema_decay) * wv[1];
      e.z = ema_decay * e.z + (1.0f - ema_decay) * wv[2];
      e.w = ema_decay * e.w + (1.0f - ema_decay) * wv[3];
      reinterpret_cast<float4*>(ema)[i] = e;
    }
    if (wb) reinterpret_cast<uint2*>(wb)[i] = make_uint2(pack_bf16(wv[0], wv[1]), pack_bf16(wv[2], wv[3]));
  }
}

// one full wave of 256-thread blocks, 4 parameters per thread and iteration
static inline int opt_grid(long long n4) {
  return int(std::min<long long>(std::max<long long>((n4 + 255) / 256, 1), (long long)num_sms() * 8));
}

static inline int grid_for(long long n) {
  return int(std::min<long long>(std::max<long long>((n + 255) / 256, 1), 148LL * 16));
}

}  // namespace t2r

using namespace t2r;

extern "C" int32_t t2r_sigmoid_logloss(const float* logit, const float* label, float* q, float* loss,
                                       float* dlogit, int64_t n, void* stream) {
  T2R_CHECK_ARG(logit && label && loss && n > 0, "sigmoid_logloss: bad args");
  sigmoid_logloss_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(logit, label, q, loss,
                                                                                      dlogit, n);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_sigmoid_f32(const float* logit, float* q, int64_t n, void* stream) {
  T2R_CHECK_ARG(logit && q && n > 0, "sigmoid: bad args");
  sigmoid_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(logit, q, n);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_cem_sample(const float* mean, const float* stddev, float* samples, int32_t B,
                                  int32_t A, int32_t D, uint64_t seed, uint64_t offset, void* stream) {
  T2R_CHECK_ARG(mean && stddev && samples && B > 0 && A > 0 && D > 0, "cem_sample: bad args");
  const long long quads = ((long long)B * A * D + 3) / 4;
  cem_sample_kernel<<<grid_for(quads), 256, 0, static_cast<cudaStream_t>(stream)>>>(mean, stddev, samples, B,
                                                                                    A, D, seed, offset);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_cem_refit(const float* samples, const float* values, float* mean, float* stddev,
                                 float* best_value, int32_t* best_index, int32_t B, int32_t A, int32_t D,
                                 int32_t num_elites, void* stream) {
  T2R_CHECK_ARG(samples && values && mean && stddev && B > 0 && D > 0, "cem_refit: bad args");
  T2R_CHECK_ARG(A > 0 && A <= 1024 && num_elites >= 2 && num_elites <= A, "cem_refit: A=%d elites=%d", A,
                num_elites);
  cem_refit_kernel<<<B, 128, sizeof(float) * 2 * A, static_cast<cudaStream_t>(stream)>>>(
      samples, values, mean, stddev, best_value, best_index, A, D, num_elites);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_bellman_target(const float* reward, const float* done, const float* max_q,
                                      float gamma, float* target, int64_t n, void* stream) {
  T2R_CHECK_ARG(reward && done && max_q && target && n > 0, "bellman_target: bad args");
  bellman_target_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(reward, done, max_q,
                                                                                     gamma, target, n);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

static bool vector_ok(const void* a, const void* b, const void* c, int64_t n, int64_t n_decay) {
  return n % 4 == 0 && n_decay % 4 == 0 && (reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                                            reinterpret_cast<uintptr_t>(c)) % 16 == 0;
}

extern "C" int32_t t2r_momentum_step(float* w, const float* g, float* accum, float* ema, void* w_bf16,
                                     int64_t n, int64_t n_decay, float lr, float momentum, float l2,
                                     float grad_scale, float ema_decay, void* stream) {
  T2R_CHECK_ARG(w && g && accum && n > 0 && n_decay >= 0 && n_decay <= n, "momentum_step: bad args");
  T2R_CHECK_ARG(vector_ok(w, g, accum, n, n_decay), "momentum_step: buffers must be 16-byte aligned, n and n_decay multiples of 4");
  const MomentumRule rule{accum, lr, momentum};
  optimizer_kernel<<<opt_grid(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, g, ema, static_cast<__nv_bfloat16*>(w_bf16), n / 4, n_decay / 4, l2, grad_scale, ema_decay, rule);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_adam_step(float* w, const float* g, float* m, float* v, float* ema, void* w_bf16,
                                 int64_t n, int64_t n_decay, float lr, float beta1, float beta2, float eps,
                                 int64_t step, float l2, float grad_scale, float ema_decay, void* stream) {
  T2R_CHECK_ARG(w && g && m && v && n > 0 && step >= 1 && n_decay >= 0 && n_decay <= n, "adam_step: bad args");
  T2R_CHECK_ARG(vector_ok(w, g, m, n, n_decay) && reinterpret_cast<uintptr_t>(v) % 16 == 0,
                "adam_step: buffers must be 16-byte aligned, n and n_decay multiples of 4");
  const double lr_t = double(lr) * sqrt(1.0 - pow(double(beta2), double(step))) /
                      (1.0 - pow(double(beta1), double(step)));
  const AdamRule rule{m, v, float(lr_t), beta1, beta2, eps};
  optimizer_kernel<<<opt_grid(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, g, ema, static_cast<__nv_bfloat16*>(w_bf16), n / 4, n_decay / 4, l2, grad_scale, ema_decay, rule);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_rmsprop_step(float* w, const float* g, float* ms, float* mom, float* ema, void* w_bf16,
                                    int64_t n, int64_t n_decay, float lr, float decay, float momentum, float eps,
                                    float l2, float grad_scale, float ema_decay, void* stream) {
  T2R_CHECK_ARG(w && g && ms && mom && n > 0 && n_decay >= 0 && n_decay <= n, "rmsprop_step: bad args");
  T2R_CHECK_ARG(vector_ok(w, g, ms, n, n_decay) && reinterpret_cast<uintptr_t>(mom) % 16 == 0,
                "rmsprop_step: buffers must be 16-byte aligned, n and n_decay multiples of 4");
  const RmsPropRule rule{ms, mom, lr, decay, momentum, eps};
  optimizer_kernel<<<opt_grid(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, g, ema, static_cast<__nv_bfloat16*>(w_bf16), n / 4, n_decay / 4, l2, grad_scale, ema_decay, rule);
  T2R_LAUNCH_OK();
  return T2R_OK;
}
