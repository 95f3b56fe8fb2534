// fp32 kernels of the small pose_env networks (layers/vision_layers.py:30-158, 277-350;
// research/pose_env/pose_env_models.py:118-181; research/dql_grasping_lib/tf_modules.py:25-93):
// 3x3 convolutions with 3 / 32 channels on 64x64 frames, slim layer_norm, the fp32 spatial softmax and the
// tile + broadcast-add action merge.  These layers have 32 channels and a few MFLOP per image: far below one
// tensor-core tile, so they are plain SIMT kernels with coalesced channel-fastest accesses; everything with
// >= 64 channels runs on the tcgen05 kernels of conv_igemm*.cu instead.
#include "common.cuh"

namespace t2r {

// ---------------------------------------------------------------------------------------------
// Direct convolution, NHWC x HWIO -> NHWC, one thread per output element (output channel fastest, so the
// weight reads are coalesced and the image reads are warp broadcasts).
// ---------------------------------------------------------------------------------------------
struct DirectConv {
  int N, H, W, Cin, Cout, KH, KW, stride, pt, pl, Ho, Wo;
};

__global__ void __launch_bounds__(256) conv_direct_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ y,
                                                              DirectConv d) {
  const long long total = (long long)d.N * d.Ho * d.Wo * d.Cout;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int co = int(i % d.Cout);
    long long p = i / d.Cout;
    const int wo = int(p % d.Wo);
    p /= d.Wo;
    const int ho = int(p % d.Ho), n = int(p / d.Ho);
    float acc = bias != nullptr ? bias[co] : 0.f;
    for (int kh = 0; kh < d.KH; ++kh) {
      const int h = ho * d.stride - d.pt + kh;
      if (h < 0 || h >= d.H) continue;
      for (int kw = 0; kw < d.KW; ++kw) {
        const int ww = wo * d.stride - d.pl + kw;
        if (ww < 0 || ww >= d.W) continue;
        const float* xp = x + (((long long)n * d.H + h) * d.W + ww) * d.Cin;
        const float* wp = w + (long long)(kh * d.KW + kw) * d.Cin * d.Cout + co;
        for (int ci = 0; ci < d.Cin; ++ci) acc = fmaf(xp[ci], wp[(long long)ci * d.Cout], acc);
      }
    }
    y[i] = acc;
  }
}

// dx[n,h,w,ci] = sum over the taps that reach (h, w) and over co of dy * w.
__global__ void __launch_bounds__(256) conv_direct_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                float* __restrict__ dx, DirectConv d) {
  const long long total = (long long)d.N * d.H * d.W * d.Cin;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int ci = int(i % d.Cin);
    long long p = i / d.Cin;
    const int ww = int(p % d.W);
    p /= d.W;
    const int h = int(p % d.H), n = int(p / d.H);
    float acc = 0.f;
    for (int kh = 0; kh < d.KH; ++kh) {
      const int th = h + d.pt - kh;
      if (th < 0 || th % d.stride != 0) continue;
      const int ho = th / d.stride;
      if (ho >= d.Ho) continue;
      for (int kw = 0; kw < d.KW; ++kw) {
        const int tw = ww + d.pl - kw;
        if (tw < 0 || tw % d.stride != 0) continue;
        const int wo = tw / d.stride;
        if (wo >= d.Wo) continue;
        const float* gp = dy + (((long long)n * d.Ho + ho) * d.Wo + wo) * d.Cout;
        const float* wp = w + ((long long)(kh * d.KW + kw) * d.Cin + ci) * d.Cout;
        for (int co = 0; co < d.Cout; ++co) acc = fmaf(gp[co], wp[co], acc);
      }
    }
    dx[i] = acc;
  }
}

// dw[kh,kw,ci,co]: block (x = tap * Cin + ci, y = slice of the output pixels); thread = (pixel lane, co);
// partial sums meet in shared memory, slices meet through atomics (dw is zeroed by the caller).
__global__ void __launch_bounds__(256) conv_direct_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                float* __restrict__ dw, DirectConv d) {
  __shared__ float sm[256];
  const int ci = blockIdx.x % d.Cin, tap = blockIdx.x / d.Cin, kh = tap / d.KW, kw = tap % d.KW;
  const int lanes = 256 / d.Cout, co = threadIdx.x % d.Cout, lane = threadIdx.x / d.Cout;
  const long long P = (long long)d.N * d.Ho * d.Wo;
  float acc = 0.f;
  if (lane < lanes) {
    for (long long p = (long long)blockIdx.y * lanes + lane; p < P; p += (long long)gridDim.y * lanes) {
      const int wo = int(p % d.Wo);
      const long long q = p / d.Wo;
      const int ho = int(q % d.Ho), n = int(q / d.Ho);
      const int h = ho * d.stride - d.pt + kh, ww = wo * d.stride - d.pl + kw;
      if (h < 0 || h >= d.H || ww < 0 || ww >= d.W) continue;
      acc = fmaf(x[(((long long)n * d.H + h) * d.W + ww) * d.Cin + ci], dy[p * d.Cout + co], acc);
    }
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  if (lane == 0) {
    for (int l = 1; l < lanes; ++l) acc += sm[l * d.Cout + co];
    atomicAdd(dw + ((long long)tap * d.Cin + ci) * d.Cout + co, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// slim.layer_norm (begin_norm_axis = 1, begin_params_axis = -1, variance_epsilon = 1e-12): moments over all
// D = H*W*C values of a sample, per-channel gamma / beta, optional fused ReLU.  One block per sample.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                       // protects sm against the previous call's readers
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += sm[i];
  return t;
}

__global__ void __launch_bounds__(256) layer_norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ y,
                                                             float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                             int D, int C, float eps, int relu) {
  __shared__ float sm[8];
  const float* xn = x + (long long)blockIdx.x * D;
  float* yn = y + (long long)blockIdx.x * D;
  float s = 0.f;
  for (int i = threadIdx.x; i < D; i += 256) s += xn[i];
  const float mean = block_sum_256(s, sm) / float(D);
  float v = 0.f;
  for (int i = threadIdx.x; i < D; i += 256) {
    const float c = xn[i] - mean;
    v = fmaf(c, c, v);
  }
  const float rstd = rsqrtf(block_sum_256(v, sm) / float(D) + eps);
  for (int i = threadIdx.x; i < D; i += 256) {
    const int c = i % C;
    float o = (xn[i] - mean) * rstd * gamma[c] + beta[c];
    if (relu) o = fmaxf(o, 0.f);
    yn[i] = o;
  }
  if (threadIdx.x == 0) {
    mean_out[blockIdx.x] = mean;
    rstd_out[blockIdx.x] = rstd;
  }
}

// dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)), g = dy masked by the ReLU;
// dgamma[c] += sum g*xhat, dbeta[c] += sum g (shared-memory partials per block, then global atomics).
__global__ void __launch_bounds__(256) layer_norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                             float* __restrict__ dx, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, int D, int C, int relu) {
  extern __shared__ float smem[];        // [8] reduction scratch, [C] dgamma, [C] dbeta
  float* sm = smem;
  float* sg = smem + 8;
  float* sb = sg + C;
  const float* xn = x + (long long)blockIdx.x * D;
  const float* gn = dy + (long long)blockIdx.x * D;
  float* dn = dx + (long long)blockIdx.x * D;
  const float mean = mean_in[blockIdx.x], rstd = rstd_in[blockIdx.x];
  for (int c = threadIdx.x; c < C; c += 256) {
    sg[c] = 0.f;
    sb[c] = 0.f;
  }
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < D; i += 256) {
    const int c = i % C;
    const float xh = (xn[i] - mean) * rstd;
    float g = gn[i];
    if (relu && xh * gamma[c] + beta[c] <= 0.f) g = 0.f;
    const float gh = g * gamma[c];
    s1 += gh;
    s2 = fmaf(gh, xh, s2);
    if (dgamma != nullptr) {
      atomicAdd(sg + c, g * xh);
      atomicAdd(sb + c, g);
    }
  }
  s1 = block_sum_256(s1, sm) / float(D);
  s2 = block_sum_256(s2, sm) / float(D);
  for (int i = threadIdx.x; i < D; i += 256) {
    const int c = i % C;
    const float xh = (xn[i] - mean) * rstd;
    float g = gn[i];
    if (relu && xh * gamma[c] + beta[c] <= 0.f) g = 0.f;
    dn[i] = rstd * (g * gamma[c] - s1 - xh * s2);
  }
  __syncthreads();
  if (dgamma != nullptr)
    for (int c = threadIdx.x; c < C; c += 256) {
      atomicAdd(dgamma + c, sg[c]);
      atomicAdd(dbeta + c, sb[c]);
    }
}

// ---------------------------------------------------------------------------------------------
// fp32 spatial softmax (layers/spatial_softmax.py:29-88), same definition as the bf16 kernel in pool.cu
// (interleaved x, y outputs): one thread per (image, channel), channel-fastest so the loads coalesce.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) spatial_softmax_f32_fwd_kernel(const float* __restrict__ x, float* __restrict__ points,
                                                                      float* __restrict__ softmax, int H, int W, int C) {
  const int n = blockIdx.x, c = blockIdx.y * 128 + threadIdx.x, HW = H * W;
  if (c >= C) return;
  const float* xn = x + (long long)n * HW * C + c;
  float mx = -INFINITY;
  for (int p = 0; p < HW; ++p) mx = fmaxf(mx, xn[(long long)p * C]);
  const float ax = W > 1 ? 2.f / float(W - 1) : 0.f, ay = H > 1 ? 2.f / float(H - 1) : 0.f;
  float se = 0.f, sx = 0.f, sy = 0.f;
  for (int p = 0; p < HW; ++p) {
    const int i = p / W, j = p - i * W;
    const float px = W > 1 ? ax * float(j) - 1.f : NAN, py = H > 1 ? ay * float(i) - 1.f : NAN;   // 0/0 in the reference
    const float e = expf(xn[(long long)p * C] - mx);
    se += e;
    sx = fmaf(e, px, sx);
    sy = fmaf(e, py, sy);
  }
  points[(long long)n * 2 * C + 2 * c] = sx / se;
  points[(long long)n * 2 * C + 2 * c + 1] = sy / se;
  if (softmax != nullptr) {
    float* sn = softmax + (long long)n * HW * C + c;
    for (int p = 0; p < HW; ++p) sn[(long long)p * C] = expf(xn[(long long)p * C] - mx) / se;
  }
}

__global__ void __launch_bounds__(128) spatial_softmax_f32_bwd_kernel(const float* __restrict__ x, const float* __restrict__ points,
                                                                      const float* __restrict__ dpoints, float* __restrict__ dx,
                                                                      int H, int W, int C) {
  const int n = blockIdx.x, c = blockIdx.y * 128 + threadIdx.x, HW = H * W;
  if (c >= C) return;
  const float* xn = x + (long long)n * HW * C + c;
  float* dn = dx + (long long)n * HW * C + c;
  float mx = -INFINITY;
  for (int p = 0; p < HW; ++p) mx = fmaxf(mx, xn[(long long)p * C]);
  float se = 0.f;
  for (int p = 0; p < HW; ++p) se += expf(xn[(long long)p * C] - mx);
  const long long o = (long long)n * 2 * C + 2 * c;
  const float ex = points[o], ey = points[o + 1], gx = dpoints[o], gy = dpoints[o + 1];
  const float ax = 2.f / float(W - 1), ay = 2.f / float(H - 1);
  for (int p = 0; p < HW; ++p) {
    const int i = p / W, j = p - i * W;
    const float px = ax * float(j) - 1.f, py = ay * float(i) - 1.f;
    const float s = expf(xn[(long long)p * C] - mx) / se;
    dn[(long long)p * C] = s * ((px - ex) * gx + (py - ey) * gy);
  }
}

// ---------------------------------------------------------------------------------------------
// Action merge of the pose_env critic (pose_env_models.py:141-149): y[j] = x[j mod Nx] + ctx[j] broadcast
// over the H*W positions (tf.tile of the whole image batch, then the broadcast add).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_add_context_fwd_kernel(const float* __restrict__ x, const float* __restrict__ ctx,
                                                                   float* __restrict__ y, int Nx, int Nc, int HW, int C) {
  const long long per = (long long)HW * C, total = per * Nc;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int j = int(i / per);
    const long long r = i - (long long)j * per;
    y[i] = x[(long long)(j % Nx) * per + r] + ctx[(long long)j * C + int(r % C)];
  }
}

// dctx[j, c] = sum_p dy[j, p, c]: grid = Nc blocks, threads stride over the channels.
__global__ void __launch_bounds__(256) tile_add_context_dctx_kernel(const float* __restrict__ dy, float* __restrict__ dctx,
                                                                    int HW, int C) {
  const float* gj = dy + (long long)blockIdx.x * HW * C;
  for (int c = threadIdx.x; c < C; c += 256) {
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += gj[(long long)p * C + c];
    dctx[(long long)blockIdx.x * C + c] = s;
  }
}

// dx[i] = sum over the tiles j = i, i + Nx, ... of dy[j].
__global__ void __launch_bounds__(256) tile_add_context_dx_kernel(const float* __restrict__ dy, float* __restrict__ dx, int Nx,
                                                                  int Nc, int HW, int C) {
  const long long per = (long long)HW * C, total = per * Nx;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int n = int(i / per);
    const long long r = i - (long long)n * per;
    float s = 0.f;
    for (int j = n; j < Nc; j += Nx) s += dy[(long long)j * per + r];
    dx[i] = s;
  }
}

__global__ void __launch_bounds__(256) relu_f32_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) y[i] = fmaxf(x[i], 0.f);
}

__global__ void __launch_bounds__(256) relu_f32_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                           float* __restrict__ dx, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

// tf.nn.elu (the MockT2RModel activation, utils/mocks.py:166-170): x > 0 ? x : exp(x) - 1.
__global__ void __launch_bounds__(256) elu_f32_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) y[i] = x[i] > 0.f ? x[i] : expm1f(x[i]);
}

// d elu = x > 0 ? 1 : elu(x) + 1, from the saved output.
__global__ void __launch_bounds__(256) elu_f32_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                          float* __restrict__ dx, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL)
    dx[i] = y[i] > 0.f ? dy[i] : dy[i] * (y[i] + 1.f);
}


// FiLM + ReLU on fp32 NHWC (layers/vision_layers.py:139-141): y = relu((1 + g[n,c]) * x + b[n,c]), film = [N][2C]
// (gammas, then betas).  One block per image, threads over channels (coalesced), loop over pixels.
__global__ void __launch_bounds__(256) film_relu_f32_fwd_kernel(const float* __restrict__ x, const float* __restrict__ film,
                                                                float* __restrict__ y, int HW, int C) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = 1.0f + film[(long long)n * 2 * C + c], b = film[(long long)n * 2 * C + C + c];
    for (int p = blockIdx.y; p < HW; p += gridDim.y) {
      const long long i = ((long long)n * HW + p) * C + c;
      y[i] = fmaxf(fmaf(g, x[i], b), 0.f);
    }
  }
}
// dx = dy * [y > 0] * (1 + g); dfilm[n, c] = sum_p dy * [y > 0] * x; dfilm[n, C + c] = sum_p dy * [y > 0]
__global__ void __launch_bounds__(256) film_relu_f32_bwd_kernel(const float* __restrict__ x, const float* __restrict__ film,
                                                                const float* __restrict__ dy, float* __restrict__ dx,
                                                                float* __restrict__ dfilm, int HW, int C) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = 1.0f + film[(long long)n * 2 * C + c], b = film[(long long)n * 2 * C + C + c];
    float sg = 0.f, sb = 0.f;
    for (int p = 0; p < HW; ++p) {
      const long long i = ((long long)n * HW + p) * C + c;
      const float xv = x[i];
      const float d = fmaf(g, xv, b) > 0.f ? dy[i] : 0.f;
      dx[i] = d * g;
      sg = fmaf(d, xv, sg);
      sb += d;
    }
    dfilm[(long long)n * 2 * C + c] = sg;
    dfilm[(long long)n * 2 * C + C + c] = sb;
  }
}

// slim.batch_norm(is_training=True) on fp32 [rows, C] (the batch-norm normaliser of layers/vision_layers.py:72-86,
// decay .99, eps 1e-4): biased batch variance normalises, the Bessel-corrected one feeds the moving average (fused
// batch norm); optional ReLU.  One block per channel; double accumulation.
__global__ void __launch_bounds__(256) bn_train_f32_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float* __restrict__ y,
                                                               float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                                               float* __restrict__ save_mean, float* __restrict__ save_rstd,
                                                               long long rows, int C, float eps, float decay, int relu) {
  __shared__ double sm[2][8];
  const int c = blockIdx.x;
  double s = 0.0, q = 0.0;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
    const double v = x[r * C + c];
    s += v;
    q += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = s; sm[1][threadIdx.x >> 5] = q; }
  __syncthreads();
  s = q = 0.0;
  for (int i = 0; i < 8; ++i) { s += sm[0][i]; q += sm[1][i]; }
  const double mean = s / double(rows);
  double var = q / double(rows) - mean * mean;
  if (var < 0) var = 0;
  const float rstd = float(1.0 / sqrt(var + double(eps)));
  const float g = gamma != nullptr ? gamma[c] : 1.0f, b = beta[c], m = float(mean);
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
    const float v = (x[r * C + c] - m) * rstd * g + b;
    y[r * C + c] = relu ? fmaxf(v, 0.f) : v;
  }
  if (threadIdx.x == 0) {
    save_mean[c] = m;
    save_rstd[c] = rstd;
    const double unbiased = rows > 1 ? var * double(rows) / double(rows - 1) : var;
    moving_mean[c] = moving_mean[c] * decay + m * (1.0f - decay);
    moving_var[c] = moving_var[c] * decay + float(unbiased) * (1.0f - decay);
  }
}
__global__ void __launch_bounds__(256) bn_train_f32_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                               const float* __restrict__ dy, const float* __restrict__ gamma,
                                                               const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                                                               float* __restrict__ dx, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta, long long rows, int C, int relu) {
  __shared__ double sm[2][8];
  const int c = blockIdx.x;
  const float m = save_mean[c], rstd = save_rstd[c], g = gamma != nullptr ? gamma[c] : 1.0f;
  double sb = 0.0, sg = 0.0;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
    const long long i = r * C + c;
    const float d = (relu && !(y[i] > 0.f)) ? 0.f : dy[i];
    sb += d;
    sg += double(d) * double((x[i] - m) * rstd);
  }
  for (int o = 16; o > 0; o >>= 1) {
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
  }
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = sb; sm[1][threadIdx.x >> 5] = sg; }
  __syncthreads();
  sb = sg = 0.0;
  for (int i = 0; i < 8; ++i) { sb += sm[0][i]; sg += sm[1][i]; }
  const float mb = float(sb / double(rows)), mg = float(sg / double(rows));
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
    const long long i = r * C + c;
    const float d = (relu && !(y[i] > 0.f)) ? 0.f : dy[i];
    const float xh = (x[i] - m) * rstd;
    dx[i] = g * rstd * (d - mb - xh * mg);
  }
  if (threadIdx.x == 0) {
    if (dgamma != nullptr) dgamma[c] = float(sg);
    if (dbeta != nullptr) dbeta[c] = float(sb);
  }
}

// tf.layers.batch_normalization(training=False) on [rows, C] fp32: y = (x - mean) * rsqrt(var + eps) * gamma + beta.
__global__ void __launch_bounds__(256) bn_infer_f32_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, const float* __restrict__ mean,
                                                               const float* __restrict__ var, float* __restrict__ y,
                                                               long long n, int C, float eps) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    const int c = int(i % C);
    y[i] = (x[i] - mean[c]) * rsqrtf(var[c] + eps) * gamma[c] + beta[c];
  }
}

// dx = dy * gamma * rstd; dgamma[c] = sum dy * (x - mean) * rstd; dbeta[c] = sum dy.  One block per 32 channels,
// threads (32 channels x 8 row lanes); rows reduced through shared memory.
__global__ void __launch_bounds__(256) bn_infer_f32_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               const float* __restrict__ gamma, const float* __restrict__ mean,
                                                               const float* __restrict__ var, float* __restrict__ dx,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                               long long rows, int C, float eps) {
  __shared__ float sg[8][33], sb[8][33];
  const int lane = threadIdx.x & 31, part = threadIdx.x >> 5, c = blockIdx.x * 32 + lane;
  float ag = 0.f, ab = 0.f;
  if (c < C) {
    const float rstd = rsqrtf(var[c] + eps), m = mean[c], g = gamma[c];
    for (long long r = part; r < rows; r += 8) {
      const float d = dy[r * C + c], xh = (x[r * C + c] - m) * rstd;
      dx[r * C + c] = d * g * rstd;
      ag = fmaf(d, xh, ag);
      ab += d;
    }
  }
  sg[part][lane] = ag;
  sb[part][lane] = ab;
  __syncthreads();
  if (part == 0 && c < C && dgamma != nullptr) {
    for (int p = 1; p < 8; ++p) {
      ag += sg[p][lane];
      ab += sb[p][lane];
    }
    dgamma[c] = ag;
    dbeta[c] = ab;
  }
}

static int grid_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = 148LL * 16;
  return int(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace t2r

using namespace t2r;

static bool direct_conv_ok(int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t KH, int32_t KW, int32_t stride,
                           int32_t pt, int32_t pl, int32_t Ho, int32_t Wo) {
  return N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && pt >= 0 && pl >= 0 && Ho > 0 &&
         Wo > 0 && (long long)(Ho - 1) * stride - pt < H && (long long)(Wo - 1) * stride - pl < W;
}

extern "C" int32_t t2r_conv2d_direct_f32_fwd(const float* x, const float* w, const float* bias, float* y, int32_t N, int32_t H,
                                             int32_t W, int32_t Cin, int32_t Cout, int32_t KH, int32_t KW, int32_t stride,
                                             int32_t pad_top, int32_t pad_left, int32_t Ho, int32_t Wo, void* stream) {
  T2R_CHECK_ARG(x && w && y && direct_conv_ok(N, H, W, Cin, Cout, KH, KW, stride, pad_top, pad_left, Ho, Wo),
                "conv2d_direct_f32_fwd: bad args");
  const DirectConv d{N, H, W, Cin, Cout, KH, KW, stride, pad_top, pad_left, Ho, Wo};
  conv_direct_fwd_kernel<<<grid_for((long long)N * Ho * Wo * Cout), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, w, bias, y, d);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_conv2d_direct_f32_dgrad(const float* dy, const float* w, float* dx, int32_t N, int32_t H, int32_t W,
                                               int32_t Cin, int32_t Cout, int32_t KH, int32_t KW, int32_t stride, int32_t pad_top,
                                               int32_t pad_left, int32_t Ho, int32_t Wo, void* stream) {
  T2R_CHECK_ARG(dy && w && dx && direct_conv_ok(N, H, W, Cin, Cout, KH, KW, stride, pad_top, pad_left, Ho, Wo),
                "conv2d_direct_f32_dgrad: bad args");
  const DirectConv d{N, H, W, Cin, Cout, KH, KW, stride, pad_top, pad_left, Ho, Wo};
  conv_direct_dgrad_kernel<<<grid_for((long long)N * H * W * Cin), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, w, dx, d);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_conv2d_direct_f32_wgrad(const float* x, const float* dy, float* dw, int32_t N, int32_t H, int32_t W,
                                               int32_t Cin, int32_t Cout, int32_t KH, int32_t KW, int32_t stride, int32_t pad_top,
                                               int32_t pad_left, int32_t Ho, int32_t Wo, void* stream) {
  T2R_CHECK_ARG(x && dy && dw && Cout <= 256 && direct_conv_ok(N, H, W, Cin, Cout, KH, KW, stride, pad_top, pad_left, Ho, Wo),
                "conv2d_direct_f32_wgrad: bad args (Cout <= 256)");
  T2R_CHECK_ARG((long long)KH * KW * Cin < (1LL << 31), "conv2d_direct_f32_wgrad: too many filter rows");
  const DirectConv d{N, H, W, Cin, Cout, KH, KW, stride, pad_top, pad_left, Ho, Wo};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  T2R_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * KH * KW * Cin * Cout, st));
  const long long P = (long long)N * Ho * Wo;
  const int lanes = 256 / Cout;
  long long slices = (P + lanes * 256LL - 1) / (lanes * 256LL);      // ~256 pixels per thread
  slices = slices < 1 ? 1 : (slices > 64 ? 64 : slices);
  conv_direct_wgrad_kernel<<<dim3(KH * KW * Cin, int(slices)), 256, 0, st>>>(x, dy, dw, d);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_layer_norm_f32_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                                          int32_t N, int32_t D, int32_t C, float eps, int32_t relu, void* stream) {
  T2R_CHECK_ARG(x && gamma && beta && y && mean && rstd && N > 0 && D > 0 && C > 0 && D % C == 0, "layer_norm_f32_fwd: bad args");
  layer_norm_fwd_kernel<<<N, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, gamma, beta, y, mean, rstd, D, C, eps, relu);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_layer_norm_f32_bwd(const float* x, const float* dy, const float* gamma, const float* beta, const float* mean,
                                          const float* rstd, float* dx, float* dgamma, float* dbeta, int32_t N, int32_t D,
                                          int32_t C, int32_t relu, void* stream) {
  T2R_CHECK_ARG(x && dy && gamma && beta && mean && rstd && dx && N > 0 && D > 0 && C > 0 && D % C == 0 && C <= 4096 &&
                    ((dgamma == nullptr) == (dbeta == nullptr)),
                "layer_norm_f32_bwd: bad args (C <= 4096)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dgamma != nullptr) {
    T2R_CUDA_OK(cudaMemsetAsync(dgamma, 0, sizeof(float) * C, st));
    T2R_CUDA_OK(cudaMemsetAsync(dbeta, 0, sizeof(float) * C, st));
  }
  layer_norm_bwd_kernel<<<N, 256, sizeof(float) * (8 + 2 * C), st>>>(x, dy, gamma, beta, mean, rstd, dx, dgamma, dbeta, D, C, relu);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_spatial_softmax_f32_fwd(const float* x, float* points, float* softmax, int32_t N, int32_t H, int32_t W,
                                               int32_t C, void* stream) {
  T2R_CHECK_ARG(x && points && N > 0 && H > 0 && W > 0 && C > 0, "spatial_softmax_f32_fwd: bad args");
  spatial_softmax_f32_fwd_kernel<<<dim3(N, (C + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(x, points, softmax, H, W, C);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_spatial_softmax_f32_bwd(const float* x, const float* points, const float* dpoints, float* dx, int32_t N,
                                               int32_t H, int32_t W, int32_t C, void* stream) {
  T2R_CHECK_ARG(x && points && dpoints && dx && N > 0 && H > 1 && W > 1 && C > 0, "spatial_softmax_f32_bwd: bad args");
  spatial_softmax_f32_bwd_kernel<<<dim3(N, (C + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(x, points, dpoints, dx, H, W, C);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_tile_add_context_f32_fwd(const float* x, const float* ctx, float* y, int32_t Nx, int32_t Nc, int32_t HW,
                                                int32_t C, void* stream) {
  T2R_CHECK_ARG(x && ctx && y && Nx > 0 && Nc >= Nx && Nc % Nx == 0 && HW > 0 && C > 0, "tile_add_context_f32_fwd: bad args");
  tile_add_context_fwd_kernel<<<grid_for((long long)Nc * HW * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ctx, y, Nx, Nc, HW, C);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_tile_add_context_f32_bwd(const float* dy, float* dx, float* dctx, int32_t Nx, int32_t Nc, int32_t HW,
                                                int32_t C, void* stream) {
  T2R_CHECK_ARG(dy && Nx > 0 && Nc >= Nx && Nc % Nx == 0 && HW > 0 && C > 0, "tile_add_context_f32_bwd: bad args");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dctx != nullptr) {
    tile_add_context_dctx_kernel<<<Nc, 256, 0, st>>>(dy, dctx, HW, C);
    T2R_LAUNCH_OK();
  }
  if (dx != nullptr) {
    tile_add_context_dx_kernel<<<grid_for((long long)Nx * HW * C), 256, 0, st>>>(dy, dx, Nx, Nc, HW, C);
    T2R_LAUNCH_OK();
  }
  return T2R_OK;
}

extern "C" int32_t t2r_relu_f32_fwd(const float* x, float* y, int64_t n, void* stream) {
  T2R_CHECK_ARG(x && y && n > 0, "relu_f32_fwd: bad args");
  relu_f32_fwd_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_relu_f32_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
  T2R_CHECK_ARG(dy && y && dx && n > 0, "relu_f32_bwd: bad args");
  relu_f32_bwd_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, y, dx, n);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_elu_f32_fwd(const float* x, float* y, int64_t n, void* stream) {
  T2R_CHECK_ARG(x && y && n > 0, "elu_f32_fwd: bad args");
  elu_f32_fwd_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_elu_f32_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
  T2R_CHECK_ARG(dy && y && dx && n > 0, "elu_f32_bwd: bad args");
  elu_f32_bwd_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, y, dx, n);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_bn_infer_f32_fwd(const float* x, const float* gamma, const float* beta, const float* mean, const float* var,
                                        float* y, int64_t rows, int32_t C, float eps, void* stream) {
  T2R_CHECK_ARG(x && gamma && beta && mean && var && y && rows > 0 && C > 0, "bn_infer_f32_fwd: bad args");
  bn_infer_f32_fwd_kernel<<<grid_for(rows * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, gamma, beta, mean, var, y,
                                                                                           rows * C, C, eps);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_bn_infer_f32_bwd(const float* x, const float* dy, const float* gamma, const float* mean, const float* var,
                                        float* dx, float* dgamma, float* dbeta, int64_t rows, int32_t C, float eps, void* stream) {
  T2R_CHECK_ARG(x && dy && gamma && mean && var && dx && rows > 0 && C > 0 && ((dgamma == nullptr) == (dbeta == nullptr)),
                "bn_infer_f32_bwd: bad args");
  bn_infer_f32_bwd_kernel<<<(C + 31) / 32, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, dy, gamma, mean, var, dx, dgamma, dbeta,
                                                                                      rows, C, eps);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_film_relu_f32_fwd(const float* x, const float* film, float* y, int32_t N, int32_t HW, int32_t C,
                                         void* stream) {
  T2R_CHECK_ARG(x && film && y && N > 0 && HW > 0 && C > 0, "film_relu_f32_fwd: bad args");
  film_relu_f32_fwd_kernel<<<dim3(N, std::min(HW, 64)), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, film, y, HW, C);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_film_relu_f32_bwd(const float* x, const float* film, const float* dy, float* dx, float* dfilm,
                                         int32_t N, int32_t HW, int32_t C, void* stream) {
  T2R_CHECK_ARG(x && film && dy && dx && dfilm && N > 0 && HW > 0 && C > 0, "film_relu_f32_bwd: bad args");
  film_relu_f32_bwd_kernel<<<N, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, film, dy, dx, dfilm, HW, C);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_bn_train_f32_fwd(const float* x, const float* gamma, const float* beta, float* y, float* moving_mean,
                                        float* moving_var, float* save_mean, float* save_rstd, int64_t rows, int32_t C,
                                        float eps, float decay, int32_t relu, void* stream) {
  T2R_CHECK_ARG(x && beta && y && moving_mean && moving_var && save_mean && save_rstd && rows > 0 && C > 0,
                "bn_train_f32_fwd: bad args");
  bn_train_f32_fwd_kernel<<<C, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, gamma, beta, y, moving_mean, moving_var,
                                                                            save_mean, save_rstd, rows, C, eps, decay, relu);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_bn_train_f32_bwd(const float* x, const float* y, const float* dy, const float* gamma,
                                        const float* save_mean, const float* save_rstd, float* dx, float* dgamma,
                                        float* dbeta, int64_t rows, int32_t C, int32_t relu, void* stream) {
  T2R_CHECK_ARG(x && y && dy && save_mean && save_rstd && dx && rows > 0 && C > 0, "bn_train_f32_bwd: bad args");
  bn_train_f32_bwd_kernel<<<C, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, dy, gamma, save_mean, save_rstd, dx, dgamma,
                                                                            dbeta, rows, C, relu);
  T2R_LAUNCH_OK();
  return T2R_OK;
}
