// pool.cu — max pooling, global spatial mean and the action-context broadcast add (bf16 NHWC).
// All HBM-bound: 16-byte vectors over the channel dimension, grid-stride loops.
#include <algorithm>

#include "common.cuh"

namespace t2r {

__device__ __forceinline__ void unpack8(const uint4 q, float (&f)[8]) {
  f[0] = bf16_lo(q.x); f[1] = bf16_hi(q.x); f[2] = bf16_lo(q.y); f[3] = bf16_hi(q.y);
  f[4] = bf16_lo(q.z); f[5] = bf16_hi(q.z); f[6] = bf16_lo(q.w); f[7] = bf16_hi(q.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 q;
  q.x = pack_bf16(f[0], f[1]); q.y = pack_bf16(f[2], f[3]);
  q.z = pack_bf16(f[4], f[5]); q.w = pack_bf16(f[6], f[7]);
  return q;
}

// slim.max_pool2d: padding never wins (acts as -inf); first maximum in row-major window order
// takes the gradient (TF MaxPoolGrad tie rule).
// One block per output row (n, oh): all index arithmetic is 32-bit and per-row invariants are
// hoisted (the 64-bit div/mod chain of a flat grid-stride loop cost 3x the memory time).
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                          uint2* __restrict__ argmax, int N, int H, int W,
                                                          int cg, int k, int stride, int pt, int pl,
                                                          int Ho, int Wo) {
  const int n = blockIdx.x / Ho, oh = blockIdx.x - n * Ho;
  const int kh_lo = max(0, pt - oh * stride), kh_hi = min(k, H + pt - oh * stride);
  const uint4* xn = x + (long long)n * H * W * cg;
  const long long orow = ((long long)n * Ho + oh) * Wo * cg;
  for (int t = threadIdx.x; t < Wo * cg; t += blockDim.x) {
    const int ow = t / cg, g = t - ow * cg;
    const int kw_lo = max(0, pl - ow * stride), kw_hi = min(k, W + pl - ow * stride);
    float best[8];
    unsigned idx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; idx[j] = 0; }
    for (int kh = kh_lo; kh < kh_hi; ++kh) {
      const uint4* xr = xn + ((oh * stride + kh - pt) * W + (ow * stride - pl)) * cg + g;
      for (int kw = kw_lo; kw < kw_hi; ++kw) {
        float f[8];
        unpack8(xr[kw * cg], f);
        const unsigned code = unsigned(kh * k + kw);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (f[j] > best[j]) { best[j] = f[j]; idx[j] = code; }
      }
    }
    y[orow + t] = pack8(best);
    if (argmax) {
      uint2 a;
      a.x = idx[0] | (idx[1] << 8) | (idx[2] << 16) | (idx[3] << 24);
      a.y = idx[4] | (idx[5] << 8) | (idx[6] << 16) | (idx[7] << 24);
      argmax[orow + t] = a;
    }
  }
}

// Gather form: every input pixel sums the dy of the (few) windows that selected it.  One block per
// input row (n, ih); the candidate output rows are the same for the whole block.
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const uint4* __restrict__ dy,
                                                          const uint2* __restrict__ argmax,
                                                          uint4* __restrict__ dx, int N, int H, int W,
                                                          int cg, int k, int stride, int pt, int pl,
                                                          int Ho, int Wo) {
  const int n = blockIdx.x / H, ih = blockIdx.x - n * H;
  // windows oh with 0 <= ih + pt - oh*stride < k
  const int oh_hi = min((ih + pt) / stride, Ho - 1);
  const int oh_lo = (ih + pt - k + 1 > 0) ? (ih + pt - k + stride) / stride : 0;
  const long long obase = (long long)n * Ho * Wo * cg;
  const long long irow = ((long long)n * H + ih) * W * cg;
  for (int t = threadIdx.x; t < W * cg; t += blockDim.x) {
    const int iw = t / cg, g = t - iw * cg;
    const int ow_hi = min((iw + pl) / stride, Wo - 1);
    const int ow_lo = (iw + pl - k + 1 > 0) ? (iw + pl - k + stride) / stride : 0;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      const int kh = ih + pt - oh * stride;
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        const int kw = iw + pl - ow * stride;
        const long long o = obase + (oh * Wo + ow) * cg + g;
        const uint2 a = argmax[o];
        float f[8];
        unpack8(dy[o], f);
        const unsigned code = unsigned(kh * k + kw);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const unsigned sel = ((j < 4 ? a.x : a.y) >> (8 * (j & 3))) & 0xFFu;
          if (sel == code) acc[j] += f[j];
        }
      }
    }
    dx[irow + t] = pack8(acc);
  }
}


// 3x3 / stride-2 pooling (the ResNet stem pool, film_resnet_model.py:567-575; 3.65 GB of activations at
// batch 512), specialised so that every byte moves through L2 once.
//
// forward: a thread produces TWO horizontally adjacent outputs from one 3 x 5 patch (15 loads instead of 18, all
// issued before the first compare).  The (kh, kw) scan order and the strict `>` keep the generic kernel's
// first-maximum tie rule.
__global__ void __launch_bounds__(256) maxpool_fwd_k3s2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                               uint2* __restrict__ argmax, int H, int W, int cg, int pt,
                                                               int pl, int Ho, int Wo) {
  const int n = blockIdx.x / Ho, oh = blockIdx.x - n * Ho;
  const uint4* xn = x + (long long)n * H * W * cg;
  const long long orow = ((long long)n * Ho + oh) * Wo * cg;
  const int pairs = (Wo + 1) >> 1;
  const uint4 ninf = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);   // bf16 -inf
  for (int t = threadIdx.x; t < pairs * cg; t += blockDim.x) {
    const int op = t / cg, g = t - op * cg;
    const int ow = 2 * op;
    uint4 q[3][5];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh * 2 + kh - pt;
#pragma unroll
      for (int kc = 0; kc < 5; ++kc) {
        const int iw = ow * 2 + kc - pl;
        q[kh][kc] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? xn[(ih * W + iw) * cg + g] : ninf;
      }
    }
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      if (ow + o >= Wo) break;
      float best[8];
      unsigned idx[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; idx[j] = 0; }
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          float f[8];
          unpack8(q[kh][2 * o + kw], f);
          const unsigned code = unsigned(kh * 3 + kw);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (f[j] > best[j]) { best[j] = f[j]; idx[j] = code; }
        }
      const long long oi = orow + (long long)(ow + o) * cg + g;
      y[oi] = pack8(best);
      if (argmax) {
        uint2 a;
        a.x = idx[0] | (idx[1] << 8) | (idx[2] << 16) | (idx[3] << 24);
        a.y = idx[4] | (idx[5] << 8) | (idx[6] << 16) | (idx[7] << 24);
        argmax[oi] = a;
      }
    }
  }
}

// backward: in padded coordinates u = ih + pad_top the rows (2a, 2a+1) belong to window a (kh = 0 / 1) and row 2a
// also to window a - 1 (kh = 2); the same for columns.  A thread therefore owns the 2 x 2 input quad (a, b) and
// reads its (at most) four candidate windows ONCE - 96 B for 64 B written instead of 216 B in the gather form.
// Windows are visited in the generic kernel's order (oh, then ow, ascending): sums are bit-identical.
__global__ void __launch_bounds__(256) maxpool_bwd_k3s2_kernel(const uint4* __restrict__ dy,
                                                               const uint2* __restrict__ argmax,
                                                               uint4* __restrict__ dx, int H, int W, int cg, int pt, int pl,
                                                               int Ho, int Wo, int Hq, int Wq) {
  const int n = blockIdx.x / Hq, a = blockIdx.x - n * Hq;
  const long long obase = (long long)n * Ho * Wo * cg;
  const long long ibase = (long long)n * H * W * cg;
  for (int t = threadIdx.x; t < Wq * cg; t += blockDim.x) {
    const int b = t / cg, g = t - b * cg;
    float acc[2][2][8];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[r][c][j] = 0.f;
#pragma unroll
    for (int da = -1; da <= 0; ++da) {
      const int oh = a + da;
      if (oh < 0 || oh >= Ho) continue;
#pragma unroll
      for (int db = -1; db <= 0; ++db) {
        const int ow = b + db;
        if (ow < 0 || ow >= Wo) continue;
        const long long o = obase + (long long)(oh * Wo + ow) * cg + g;
        const uint2 am = argmax[o];
        float f[8];
        unpack8(dy[o], f);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          if (da < 0 && r == 1) continue;           // row 2a+1 is not in window a-1
          const int kh = da < 0 ? 2 : r;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (db < 0 && c == 1) continue;
            const unsigned code = unsigned(kh * 3 + (db < 0 ? 2 : c));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const unsigned sel = ((j < 4 ? am.x : am.y) >> (8 * (j & 3))) & 0xFFu;
              if (sel == code) acc[r][c][j] += f[j];
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int ih = 2 * a + r - pt;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int iw = 2 * b + c - pl;
        if (iw < 0 || iw >= W) continue;
        dx[ibase + (long long)(ih * W + iw) * cg + g] = pack8(acc[r][c]);
      }
    }
  }
}

// x [N, HW, C] -> y [N, C]: one block per (image, 32 column groups), 8 row lanes.
__global__ void __launch_bounds__(256) global_mean_fwd_kernel(const uint4* __restrict__ x,
                                                              uint4* __restrict__ y, int HW, int cg) {
  __shared__ float sm[8][32][9];
  const int n = blockIdx.x;
  const int g = blockIdx.y * 32 + threadIdx.x;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (g < cg)
    for (int p = threadIdx.y; p < HW; p += 8) {
      float f[8];
      unpack8(x[((long long)n * HW + p) * cg + g], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[threadIdx.y][threadIdx.x][j] = acc[j];
  __syncthreads();
  if (threadIdx.y == 0 && g < cg) {
    float o[8];
    const float inv = 1.f / float(HW);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
      for (int r = 0; r < 8; ++r) s += sm[r][threadIdx.x][j];
      o[j] = s * inv;
    }
    y[(long long)n * cg + g] = pack8(o);
  }
}

__global__ void __launch_bounds__(256) global_mean_bwd_kernel(const uint4* __restrict__ dy,
                                                              uint4* __restrict__ dx, long long total8,
                                                              int HW, int cg) {
  const float inv = 1.f / float(HW);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = int(i % cg);
    const long long n = (i / cg) / HW;
    float f[8];
    unpack8(dy[n * cg + g], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= inv;
    dx[i] = pack8(f);
  }
}

// y[(b*A+a), p, :] = x[b, p, :] + ctx[(b*A+a), :]
__global__ void __launch_bounds__(256) add_context_fwd_kernel(const uint4* __restrict__ x,
                                                              const uint4* __restrict__ ctx,
                                                              uint4* __restrict__ y, long long total8,
                                                              int A, int HW, int cg) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = int(i % cg);
    long long r = i / cg;
    const int p = int(r % HW);
    const long long ba = r / HW;
    const long long b = ba / A;
    float fx[8], fc[8];
    unpack8(x[(b * HW + p) * cg + g], fx);
    unpack8(ctx[ba * cg + g], fc);
#pragma unroll
    for (int j = 0; j < 8; ++j) fx[j] += fc[j];
    y[i] = pack8(fx);
  }
}

// Inference fusion of the merge with the batch norm (+ReLU) that follows it:
// y[(b*A+a), p, :] = relu?((x[b, p, :] + ctx[(b*A+a), :]) * scale + shift)
__global__ void __launch_bounds__(256) add_context_affine_kernel(const uint4* __restrict__ x,
                                                                 const uint4* __restrict__ ctx,
                                                                 const float* __restrict__ scale,
                                                                 const float* __restrict__ shift,
                                                                 uint4* __restrict__ y, long long total8, int A,
                                                                 int HW, int cg, int relu) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = int(i % cg);
    long long r = i / cg;
    const int p = int(r % HW);
    const long long ba = r / HW;
    const long long b = ba / A;
    float fx[8], fc[8];
    unpack8(x[(b * HW + p) * cg + g], fx);
    unpack8(ctx[ba * cg + g], fc);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * g);
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * g + 1);
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift) + 2 * g);
    const float4 h1 = __ldg(reinterpret_cast<const float4*>(shift) + 2 * g + 1);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // the unfused path stores x + ctx in bf16 before normalising: keep that rounding for parity
      const float sum = __bfloat162float(__float2bfloat16_rn(fx[j] + fc[j]));
      fx[j] = fmaf(sum, sc[j], sh[j]);
      if (relu) fx[j] = fmaxf(fx[j], 0.f);
    }
    y[i] = pack8(fx);
  }
}

// dx[b,p,:] = sum_a dy[(b*A+a),p,:]
__global__ void __launch_bounds__(256) add_context_bwd_x_kernel(const uint4* __restrict__ dy,
                                                                uint4* __restrict__ dx, long long total8,
                                                                int A, int HW, int cg) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = int(i % cg);
    long long r = i / cg;
    const int p = int(r % HW);
    const long long b = r / HW;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int a = 0; a < A; ++a) {
      float f[8];
      unpack8(dy[((b * A + a) * HW + p) * cg + g], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    dx[i] = pack8(acc);
  }
}

// dctx[(b*A+a),:] = sum_p dy[(b*A+a),p,:]   (same shape of work as global mean without the 1/HW)
__global__ void __launch_bounds__(256) add_context_bwd_ctx_kernel(const uint4* __restrict__ dy,
                                                                  uint4* __restrict__ dctx, int HW, int cg) {
  __shared__ float sm[8][32][9];
  const long long n = blockIdx.x;
  const int g = blockIdx.y * 32 + threadIdx.x;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (g < cg)
    for (int p = threadIdx.y; p < HW; p += 8) {
      float f[8];
      unpack8(dy[(n * HW + p) * cg + g], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[threadIdx.y][threadIdx.x][j] = acc[j];
  __syncthreads();
  if (threadIdx.y == 0 && g < cg) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
      for (int r = 0; r < 8; ++r) s += sm[r][threadIdx.x][j];
      o[j] = s;
    }
    dctx[n * cg + g] = pack8(o);
  }
}

static inline int grid_for(long long n) {
  return int(std::min<long long>(std::max<long long>((n + 255) / 256, 1), 148LL * 16));
}

// ---------------------------------------------------------------------------------------------
// Spatial softmax (layers/spatial_softmax.py:29-88): per (image, channel) softmax over the H*W
// positions and the expected coordinates x_j = 2j/(W-1) - 1, y_i = 2i/(H-1) - 1.  The reference
// reshapes concat([x, y], 1) of shape [B*C, 2] to [B, 2C], i.e. the points come out INTERLEAVED
// (x_1, y_1, x_2, y_2, ...) - the code, not its docstring, is restated here.
// One block per (image, group of 8 channels); 256 threads stride over the positions.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_reduce8(float (&v)[8], float (*sm)[9], bool is_max) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[tid][j] = v[j];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sm[tid][j] = is_max ? fmaxf(sm[tid][j], sm[tid + s][j]) : sm[tid][j] + sm[tid + s][j];
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = sm[0][j];
  __syncthreads();
}

__global__ void __launch_bounds__(256) spatial_softmax_fwd_kernel(const uint4* __restrict__ x, float* __restrict__ points,
                                                                  uint4* __restrict__ softmax, int H, int W, int cg) {
  __shared__ float sm[256][9];
  const int n = blockIdx.x, g = blockIdx.y, HW = H * W, C = cg * 8;
  const uint4* xn = x + (long long)n * HW * cg + g;
  float mx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) mx[j] = -INFINITY;
  for (int p = threadIdx.x; p < HW; p += 256) {
    float f[8];
    unpack8(xn[(long long)p * cg], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) mx[j] = fmaxf(mx[j], f[j]);
  }
  block_reduce8(mx, sm, true);
  float se[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sx[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sy[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const float ax = W > 1 ? 2.f / float(W - 1) : 0.f, ay = H > 1 ? 2.f / float(H - 1) : 0.f;
  for (int p = threadIdx.x; p < HW; p += 256) {
    const int i = p / W, j0 = p - i * W;
    const float px = W > 1 ? ax * float(j0) - 1.f : NAN, py = H > 1 ? ay * float(i) - 1.f : NAN;  // 0/0 in the reference
    float f[8];
    unpack8(xn[(long long)p * cg], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float e = __expf(f[j] - mx[j]);
      se[j] += e;
      sx[j] = fmaf(e, px, sx[j]);
      sy[j] = fmaf(e, py, sy[j]);
    }
  }
  block_reduce8(se, sm, false);
  block_reduce8(sx, sm, false);
  block_reduce8(sy, sm, false);
  if (threadIdx.x < 8) {
    const int c = g * 8 + threadIdx.x;
    points[(long long)n * 2 * C + 2 * c] = sx[threadIdx.x] / se[threadIdx.x];
    points[(long long)n * 2 * C + 2 * c + 1] = sy[threadIdx.x] / se[threadIdx.x];
  }
  if (softmax != nullptr) {
    uint4* sn = softmax + (long long)n * HW * cg + g;
    for (int p = threadIdx.x; p < HW; p += 256) {
      float f[8];
      unpack8(xn[(long long)p * cg], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __expf(f[j] - mx[j]) / se[j];
      sn[(long long)p * cg] = pack8(f);
    }
  }
}

// d logit_p = s_p * ((x_p - E[x]) * dEx + (y_p - E[y]) * dEy), recomputing the softmax from x.
__global__ void __launch_bounds__(256) spatial_softmax_bwd_kernel(const uint4* __restrict__ x, const float* __restrict__ points,
                                                                  const float* __restrict__ dpoints, uint4* __restrict__ dx,
                                                                  int H, int W, int cg) {
  __shared__ float sm[256][9];
  const int n = blockIdx.x, g = blockIdx.y, HW = H * W, C = cg * 8;
  const uint4* xn = x + (long long)n * HW * cg + g;
  float mx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) mx[j] = -INFINITY;
  for (int p = threadIdx.x; p < HW; p += 256) {
    float f[8];
    unpack8(xn[(long long)p * cg], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) mx[j] = fmaxf(mx[j], f[j]);
  }
  block_reduce8(mx, sm, true);
  float se[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = threadIdx.x; p < HW; p += 256) {
    float f[8];
    unpack8(xn[(long long)p * cg], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) se[j] += __expf(f[j] - mx[j]);
  }
  block_reduce8(se, sm, false);
  float ex[8], ey[8], gx[8], gy[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long o = (long long)n * 2 * C + 2 * (g * 8 + j);
    ex[j] = points[o]; ey[j] = points[o + 1];
    gx[j] = dpoints[o]; gy[j] = dpoints[o + 1];
  }
  const float ax = 2.f / float(W - 1), ay = 2.f / float(H - 1);
  uint4* dn = dx + (long long)n * HW * cg + g;
  for (int p = threadIdx.x; p < HW; p += 256) {
    const int i = p / W, j0 = p - i * W;
    const float px = ax * float(j0) - 1.f, py = ay * float(i) - 1.f;
    float f[8];
    unpack8(xn[(long long)p * cg], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sft = __expf(f[j] - mx[j]) / se[j];
      f[j] = sft * ((px - ex[j]) * gx[j] + (py - ey[j]) * gy[j]);
    }
    dn[(long long)p * cg] = pack8(f);
  }
}

}  // namespace t2r

using namespace t2r;

extern "C" int32_t t2r_maxpool_fwd(const void* x, void* y, uint8_t* argmax, int32_t N, int32_t H,
                                   int32_t W, int32_t C, int32_t k, int32_t stride, int32_t pad_top,
                                   int32_t pad_left, int32_t Ho, int32_t Wo, void* stream) {
  T2R_CHECK_ARG(x && y && C % 8 == 0 && k >= 1 && k * k <= 255 && stride >= 1, "maxpool_fwd: bad args");
  T2R_CHECK_ARG((long long)N * Ho < (1LL << 31) && (long long)H * W * (C / 8) < (1LL << 31), "maxpool_fwd: too large");
  if (k == 3 && stride == 2) {
    maxpool_fwd_k3s2_kernel<<<N * Ho, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(x), static_cast<uint4*>(y), reinterpret_cast<uint2*>(argmax), H, W, C / 8, pad_top,
        pad_left, Ho, Wo);
    T2R_LAUNCH_OK();
    return T2R_OK;
  }
  maxpool_fwd_kernel<<<N * Ho, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), reinterpret_cast<uint2*>(argmax), N, H, W,
      C / 8, k, stride, pad_top, pad_left, Ho, Wo);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_maxpool_bwd(const void* dy, const uint8_t* argmax, void* dx, int32_t N,
                                   int32_t H, int32_t W, int32_t C, int32_t k, int32_t stride,
                                   int32_t pad_top, int32_t pad_left, int32_t Ho, int32_t Wo,
                                   void* stream) {
  T2R_CHECK_ARG(dy && argmax && dx && C % 8 == 0, "maxpool_bwd: bad args");
  T2R_CHECK_ARG((long long)N * H < (1LL << 31) && (long long)Ho * Wo * (C / 8) < (1LL << 31), "maxpool_bwd: too large");
  if (k == 3 && stride == 2) {
    const int Hq = (H + pad_top + 1) / 2, Wq = (W + pad_left + 1) / 2;   // quads of the padded image
    T2R_CHECK_ARG((long long)N * Hq < (1LL << 31), "maxpool_bwd: too large");
    maxpool_bwd_k3s2_kernel<<<N * Hq, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(dy), reinterpret_cast<const uint2*>(argmax), static_cast<uint4*>(dx), H, W, C / 8,
        pad_top, pad_left, Ho, Wo, Hq, Wq);
    T2R_LAUNCH_OK();
    return T2R_OK;
  }
  maxpool_bwd_kernel<<<N * H, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dy), reinterpret_cast<const uint2*>(argmax), static_cast<uint4*>(dx), N,
      H, W, C / 8, k, stride, pad_top, pad_left, Ho, Wo);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_global_mean_fwd(const void* x, void* y, int32_t N, int32_t HW, int32_t C,
                                       void* stream) {
  T2R_CHECK_ARG(x && y && N > 0 && HW > 0 && C % 8 == 0, "global_mean_fwd: bad args");
  const int cg = C / 8;
  global_mean_fwd_kernel<<<dim3(N, (cg + 31) / 32), dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), HW, cg);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_global_mean_bwd(const void* dy, void* dx, int32_t N, int32_t HW, int32_t C,
                                       void* stream) {
  T2R_CHECK_ARG(dy && dx && N > 0 && HW > 0 && C % 8 == 0, "global_mean_bwd: bad args");
  const long long total8 = (long long)N * HW * (C / 8);
  global_mean_bwd_kernel<<<grid_for(total8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dy), static_cast<uint4*>(dx), total8, HW, C / 8);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_add_context_fwd(const void* x, const void* ctx, void* y, int32_t B, int32_t A,
                                       int32_t HW, int32_t C, void* stream) {
  T2R_CHECK_ARG(x && ctx && y && B > 0 && A > 0 && HW > 0 && C % 8 == 0, "add_context_fwd: bad args");
  const long long total8 = (long long)B * A * HW * (C / 8);
  add_context_fwd_kernel<<<grid_for(total8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<const uint4*>(ctx), static_cast<uint4*>(y), total8, A, HW,
      C / 8);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_add_context_affine_fwd(const void* x, const void* ctx, const float* scale, const float* shift,
                                              void* y, int32_t B, int32_t A, int32_t HW, int32_t C, int32_t relu,
                                              void* stream) {
  T2R_CHECK_ARG(x && ctx && scale && shift && y && B > 0 && A > 0 && HW > 0 && C % 8 == 0,
                "add_context_affine_fwd: bad args");
  const long long total8 = (long long)B * A * HW * (C / 8);
  add_context_affine_kernel<<<grid_for(total8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<const uint4*>(ctx), scale, shift, static_cast<uint4*>(y), total8, A,
      HW, C / 8, relu);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_add_context_bwd(const void* dy, void* dx, void* dctx, int32_t B, int32_t A,
                                       int32_t HW, int32_t C, void* stream) {
  T2R_CHECK_ARG(dy && (dx || dctx) && B > 0 && A > 0 && HW > 0 && C % 8 == 0, "add_context_bwd: bad args");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cg = C / 8;
  if (dx) {
    const long long total8 = (long long)B * HW * cg;
    add_context_bwd_x_kernel<<<grid_for(total8), 256, 0, st>>>(static_cast<const uint4*>(dy),
                                                               static_cast<uint4*>(dx), total8, A, HW, cg);
    T2R_LAUNCH_OK();
  }
  if (dctx) {
    add_context_bwd_ctx_kernel<<<dim3(B * A, (cg + 31) / 32), dim3(32, 8), 0, st>>>(
        static_cast<const uint4*>(dy), static_cast<uint4*>(dctx), HW, cg);
    T2R_LAUNCH_OK();
  }
  return T2R_OK;
}

extern "C" int32_t t2r_spatial_softmax_fwd(const void* x, float* points, void* softmax, int32_t N, int32_t H,
                                           int32_t W, int32_t C, void* stream) {
  T2R_CHECK_ARG(x && points && N > 0 && H > 0 && W > 0 && C % 8 == 0, "spatial_softmax_fwd: bad args");
  spatial_softmax_fwd_kernel<<<dim3(N, C / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), points, static_cast<uint4*>(softmax), H, W, C / 8);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_spatial_softmax_bwd(const void* x, const float* points, const float* dpoints, void* dx,
                                           int32_t N, int32_t H, int32_t W, int32_t C, void* stream) {
  T2R_CHECK_ARG(x && points && dpoints && dx && N > 0 && H > 1 && W > 1 && C % 8 == 0, "spatial_softmax_bwd: bad args");
  spatial_softmax_bwd_kernel<<<dim3(N, C / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), points, dpoints, static_cast<uint4*>(dx), H, W, C / 8);
  T2R_LAUNCH_OK();
  return T2R_OK;
}
