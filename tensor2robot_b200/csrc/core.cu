// core.cu — library plumbing (errors, launch accounting, TMA descriptor encoding) and the small
// utility kernels: weight packing, small-Cin im2col, casts, adds, fp32 SGEMM for tiny layers.
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "conv_common.cuh"

namespace t2r {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// cuTensorMapEncodeTiled is fetched through the runtime so that the library loads (and its
// symbol table can be checked) on machines without libcuda.so.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return -1;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr,
                  bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box "
              "[%u,%u,%u,%u] base %p",
              int(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
    return -1;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------

// w fp32 [Cout][taps][Cin] -> wf bf16 same layout, wd bf16 [Cin][taps][Cout].
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                    __nv_bfloat16* __restrict__ wd, int Cout, int taps, int Cin) {
  const long long total = (long long)Cout * taps * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = w[i];
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    if (wf) wf[i] = b;
    if (wd) {
      const int ci = int(i % Cin);
      const long long r = i / Cin;
      const int t = int(r % taps);
      const int co = int(r / taps);
      wd[((long long)ci * taps + t) * Cout + co] = b;
    }
  }
}

// x bf16 [N,H,W,Cin] -> a bf16 [N*Ho*Wo, Kpad]; one thread per (pixel, 8-column group).
__global__ void im2col_small_cin_kernel(const __nv_bfloat16* __restrict__ x,
                                        __nv_bfloat16* __restrict__ a, int N, int H, int W, int Cin,
                                        int KH, int KW, int stride, int pad_top, int pad_left,
                                        int Ho, int Wo, int Kpad) {
  // Per-k lookup table (shared memory) instead of a div/mod chain per element:
  //   tab_hw[k] = kh << 16 | kw (or -1 for the zero padding columns), tab_d[k] = element offset of
  //   (kh, kw, c) relative to the window origin.
  extern __shared__ int im2col_tab[];
  int* tab_hw = im2col_tab;
  int* tab_d = im2col_tab + Kpad;
  const int Kreal = KH * KW * Cin;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    if (k < Kreal) {
      const int c = k % Cin, t = k / Cin;
      const int kw = t % KW, kh = t / KW;
      tab_hw[k] = (kh << 16) | kw;
      tab_d[k] = (kh * W + kw) * Cin + c;
    } else {
      tab_hw[k] = -1;
      tab_d[k] = 0;
    }
  }
  __syncthreads();
  const int groups = Kpad / 8;
  const long long total = (long long)N * Ho * Wo * groups;
  const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = int(i % groups);
    long long pix = i / groups;
    const int ow = int(pix % Wo);
    long long r = pix / Wo;
    const int oh = int(r % Ho);
    const int n = int(r / Ho);
    const int ih0 = oh * stride - pad_top, iw0 = ow * stride - pad_left;
    const long long base = (((long long)n * H + ih0) * W + iw0) * Cin;
    unsigned short v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      const int hw = tab_hw[k];
      unsigned short val = 0;
      if (hw >= 0) {
        const int ih = ih0 + (hw >> 16), iw = iw0 + (hw & 0xFFFF);
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) val = xs[base + tab_d[k]];
      }
      v[j] = val;
    }
    uint4 q;
    q.x = v[0] | (uint32_t(v[1]) << 16);
    q.y = v[2] | (uint32_t(v[3]) << 16);
    q.z = v[4] | (uint32_t(v[5]) << 16);
    q.w = v[6] | (uint32_t(v[7]) << 16);
    reinterpret_cast<uint4*>(a)[i] = q;
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                     long long n) {
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    reinterpret_cast<uint2*>(y)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = n4 * 4 + threadIdx.x;
    y[i] = __float2bfloat16_rn(x[i]);
  }
}

__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y,
                                     long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}

__global__ void add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                uint4* __restrict__ y, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 p = a[i], q = b[i];
    uint4 o;
    o.x = pack_bf16(bf16_lo(p.x) + bf16_lo(q.x), bf16_hi(p.x) + bf16_hi(q.x));
    o.y = pack_bf16(bf16_lo(p.y) + bf16_lo(q.y), bf16_hi(p.y) + bf16_hi(q.y));
    o.z = pack_bf16(bf16_lo(p.z) + bf16_lo(q.z), bf16_hi(p.z) + bf16_hi(q.z));
    o.w = pack_bf16(bf16_lo(p.w) + bf16_lo(q.w), bf16_hi(p.w) + bf16_hi(q.w));
    y[i] = o;
  }
}

// y = relu(a + b): the tail of a ResNet v1 block (film_resnet_model.py:156-166)
__global__ void add_relu_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                     uint4* __restrict__ y, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 p = a[i], q = b[i];
    uint4 o;
    o.x = pack_bf16(fmaxf(bf16_lo(p.x) + bf16_lo(q.x), 0.f), fmaxf(bf16_hi(p.x) + bf16_hi(q.x), 0.f));
    o.y = pack_bf16(fmaxf(bf16_lo(p.y) + bf16_lo(q.y), 0.f), fmaxf(bf16_hi(p.y) + bf16_hi(q.y), 0.f));
    o.z = pack_bf16(fmaxf(bf16_lo(p.z) + bf16_lo(q.z), 0.f), fmaxf(bf16_hi(p.z) + bf16_hi(q.z), 0.f));
    o.w = pack_bf16(fmaxf(bf16_lo(p.w) + bf16_lo(q.w), 0.f), fmaxf(bf16_hi(p.w) + bf16_hi(q.w), 0.f));
    y[i] = o;
  }
}

__global__ void add_bf16_tail_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b,
                                     __nv_bfloat16* y, long long start, long long n) {
  const long long i = start + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2bfloat16_rn(__bfloat162float(a[i]) + __bfloat162float(b[i]));
}

__global__ void relu_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ dy,
                                     const __nv_bfloat16* __restrict__ y,
                                     __nv_bfloat16* __restrict__ dx, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dx[i] = (__bfloat162float(y[i]) > 0.f) ? dy[i] : __float2bfloat16_rn(0.f);
}

__global__ void bias_add_f32_kernel(float* y, const float* __restrict__ bias, long long total,
                                    int C) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    y[i] += bias[i % C];
}

// out[c] = sum_r x[r,c]; one block per 32 columns, 8 row lanes.
__global__ void colsum_f32_kernel(const float* __restrict__ x, float* __restrict__ out,
                                  long long rows, int C) {
  __shared__ float sm[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < C)
    for (long long r = threadIdx.y; r < rows; r += 8) acc += x[r * C + c];
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
    for (int j = 0; j < 8; ++j) s += sm[j][threadIdx.x];
    out[c] = s;
  }
}

// Plain smem-tiled fp32 GEMM (CUDA cores).  Only used for layers whose inner dimension is not a
// multiple of 64 (action context FC 10->256, logits 64->1): < 0.1 % of the step's FLOPs.
template <bool TA, bool TB>
__global__ void sgemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda,
                             const float* __restrict__ B, int ldb, float beta, float* C, int ldc) {
  __shared__ float As[16][17], Bs[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    const int ka = k0 + tx, kb = k0 + ty;
    As[ty][tx] = (row < M && ka < K) ? (TA ? A[(long long)ka * lda + row] : A[(long long)row * lda + ka]) : 0.f;
    Bs[ty][tx] = (col < N && kb < K) ? (TB ? B[(long long)col * ldb + kb] : B[(long long)kb * ldb + col]) : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += As[ty][k] * Bs[k][tx];
    __syncthreads();
  }
  if (row < M && col < N) {
    float* c = C + (long long)row * ldc + col;
    *c = alpha * acc + (beta == 0.f ? 0.f : beta * *c);
  }
}

static inline int grid_for(long long n, int block = 256) {
  return int(std::min<long long>(std::max<long long>((n + block - 1) / block, 1), 148LL * 32));
}

}  // namespace t2r

using namespace t2r;

extern "C" int32_t t2r_version(void) { return 100; }
extern "C" const char* t2r_last_error(void) { return g_err; }
extern "C" int64_t t2r_launch_count(void) { return g_launch_count.load(); }
extern "C" void t2r_launch_count_reset(void) { g_launch_count.store(0); }

extern "C" int32_t t2r_pack_weights(const float* w, void* w_fprop, void* w_dgrad, int32_t Cout,
                                    int32_t taps, int32_t Cin, void* stream) {
  T2R_CHECK_ARG(w && (w_fprop || w_dgrad), "null pointer");
  const long long total = (long long)Cout * taps * Cin;
  pack_weights_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, static_cast<__nv_bfloat16*>(w_fprop), static_cast<__nv_bfloat16*>(w_dgrad), Cout, taps, Cin);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_im2col_small_cin(const T2RConvDesc* d, const void* x, void* a, int32_t Kpad,
                                        void* stream) {
  T2R_CHECK_ARG(d && d->struct_size == sizeof(T2RConvDesc), "bad T2RConvDesc");
  T2R_CHECK_ARG(Kpad % 64 == 0 && Kpad >= d->KH * d->KW * d->Cin, "Kpad=%d invalid", Kpad);
  const long long total = (long long)d->N * d->Ho * d->Wo * (Kpad / 8);
  im2col_small_cin_kernel<<<grid_for(total), 256, 2 * Kpad * sizeof(int), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(a), d->N, d->H, d->W, d->Cin,
      d->KH, d->KW, d->stride, d->pad_top, d->pad_left, d->Ho, d->Wo, Kpad);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_sgemm(int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K,
                             float alpha, const float* A, int32_t lda, const float* B, int32_t ldb,
                             float beta, float* C, int32_t ldc, void* stream) {
  T2R_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0, "bad sgemm args");
  dim3 block(16, 16), grid((N + 15) / 16, (M + 15) / 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!transA && !transB) sgemm_kernel<false, false><<<grid, block, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (!transA && transB) sgemm_kernel<false, true><<<grid, block, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (transA && !transB) sgemm_kernel<true, false><<<grid, block, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else sgemm_kernel<true, true><<<grid, block, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_bias_add_f32(float* y, const float* bias, int64_t rows, int32_t C,
                                    void* stream) {
  T2R_CHECK_ARG(y && bias && rows > 0 && C > 0, "bad bias_add args");
  bias_add_f32_kernel<<<grid_for(rows * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, bias, rows * C, C);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_colsum_f32(const float* x, float* out, int64_t rows, int32_t C, void* stream) {
  T2R_CHECK_ARG(x && out && rows > 0 && C > 0, "bad colsum args");
  colsum_f32_kernel<<<(C + 31) / 32, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(x, out, rows, C);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
  T2R_CHECK_ARG(x && y && n > 0, "bad cast args");
  cast_f32_bf16_kernel<<<grid_for(n / 4 + 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(y), n);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

namespace t2r {
// w' = bf16(w * scale[row]) for a [rows, K] fp32 matrix (K % 4 == 0): inference-mode batch norm folded
// into the weights of the convolution that feeds it.
__global__ void fold_scale_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                  __nv_bfloat16* __restrict__ out, long long n4, int K4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float s = scale[i / K4];
    const float4 v = reinterpret_cast<const float4*>(w)[i];
    uint2 o;
    o.x = pack_bf16(v.x * s, v.y * s);
    o.y = pack_bf16(v.z * s, v.w * s);
    reinterpret_cast<uint2*>(out)[i] = o;
  }
}
}  // namespace t2r

extern "C" int32_t t2r_fold_bn_weights(const float* w, const float* scale, void* w_bf16, int32_t Cout, int64_t K,
                                       void* stream) {
  T2R_CHECK_ARG(w && scale && w_bf16 && Cout > 0 && K > 0 && K % 4 == 0, "fold_bn_weights: bad args");
  const long long n4 = (long long)Cout * K / 4;
  t2r::fold_scale_kernel<<<t2r::grid_for(n4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, scale, static_cast<__nv_bfloat16*>(w_bf16), n4, int(K / 4));
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_cast_bf16_to_f32(const void* x, float* y, int64_t n, void* stream) {
  T2R_CHECK_ARG(x && y && n > 0, "bad cast args");
  cast_bf16_f32_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), y, n);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_add_bf16(const void* a, const void* b, void* y, int64_t n, void* stream) {
  T2R_CHECK_ARG(a && b && y && n > 0, "bad add args");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n8 = n / 8;
  if (n8 > 0) {
    add_bf16_kernel<<<grid_for(n8), 256, 0, st>>>(static_cast<const uint4*>(a), static_cast<const uint4*>(b),
                                                  static_cast<uint4*>(y), n8);
    T2R_LAUNCH_OK();
  }
  if (n % 8) {
    add_bf16_tail_kernel<<<1, 32, 0, st>>>(static_cast<const __nv_bfloat16*>(a),
                                           static_cast<const __nv_bfloat16*>(b),
                                           static_cast<__nv_bfloat16*>(y), n8 * 8, n);
    T2R_LAUNCH_OK();
  }
  return T2R_OK;
}

extern "C" int32_t t2r_add_relu_bf16(const void* a, const void* b, void* y, int64_t n, void* stream) {
  T2R_CHECK_ARG(a && b && y && n > 0 && n % 8 == 0, "add_relu_bf16: n must be a positive multiple of 8");
  add_relu_bf16_kernel<<<grid_for(n / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<uint4*>(y), n / 8);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_relu_bwd_bf16(const void* dy, const void* y, void* dx, int64_t n, void* stream) {
  T2R_CHECK_ARG(dy && y && dx && n > 0, "bad relu_bwd args");
  relu_bwd_bf16_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(y),
      static_cast<__nv_bfloat16*>(dx), n);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

// sum of squares of a flat fp32 buffer (slim l2_regularizer term: l2 * sum(w^2) / 2), accumulated
// into out[0] (caller zeroes).
namespace t2r {
__global__ void __launch_bounds__(256) sumsq_f32_kernel(const float* __restrict__ x, float* out, long long n,
                                                        float scale) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    acc += x[i] * x[i];
  __shared__ float sm[8];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int j = 0; j < 8; ++j) s += sm[j];
    atomicAdd(out, s * scale);
  }
}
}  // namespace t2r

extern "C" int32_t t2r_sumsq_f32(const float* x, float* out, int64_t n, float scale, void* stream) {
  T2R_CHECK_ARG(x && out && n > 0, "bad sumsq args");
  t2r::sumsq_f32_kernel<<<t2r::grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, n, scale);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

namespace t2r {
// x bf16 [N,H,W,3] -> interior of the zero-padded x4p bf16 [N,Hp,Wp,4] (channel 3 = 0).  The border
// of x4p is never written: the caller zeroes the buffer once and reuses it.
__global__ void __launch_bounds__(256) pad_nhwc3_c4_kernel(const unsigned short* __restrict__ x,
                                                           uint2* __restrict__ x4p, int N, int H, int W, int Hp,
                                                           int Wp, int pad_top, int pad_left) {
  const long long total = (long long)N * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = int(i % W);
    const long long r = i / W;
    const int h = int(r % H);
    const int n = int(r / H);
    const unsigned short* px = x + i * 3;
    uint2 o;
    o.x = uint32_t(px[0]) | (uint32_t(px[1]) << 16);
    o.y = uint32_t(px[2]);
    x4p[((long long)n * Hp + h + pad_top) * Wp + w + pad_left] = o;
  }
}

// Row-pair variant: x bf16 [N,H,W,3] -> x8p bf16 [N,Hp/2,Wp,8], where channels 0..3 of pair row p
// hold padded image row 2p and channels 4..7 row 2p+1 (same column).  8 consecutive columns are then
// 128 contiguous bytes = 2 filter rows x 8 pixels x 4 channels: one K chunk of a stride-2 stem.
__global__ void __launch_bounds__(256) pad_nhwc3_pairs_kernel(const unsigned short* __restrict__ x,
                                                              uint2* __restrict__ x8p, int N, int H, int W, int Hp2,
                                                              int Wp, int pad_top, int pad_left) {
  const long long total = (long long)N * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = int(i % W);
    const long long r = i / W;
    const int h = int(r % H);
    const int n = int(r / H);
    const unsigned short* px = x + i * 3;
    uint2 o;
    o.x = uint32_t(px[0]) | (uint32_t(px[1]) << 16);
    o.y = uint32_t(px[2]);
    const int pr = h + pad_top;
    x8p[(((long long)n * Hp2 + (pr >> 1)) * Wp + w + pad_left) * 2 + (pr & 1)] = o;
  }
}

// Clears the padded slots of a stem weight gradient.  rows == 1: [Cout][KH][16 px][4 ch];
// rows == 2: [Cout][chunks][8 px][2 rows][4 ch].
__global__ void stem_mask_grad_kernel(float* dw, long long total, int KH, int KW, int rows, int chunks) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int slot = int(i & 63);
    const int chunk = int((i >> 6) % chunks);
    const int pxi = rows == 2 ? (slot >> 3) : (slot >> 2);
    const int r = rows == 2 ? ((slot >> 2) & 1) : 0;
    if (pxi >= KW || (slot & 3) == 3 || chunk * rows + r >= KH) dw[i] = 0.f;
  }
}
}  // namespace t2r

extern "C" int32_t t2r_pad_nhwc3_c4(const void* x, void* x4p, int32_t N, int32_t H, int32_t W, int32_t Hp,
                                    int32_t Wp, int32_t pad_top, int32_t pad_left, void* stream) {
  T2R_CHECK_ARG(x && x4p && N > 0 && H + pad_top <= Hp && W + pad_left <= Wp, "pad_nhwc3_c4: bad args");
  const long long total = (long long)N * H * W;
  t2r::pad_nhwc3_c4_kernel<<<t2r::grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const unsigned short*>(x), static_cast<uint2*>(x4p), N, H, W, Hp, Wp, pad_top, pad_left);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_stem_pack_image(const void* x, void* xp, int32_t N, int32_t H, int32_t W, int32_t Hp, int32_t Wp,
                                       int32_t pad_top, int32_t pad_left, int32_t KW, int32_t stride, void* stream) {
  if (t2r::stem_rows_per_chunk(KW, stride) == 1)
    return t2r_pad_nhwc3_c4(x, xp, N, H, W, Hp, Wp, pad_top, pad_left, stream);
  T2R_CHECK_ARG(x && xp && N > 0 && Hp % 2 == 0 && H + pad_top <= Hp && W + pad_left <= Wp, "stem_pack_image: bad args");
  const long long total = (long long)N * H * W;
  t2r::pad_nhwc3_pairs_kernel<<<t2r::grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const unsigned short*>(x), static_cast<uint2*>(xp), N, H, W, Hp / 2, Wp, pad_top, pad_left);
  T2R_LAUNCH_OK();
  return T2R_OK;
}

extern "C" int32_t t2r_stem_k(int32_t KH, int32_t KW, int32_t stride) {
  const int rows = t2r::stem_rows_per_chunk(KW, stride);
  return ((KH + rows - 1) / rows) * 64;
}

extern "C" int32_t t2r_stem_mask_grad(float* dw_stem, int32_t Cout, int32_t KH, int32_t KW, int32_t stride,
                                      void* stream) {
  T2R_CHECK_ARG(dw_stem && Cout > 0 && KH > 0 && KW > 0 && KW <= 16, "stem_mask_grad: bad args");
  const int rows = t2r::stem_rows_per_chunk(KW, stride);
  const int chunks = (KH + rows - 1) / rows;
  const long long total = (long long)Cout * chunks * 64;
  t2r::stem_mask_grad_kernel<<<t2r::grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(dw_stem, total, KH,
                                                                                                  KW, rows, chunks);
  T2R_LAUNCH_OK();
  return T2R_OK;
}
