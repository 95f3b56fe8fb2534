// test_kernels.cu — standalone GPU sanity harness for libt2r_b200.so (no Python, no torch).
// Every kernel is checked against a naive CPU loop written here (bf16-rounded inputs, double
// accumulation).  Usage: ./test_kernels [filter-substring]; exit code = number of failures.
// The real parity tests (against oracle/) live in tests/test_*_gpu.py; this binary exists so a
// kernel bug is found in seconds of GPU time.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <string>
#include <vector>

#include "../../include/t2r_b200.h"

static int g_fail = 0, g_pass = 0;
static const char* g_filter = nullptr;

#define CK(expr)                                                                   \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      printf("[FATAL] %s -> %s (%s:%d)\n", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      printf("SUMMARY pass=%d fail=%d (aborted)\n", g_pass, g_fail + 1);          \
      exit(100);                                                                   \
    }                                                                              \
  } while (0)
#define T2R(expr)                                                                  \
  do {                                                                             \
    int _rc = (expr);                                                              \
    if (_rc != 0) {                                                                \
      printf("[FAIL] %s rc=%d: %s\n", #expr, _rc, t2r_last_error());               \
      ++g_fail;                                                                    \
      return;                                                                      \
    }                                                                              \
  } while (0)

static uint32_t rng_state = 12345;
static float frand() {  // uniform [-1,1)
  rng_state = rng_state * 1664525u + 1013904223u;
  return float(rng_state >> 8) * (2.0f / 16777216.0f) - 1.0f;
}
static float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <class T>
struct Dev {
  T* p = nullptr;
  size_t n = 0;
  explicit Dev(size_t n_) : n(n_) { CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T))); CK(cudaMemset(p, 0, std::max<size_t>(n,1) * sizeof(T))); }
  ~Dev() { cudaFree(p); }
  void up(const std::vector<T>& h) { CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); }
  std::vector<T> down() const { std::vector<T> h(n); CK(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost)); return h; }
};

static std::vector<__nv_bfloat16> to_bf16(const std::vector<float>& v) {
  std::vector<__nv_bfloat16> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = __float2bfloat16_rn(v[i]);
  return o;
}
static std::vector<float> from_bf16(const std::vector<__nv_bfloat16>& v) {
  std::vector<float> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = __bfloat162float(v[i]);
  return o;
}
static std::vector<float> randv(size_t n, float scale = 1.f, bool round_bf16 = true) {
  std::vector<float> v(n);
  for (auto& x : v) { x = frand() * scale; if (round_bf16) x = bf16r(x); }
  return v;
}

static void report(const std::string& name, const std::vector<float>& got, const std::vector<float>& ref,
                   float rtol, float atol) {
  double maxerr = 0, maxref = 0;
  size_t bad = 0, first_bad = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    const double e = fabs(double(got[i]) - double(ref[i]));
    if (!(e <= atol + rtol * fabs(ref[i]))) { if (!bad) first_bad = i; ++bad; }
    if (e > maxerr || e != e) maxerr = e;
    if (fabs(ref[i]) > maxref) maxref = fabs(ref[i]);
  }
  if (bad) {
    printf("[FAIL] %-44s bad=%zu/%zu maxerr=%.4g maxref=%.4g first_bad=%zu got=%.6g ref=%.6g\n", name.c_str(), bad,
           ref.size(), maxerr, maxref, first_bad, got[first_bad], ref[first_bad]);
    ++g_fail;
  } else {
    printf("[PASS] %-44s n=%zu maxerr=%.4g maxref=%.4g\n", name.c_str(), ref.size(), maxerr, maxref);
    ++g_pass;
  }
  fflush(stdout);
}

static void sync_check(const char* what) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[FATAL] %s: %s\nSUMMARY pass=%d fail=%d (aborted)\n", what, cudaGetErrorString(e), g_pass, g_fail + 1);
    exit(100);
  }
}

// ------------------------------------------------------------------------------------------
// convolution: fprop / dgrad / wgrad against naive loops
// ------------------------------------------------------------------------------------------
struct ConvCase {
  const char* name;
  int N, H, W, Cin, Cout, KH, KW, stride, pt, pl, Ho, Wo, flags;
};

static T2RConvDesc mkdesc(const ConvCase& c) {
  T2RConvDesc d;
  memset(&d, 0, sizeof(d));
  d.struct_size = sizeof(d);
  d.N = c.N; d.H = c.H; d.W = c.W; d.Cin = c.Cin; d.Cout = c.Cout; d.KH = c.KH; d.KW = c.KW;
  d.stride = c.stride; d.pad_top = c.pt; d.pad_left = c.pl; d.Ho = c.Ho; d.Wo = c.Wo; d.flags = c.flags;
  return d;
}

static void test_conv(const ConvCase& c) {
  const T2RConvDesc d = mkdesc(c);
  const size_t nx = size_t(c.N) * c.H * c.W * c.Cin, ny = size_t(c.N) * c.Ho * c.Wo * c.Cout;
  const int taps = c.KH * c.KW;
  const size_t nw = size_t(c.Cout) * taps * c.Cin;
  const float wscale = 1.0f / sqrtf(float(taps * c.Cin));
  std::vector<float> x = randv(nx), w = randv(nw, wscale, false), bias = randv(c.Cout, 1.f, false),
                     res = randv(ny), dy = randv(ny);
  Dev<__nv_bfloat16> dx_(nx), dxg(nx), dres(ny), ddy(ny), dyo(ny), dwf(nw), dwd(nw);
  Dev<float> dw32(nw), dbias(c.Cout), dyo32(ny), dwg(nw);
  dx_.up(to_bf16(x)); dres.up(to_bf16(res)); ddy.up(to_bf16(dy)); dw32.up(w); dbias.up(bias);
  T2R(t2r_pack_weights(dw32.p, dwf.p, dwd.p, c.Cout, taps, c.Cin, nullptr));
  sync_check("pack");
  std::vector<float> wb(nw);
  for (size_t i = 0; i < nw; ++i) wb[i] = bf16r(w[i]);

  // ---- fprop ----
  const bool f32out = c.flags & T2R_EPI_OUT_F32;
  T2R(t2r_conv2d_fprop(&d, dx_.p, dwf.p, dbias.p, dres.p, f32out ? (void*)dyo32.p : (void*)dyo.p, nullptr));
  sync_check((std::string(c.name) + " fprop").c_str());
  std::vector<float> ref(ny);
  for (int n = 0; n < c.N; ++n)
    for (int oh = 0; oh < c.Ho; ++oh)
      for (int ow = 0; ow < c.Wo; ++ow)
        for (int co = 0; co < c.Cout; ++co) {
          double acc = 0;
          for (int kh = 0; kh < c.KH; ++kh) {
            const int ih = oh * c.stride + kh - c.pt;
            if (ih < 0 || ih >= c.H) continue;
            for (int kw = 0; kw < c.KW; ++kw) {
              const int iw = ow * c.stride + kw - c.pl;
              if (iw < 0 || iw >= c.W) continue;
              const float* xp = &x[((size_t(n) * c.H + ih) * c.W + iw) * c.Cin];
              const float* wp = &wb[(size_t(co) * taps + kh * c.KW + kw) * c.Cin];
              for (int ci = 0; ci < c.Cin; ++ci) acc += double(xp[ci]) * wp[ci];
            }
          }
          const size_t o = ((size_t(n) * c.Ho + oh) * c.Wo + ow) * c.Cout + co;
          float v = float(acc);
          if (c.flags & T2R_EPI_BIAS) v += bias[co];
          if (c.flags & T2R_EPI_RESIDUAL) v += res[o];
          if (c.flags & T2R_EPI_RELU) v = fmaxf(v, 0.f);
          ref[o] = v;
        }
  report(std::string(c.name) + " fprop", f32out ? dyo32.down() : from_bf16(dyo.down()), ref, f32out ? 1e-4f : 8e-3f,
         f32out ? 1e-4f : 4e-3f);
  if (!f32out && c.Cout <= 2048) {
    // fused bn_stats: column sums / sums of squares of exactly the bf16 values that were stored
    Dev<double> dstats(2 * c.Cout);
    CK(cudaMemset(dyo.p, 0, ny * 2));
    T2R(t2r_conv2d_fprop_stats(&d, dx_.p, dwf.p, dbias.p, dres.p, dyo.p, dstats.p, nullptr));
    sync_check((std::string(c.name) + " fprop+stats").c_str());
    const std::vector<float> yv = from_bf16(dyo.down());
    const std::vector<double> st = dstats.down();
    std::vector<float> got(2 * c.Cout), want(2 * c.Cout);
    std::vector<double> acc(2 * c.Cout, 0.0);
    for (size_t i = 0; i < ny; ++i) { const int co = int(i % c.Cout); acc[co] += yv[i]; acc[c.Cout + co] += double(yv[i]) * yv[i]; }
    double amax = 0;
    for (int i = 0; i < 2 * c.Cout; ++i) { got[i] = float(st[i]); want[i] = float(acc[i]); amax = fmax(amax, fabs(acc[i])); }
    report(std::string(c.name) + " fused bn_stats", got, want, 1e-4f, float(1e-5 * amax + 1e-4));
  }

  // ---- dgrad ----
  CK(cudaMemset(dxg.p, 0x7f, nx * 2));  // poison
  T2R(t2r_conv2d_dgrad(&d, ddy.p, dwd.p, dxg.p, 0, nullptr));
  sync_check((std::string(c.name) + " dgrad").c_str());
  std::vector<double> dxref(nx, 0.0);
  for (int n = 0; n < c.N; ++n)
    for (int oh = 0; oh < c.Ho; ++oh)
      for (int ow = 0; ow < c.Wo; ++ow)
        for (int kh = 0; kh < c.KH; ++kh) {
          const int ih = oh * c.stride + kh - c.pt;
          if (ih < 0 || ih >= c.H) continue;
          for (int kw = 0; kw < c.KW; ++kw) {
            const int iw = ow * c.stride + kw - c.pl;
            if (iw < 0 || iw >= c.W) continue;
            double* dxp = &dxref[((size_t(n) * c.H + ih) * c.W + iw) * c.Cin];
            const float* dyp = &dy[((size_t(n) * c.Ho + oh) * c.Wo + ow) * c.Cout];
            for (int co = 0; co < c.Cout; ++co) {
              const float* wp = &wb[(size_t(co) * taps + kh * c.KW + kw) * c.Cin];
              const double g = dyp[co];
              for (int ci = 0; ci < c.Cin; ++ci) dxp[ci] += g * wp[ci];
            }
          }
        }
  std::vector<float> dxr(nx);
  for (size_t i = 0; i < nx; ++i) dxr[i] = float(dxref[i]);
  report(std::string(c.name) + " dgrad", from_bf16(dxg.down()), dxr, 8e-3f, 4e-3f);
  // accumulate variant: dx = dx + dgrad  (starting from the previous result => 2x)
  T2R(t2r_conv2d_dgrad(&d, ddy.p, dwd.p, dxg.p, 1, nullptr));
  sync_check((std::string(c.name) + " dgrad-acc").c_str());
  std::vector<float> dxr2(nx);
  {
    std::vector<float> prev(nx);
    for (size_t i = 0; i < nx; ++i) prev[i] = bf16r(dxr[i]);
    for (size_t i = 0; i < nx; ++i) dxr2[i] = prev[i] + dxr[i];
  }
  report(std::string(c.name) + " dgrad-acc", from_bf16(dxg.down()), dxr2, 1.6e-2f, 8e-3f);

  // ---- wgrad ----
  CK(cudaMemset(dwg.p, 0, nw * 4));
  T2R(t2r_conv2d_wgrad(&d, dx_.p, ddy.p, dwg.p, nullptr));
  sync_check((std::string(c.name) + " wgrad").c_str());
  std::vector<double> dwref(nw, 0.0);
  for (int n = 0; n < c.N; ++n)
    for (int oh = 0; oh < c.Ho; ++oh)
      for (int ow = 0; ow < c.Wo; ++ow) {
        const float* dyp = &dy[((size_t(n) * c.Ho + oh) * c.Wo + ow) * c.Cout];
        for (int kh = 0; kh < c.KH; ++kh) {
          const int ih = oh * c.stride + kh - c.pt;
          if (ih < 0 || ih >= c.H) continue;
          for (int kw = 0; kw < c.KW; ++kw) {
            const int iw = ow * c.stride + kw - c.pl;
            if (iw < 0 || iw >= c.W) continue;
            const float* xp = &x[((size_t(n) * c.H + ih) * c.W + iw) * c.Cin];
            for (int co = 0; co < c.Cout; ++co) {
              double* dwp = &dwref[(size_t(co) * taps + kh * c.KW + kw) * c.Cin];
              const double g = dyp[co];
              for (int ci = 0; ci < c.Cin; ++ci) dwp[ci] += g * xp[ci];
            }
          }
        }
      }
  std::vector<float> dwr(nw);
  double wmax = 0;
  for (size_t i = 0; i < nw; ++i) { dwr[i] = float(dwref[i]); wmax = fmax(wmax, fabs(dwref[i])); }
  report(std::string(c.name) + " wgrad", dwg.down(), dwr, 2e-3f, float(2e-4 * wmax + 1e-4));
}

static void tests_conv() {
  const ConvCase cases[] = {
      {"gemm 1x1 W=300 128->64", 1, 1, 300, 128, 64, 1, 1, 1, 0, 0, 1, 300, 0},
      {"3x3 s1 SAME 20x20 64->64", 2, 20, 20, 64, 64, 3, 3, 1, 1, 1, 20, 20, 0},
      {"3x3 s1 SAME +bias+res+relu", 2, 20, 20, 64, 64, 3, 3, 1, 1, 1, 20, 20,
       T2R_EPI_BIAS | T2R_EPI_RESIDUAL | T2R_EPI_RELU},
      {"5x5 s1 SAME 19x23 64->128", 1, 19, 23, 64, 128, 5, 5, 1, 2, 2, 19, 23, 0},
      {"3x3 s2 pad1 21x21 128->256", 2, 21, 21, 128, 256, 3, 3, 2, 1, 1, 11, 11, 0},
      {"1x1 s2 15x15 64->64", 2, 15, 15, 64, 64, 1, 1, 2, 0, 0, 8, 8, 0},
      {"3x3 VALID 14->12 64->64", 3, 14, 14, 64, 64, 3, 3, 1, 0, 0, 12, 12, 0},
      {"1x1 256->512 f32 out", 2, 9, 9, 256, 512, 1, 1, 1, 0, 0, 9, 9, T2R_EPI_OUT_F32},
      {"fc 4096->64 rows=70", 1, 1, 70, 4096, 64, 1, 1, 1, 0, 0, 1, 70, T2R_EPI_BIAS},
      {"3x3 s2 even 16x16 64->128", 1, 16, 16, 64, 128, 3, 3, 2, 1, 1, 8, 8, 0},
      // shared-memory halo kernel (conv_halo.cu): streaming 5x5, two K chunks, odd tile count, big image
      {"halo 5x5 s1 SAME 19x23 64->64", 2, 19, 23, 64, 64, 5, 5, 1, 2, 2, 19, 23, T2R_EPI_BIAS | T2R_EPI_RELU},
      {"halo 3x3 s1 SAME 37x21 128->64", 1, 37, 21, 128, 64, 3, 3, 1, 1, 1, 37, 21, 0},
      {"halo 3x3 s1 SAME 16x8 64->64 odd tiles", 3, 16, 8, 64, 64, 3, 3, 1, 1, 1, 16, 8, T2R_EPI_RESIDUAL},
      {"halo 3x3 s1 SAME 118x118 64->64", 1, 118, 118, 64, 64, 3, 3, 1, 1, 1, 118, 118, 0},
  };
  for (const auto& c : cases) {
    if (g_filter && !strstr(c.name, g_filter) && !strstr("conv", g_filter)) continue;
    test_conv(c);
  }
}


// ------------------------------------------------------------------------------------------
// batch-norm operand fusion: fprop / wgrad / dgrad *_bnrelu against the unfused pipeline
// (t2r_bn_apply -> t2r_conv2d_*) on the device, and the reduction against a CPU loop
// ------------------------------------------------------------------------------------------
static void test_bnfuse(const ConvCase& c, bool with_res, bool accumulate) {
  std::string name = std::string("bnfuse ") + c.name + (with_res ? " +res" : "") + (accumulate ? " +acc" : "");
  if (g_filter && !strstr(name.c_str(), g_filter)) return;
  const T2RConvDesc d = mkdesc(c);
  const size_t nx = size_t(c.N) * c.H * c.W * c.Cin, ny = size_t(c.N) * c.Ho * c.Wo * c.Cout;
  const int taps = c.KH * c.KW;
  const size_t nw = size_t(c.Cout) * taps * c.Cin;
  const float wscale = 1.0f / sqrtf(float(taps * c.Cin));
  std::vector<float> x = randv(nx, 2.f), w = randv(nw, wscale, false), res = randv(ny), dy = randv(ny), prev = randv(nx),
                     scale = randv(c.Cin, 1.f, false), shift = randv(c.Cin, 1.f, false);
  for (auto& v : scale) v = 0.75f + 0.5f * v;
  Dev<__nv_bfloat16> dx_(nx), dz(nx), dres(ny), ddy(ny), y0(ny), y1(ny), dwf(nw), dwd(nw), g0(nx), g1(nx);
  Dev<float> dw32(nw), dsc(c.Cin), dsh(c.Cin), dw0(nw), dw1(nw);
  Dev<double> st0(2 * c.Cout), st1(2 * c.Cout), red(2 * c.Cin);
  dx_.up(to_bf16(x)); dres.up(to_bf16(res)); ddy.up(to_bf16(dy)); dw32.up(w); dsc.up(scale); dsh.up(shift);
  T2R(t2r_pack_weights(dw32.p, dwf.p, dwd.p, c.Cout, taps, c.Cin, nullptr));
  T2R(t2r_bn_apply(dx_.p, dz.p, int64_t(c.N) * c.H * c.W, c.Cin, dsc.p, dsh.p, nullptr, 1, 1, nullptr));
  const bool one = c.KH == 1 && c.KW == 1;
  if (one) {
    T2RConvDesc df = d;
    df.flags = with_res ? T2R_EPI_RESIDUAL : 0;
    T2R(t2r_conv2d_fprop_stats(&df, dz.p, dwf.p, nullptr, with_res ? dres.p : nullptr, y0.p, st0.p, nullptr));
    T2R(t2r_conv2d_fprop_bnrelu(&df, dx_.p, dsc.p, dsh.p, dwf.p, with_res ? dres.p : nullptr, y1.p, st1.p, nullptr));
    T2R(t2r_conv2d_wgrad(&d, dz.p, ddy.p, dw0.p, nullptr));
    T2R(t2r_conv2d_wgrad_bnrelu(&d, dx_.p, dsc.p, dsh.p, ddy.p, dw1.p, nullptr));
  }
  if (accumulate) { g0.up(to_bf16(prev)); g1.up(to_bf16(prev)); }
  T2R(t2r_conv2d_dgrad(&d, ddy.p, dwd.p, g0.p, accumulate ? 1 : 0, nullptr));
  T2R(t2r_conv2d_dgrad_bnrelu(&d, ddy.p, dwd.p, dx_.p, dsc.p, dsh.p, g1.p, accumulate ? 1 : 0, red.p, nullptr));
  sync_check(name.c_str());
  if (one) {
    report(name + " fprop (vs bn_apply+fprop)", from_bf16(y1.down()), from_bf16(y0.down()), 0.f, 0.f);
    std::vector<double> s0 = st0.down(), s1 = st1.down();
    std::vector<float> a(s0.begin(), s0.end()), b(s1.begin(), s1.end());
    report(name + " fprop fused stats", b, a, 1e-5f, 1e-3f);
    double wmax = 0;
    std::vector<float> r0 = dw0.down();
    for (float v : r0) wmax = fmax(wmax, fabs(v));
    report(name + " wgrad (vs bn_apply+wgrad)", dw1.down(), r0, 1e-4f, float(1e-5 * wmax + 1e-6));
  }
  // reference: mask the plain data gradient on the CPU, reduce in double
  std::vector<float> dzr = from_bf16(g0.down()), gref(nx), redref(2 * c.Cin);
  std::vector<double> acc(2 * c.Cin, 0.0);
  for (size_t i = 0; i < nx; ++i) {
    const int ch = int(i % c.Cin);
    const bool on = fmaf(x[i], scale[ch], shift[ch]) > 0.f;
    gref[i] = on ? dzr[i] : 0.f;
    acc[ch] += gref[i];
    acc[c.Cin + ch] += double(gref[i]) * x[i];
  }
  double rmax = 0;
  for (int i = 0; i < 2 * c.Cin; ++i) { redref[i] = float(acc[i]); rmax = fmax(rmax, fabs(acc[i])); }
  std::vector<float> gout = from_bf16(g1.down());
  // fused launches store the masked gradient; fallback launches (halo / accumulate) store the plain one
  bool masked = true;
  for (size_t i = 0; i < nx && masked; ++i) masked = gout[i] == gref[i];
  report(name + (masked ? " dgrad g (masked)" : " dgrad g (plain, reduce pass)"), gout, masked ? gref : dzr, 0.f, 0.f);
  std::vector<double> rd = red.down();
  std::vector<float> rdf(rd.begin(), rd.end());
  report(name + " dgrad sums", rdf, redref, 2e-4f, float(2e-5 * rmax + 1e-4));
}

static void tests_bnfuse() {
  const ConvCase cases[] = {
      {"1x1 19x17 64->256", 2, 19, 17, 64, 256, 1, 1, 1, 0, 0, 19, 17, 0},       // tma<128> x2 / dgrad tma<64>
      {"1x1 19x17 256->64", 2, 19, 17, 256, 64, 1, 1, 1, 0, 0, 19, 17, 0},       // tma<64>, 4 K chunks / dgrad tma<128>
      {"1x1 30x30 512->128", 3, 30, 30, 512, 128, 1, 1, 1, 0, 0, 30, 30, 0},     // dgrad 512 out ch, K = 128
      {"1x1 9x9 1024->512", 2, 9, 9, 1024, 512, 1, 1, 1, 0, 0, 9, 9, 0},         // igemm<256> register epilogue both ways
      {"1x1 s2 15x15 256->512", 2, 15, 15, 256, 512, 1, 1, 2, 0, 0, 8, 8, 0},    // strided projection: phase launches
      {"3x3 s1 20x20 128->128", 2, 20, 20, 128, 128, 3, 3, 1, 1, 1, 20, 20, 0},  // dgrad only (tma<128>)
      {"3x3 s2 21x21 256->256", 2, 21, 21, 256, 256, 3, 3, 2, 1, 1, 11, 11, 0},  // dgrad only, 4 phases, igemm<256>
      {"3x3 s1 37x21 64->64 halo", 1, 37, 21, 64, 64, 3, 3, 1, 1, 1, 37, 21, 0}, // halo kernel
  };
  for (const auto& c : cases) {
    test_bnfuse(c, false, false);
  }
  test_bnfuse(cases[0], true, false);
  test_bnfuse(cases[1], false, true);
  test_bnfuse(cases[3], true, false);
}

// ------------------------------------------------------------------------------------------
// batch norm
// ------------------------------------------------------------------------------------------
static void test_bn(int rows, int C, bool with_gamma, int relu, bool with_res) {
  char name[128];
  snprintf(name, sizeof(name), "bn rows=%d C=%d gamma=%d relu=%d res=%d", rows, C, int(with_gamma), relu, int(with_res));
  if (g_filter && !strstr(name, g_filter)) return;
  const size_t n = size_t(rows) * C;
  std::vector<float> x = randv(n, 2.f), dy = randv(n), dres = randv(n), gamma = randv(C, 1.f, false),
                     beta = randv(C, 1.f, false), mm(C, 0.25f), mv(C, 2.f);
  for (auto& g : gamma) g = 1.f + 0.5f * g;
  for (size_t i = 0; i < n; ++i) x[i] = bf16r(x[i] + 0.5f * float(i % C) / C);
  Dev<__nv_bfloat16> dx(n), dyv(n), dr(n), dy_out(n), dxg(n);
  Dev<double> stats(2 * C), red(2 * C);
  Dev<float> dgamma(C), dbeta(C), dmm(C), dmv(C), dmean(C), dinv(C), dscale(C), dshift(C), dg(C), db(C);
  dx.up(to_bf16(x)); dyv.up(to_bf16(dy)); dr.up(to_bf16(dres)); dgamma.up(gamma); dbeta.up(beta); dmm.up(mm); dmv.up(mv);
  const float eps = 1e-3f, mom = 0.9f;
  T2R(t2r_bn_stats(dx.p, rows, C, stats.p, nullptr));
  T2R(t2r_bn_finalize(stats.p, rows, C, with_gamma ? dgamma.p : nullptr, dbeta.p, eps, mom, dmm.p, dmv.p, dmean.p,
                      dinv.p, dscale.p, dshift.p, nullptr));
  T2R(t2r_bn_apply(dx.p, dy_out.p, rows, C, dscale.p, dshift.p, nullptr, 1, relu, nullptr));
  T2R(t2r_bn_backward(dyv.p, dx.p, with_res ? dr.p : nullptr, dxg.p, rows, C, with_gamma ? dgamma.p : nullptr,
                      dmean.p, dinv.p, dscale.p, dshift.p, relu, red.p, dg.p, db.p, nullptr));
  sync_check(name);
  // reference
  std::vector<float> yref(n), dxref(n), mean(C), var(C), mmr(C), mvr(C), dgr(C), dbr(C);
  for (int c = 0; c < C; ++c) {
    double s = 0, q = 0;
    for (int r = 0; r < rows; ++r) { const double v = x[size_t(r) * C + c]; s += v; q += v * v; }
    const double m = s / rows, va = fmax(q / rows - m * m, 0.0);
    mean[c] = float(m); var[c] = float(va);
    mmr[c] = mm[c] * mom + float(m) * (1 - mom);
    mvr[c] = mv[c] * mom + float(va * rows / (rows - 1)) * (1 - mom);
    const double inv = 1.0 / sqrt(va + eps), g = with_gamma ? gamma[c] : 1.0;
    double sdz = 0, sdzx = 0;
    for (int r = 0; r < rows; ++r) {
      const size_t i = size_t(r) * C + c;
      const double xh = (x[i] - m) * inv, z = g * xh + beta[c];
      yref[i] = float(relu ? fmax(z, 0.0) : z);
      const double dz = (relu && !(z > 0)) ? 0.0 : dy[i];
      sdz += dz; sdzx += dz * xh;
    }
    dbr[c] = float(sdz); dgr[c] = float(sdzx);
    for (int r = 0; r < rows; ++r) {
      const size_t i = size_t(r) * C + c;
      const double xh = (x[i] - m) * inv, z = g * xh + beta[c];
      const double dz = (relu && !(z > 0)) ? 0.0 : dy[i];
      dxref[i] = float(g * inv * (dz - sdz / rows - xh * sdzx / rows) + (with_res ? dres[i] : 0.f));
    }
  }
  report(std::string(name) + " mean", dmean.down(), mean, 1e-4f, 1e-5f);
  report(std::string(name) + " moving_mean", dmm.down(), mmr, 1e-4f, 1e-5f);
  report(std::string(name) + " moving_var", dmv.down(), mvr, 1e-4f, 1e-5f);
  report(std::string(name) + " y", from_bf16(dy_out.down()), yref, 8e-3f, 8e-3f);
  report(std::string(name) + " dbeta", db.down(), dbr, 1e-3f, 2e-2f);
  report(std::string(name) + " dgamma", dg.down(), dgr, 1e-3f, 2e-2f);
  report(std::string(name) + " dx", from_bf16(dxg.down()), dxref, 8e-3f, 8e-3f);
}

// ------------------------------------------------------------------------------------------
// pooling etc.
// ------------------------------------------------------------------------------------------
static void test_maxpool(int N, int H, int W, int C, int k, int s, bool same) {
  char name[128];
  snprintf(name, sizeof(name), "maxpool %dx%d k%d s%d %s C=%d", H, W, k, s, same ? "SAME" : "VALID", C);
  if (g_filter && !strstr(name, g_filter)) return;
  int Ho, Wo, pt = 0, pl = 0;
  if (same) { t2r_conv_same_padding(H, k, s, &Ho, &pt); t2r_conv_same_padding(W, k, s, &Wo, &pl); }
  else { Ho = (H - k) / s + 1; Wo = (W - k) / s + 1; }
  const size_t nx = size_t(N) * H * W * C, ny = size_t(N) * Ho * Wo * C;
  std::vector<float> x = randv(nx), dy = randv(ny);
  // introduce ties
  for (size_t i = 0; i < nx; i += 7) x[i] = 0.5f;
  Dev<__nv_bfloat16> dx(nx), dyo(ny), ddy(ny), dxg(nx);
  Dev<uint8_t> arg(ny);
  dx.up(to_bf16(x)); ddy.up(to_bf16(dy));
  T2R(t2r_maxpool_fwd(dx.p, dyo.p, arg.p, N, H, W, C, k, s, pt, pl, Ho, Wo, nullptr));
  T2R(t2r_maxpool_bwd(ddy.p, arg.p, dxg.p, N, H, W, C, k, s, pt, pl, Ho, Wo, nullptr));
  sync_check(name);
  std::vector<float> yref(ny), dxref(nx, 0.f);
  for (int n = 0; n < N; ++n) for (int oh = 0; oh < Ho; ++oh) for (int ow = 0; ow < Wo; ++ow) for (int c = 0; c < C; ++c) {
    float best = -INFINITY; size_t bi = 0;
    for (int kh = 0; kh < k; ++kh) for (int kw = 0; kw < k; ++kw) {
      const int ih = oh * s + kh - pt, iw = ow * s + kw - pl;
      if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
      const size_t i = ((size_t(n) * H + ih) * W + iw) * C + c;
      if (x[i] > best) { best = x[i]; bi = i; }
    }
    const size_t o = ((size_t(n) * Ho + oh) * Wo + ow) * C + c;
    yref[o] = best; dxref[bi] += dy[o];
  }
  report(std::string(name) + " fwd", from_bf16(dyo.down()), yref, 0.f, 0.f);
  report(std::string(name) + " bwd", from_bf16(dxg.down()), dxref, 8e-3f, 1e-6f);
}

static void test_misc() {
  if (g_filter && !strstr("misc", g_filter)) return;
  {  // global mean fwd/bwd
    const int N = 3, HW = 225, C = 128;
    std::vector<float> x = randv(size_t(N) * HW * C), dy = randv(size_t(N) * C);
    Dev<__nv_bfloat16> dx(x.size()), y(size_t(N) * C), ddy(dy.size()), dxg(x.size());
    dx.up(to_bf16(x)); ddy.up(to_bf16(dy));
    T2R(t2r_global_mean_fwd(dx.p, y.p, N, HW, C, nullptr));
    T2R(t2r_global_mean_bwd(ddy.p, dxg.p, N, HW, C, nullptr));
    sync_check("global_mean");
    std::vector<float> yr(size_t(N) * C), dxr(x.size());
    for (int n = 0; n < N; ++n) for (int c = 0; c < C; ++c) {
      double s = 0; for (int p = 0; p < HW; ++p) s += x[(size_t(n) * HW + p) * C + c];
      yr[size_t(n) * C + c] = float(s / HW);
      for (int p = 0; p < HW; ++p) dxr[(size_t(n) * HW + p) * C + c] = dy[size_t(n) * C + c] / HW;
    }
    report("global_mean fwd", from_bf16(y.down()), yr, 8e-3f, 1e-3f);
    report("global_mean bwd", from_bf16(dxg.down()), dxr, 8e-3f, 1e-5f);
  }
  {  // add context fwd/bwd
    const int B = 3, A = 5, HW = 49, C = 64;
    std::vector<float> x = randv(size_t(B) * HW * C), ctx = randv(size_t(B) * A * C), dy = randv(size_t(B) * A * HW * C);
    Dev<__nv_bfloat16> dx(x.size()), dc(ctx.size()), y(dy.size()), ddy(dy.size()), dxg(x.size()), dcg(ctx.size());
    dx.up(to_bf16(x)); dc.up(to_bf16(ctx)); ddy.up(to_bf16(dy));
    T2R(t2r_add_context_fwd(dx.p, dc.p, y.p, B, A, HW, C, nullptr));
    T2R(t2r_add_context_bwd(ddy.p, dxg.p, dcg.p, B, A, HW, C, nullptr));
    sync_check("add_context");
    std::vector<float> yr(dy.size()), dxr(x.size(), 0.f), dcr(ctx.size(), 0.f);
    for (int b = 0; b < B; ++b) for (int a = 0; a < A; ++a) for (int p = 0; p < HW; ++p) for (int c = 0; c < C; ++c) {
      const size_t o = ((size_t(b * A + a)) * HW + p) * C + c;
      yr[o] = x[(size_t(b) * HW + p) * C + c] + ctx[size_t(b * A + a) * C + c];
      dxr[(size_t(b) * HW + p) * C + c] += dy[o];
      dcr[size_t(b * A + a) * C + c] += dy[o];
    }
    report("add_context fwd", from_bf16(y.down()), yr, 8e-3f, 1e-3f);
    report("add_context bwd dx", from_bf16(dxg.down()), dxr, 8e-3f, 8e-3f);
    report("add_context bwd dctx", from_bf16(dcg.down()), dcr, 8e-3f, 2e-2f);
  }
  {  // sgemm NN/NT/TN + bias + colsum
    const int M = 37, N = 29, K = 10;
    std::vector<float> A = randv(size_t(M) * K, 1.f, false), Bm = randv(size_t(N) * K, 1.f, false), Cm(size_t(M) * N, 0.f);
    Dev<float> dA(A.size()), dB(Bm.size()), dC(Cm.size());
    dA.up(A); dB.up(Bm);
    T2R(t2r_sgemm(0, 1, M, N, K, 1.f, dA.p, K, dB.p, K, 0.f, dC.p, N, nullptr));  // C = A * B^T
    sync_check("sgemm");
    std::vector<float> ref(Cm.size());
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += double(A[m * K + k]) * Bm[n * K + k]; ref[m * N + n] = float(s); }
    report("sgemm NT", dC.down(), ref, 1e-5f, 1e-5f);
    // TN: D[K,N'] = A^T[K,M] * C[M,N]
    Dev<float> dD(size_t(K) * N);
    T2R(t2r_sgemm(1, 0, K, N, M, 1.f, dA.p, K, dC.p, N, 0.f, dD.p, N, nullptr));
    sync_check("sgemm TN");
    std::vector<float> ref2(size_t(K) * N);
    for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) { double s = 0; for (int m = 0; m < M; ++m) s += double(A[m * K + k]) * ref[m * N + n]; ref2[k * N + n] = float(s); }
    report("sgemm TN", dD.down(), ref2, 1e-4f, 1e-4f);
    Dev<float> dcs(N);
    T2R(t2r_colsum_f32(dC.p, dcs.p, M, N, nullptr));
    sync_check("colsum");
    std::vector<float> cs(N);
    for (int n = 0; n < N; ++n) { double s = 0; for (int m = 0; m < M; ++m) s += ref[m * N + n]; cs[n] = float(s); }
    report("colsum", dcs.down(), cs, 1e-5f, 1e-5f);
  }
  {  // sigmoid log loss
    const int n = 1000;
    std::vector<float> z = randv(n, 4.f, false), y(n);
    for (int i = 0; i < n; ++i) y[i] = (i % 3 == 0) ? 1.f : 0.f;
    Dev<float> dz(n), dyl(n), q(n), loss(1), dl(n);
    dz.up(z); dyl.up(y);
    T2R(t2r_sigmoid_logloss(dz.p, dyl.p, q.p, loss.p, dl.p, n, nullptr));
    sync_check("logloss");
    double L = 0; std::vector<float> qr(n), dr(n);
    for (int i = 0; i < n; ++i) {
      const double qq = 1.0 / (1.0 + exp(-double(z[i]))), e = 1e-7;
      L += -(y[i] * log(qq + e) + (1 - y[i]) * log(1 - qq + e));
      qr[i] = float(qq);
      dr[i] = float(qq * (1 - qq) * (-(y[i] / (qq + e)) + (1 - y[i]) / (1 - qq + e)) / n);
    }
    report("logloss q", q.down(), qr, 1e-5f, 1e-6f);
    report("logloss loss", loss.down(), std::vector<float>{float(L / n)}, 1e-5f, 1e-6f);
    report("logloss dlogit", dl.down(), dr, 1e-4f, 1e-8f);
  }
  {  // optimizers
    const int n = 1004, nd = 600;   // 16-byte vectors: n and nd are multiples of 4
    std::vector<float> w = randv(n, 1.f, false), g = randv(n, 1.f, false), a = randv(n, 0.1f, false), e = w, v(n);
    for (int i = 0; i < n; ++i) v[i] = fabsf(a[i]);
    Dev<float> dw(n), dg(n), da(n), de(n), dm(n), dv(n), dw2(n), de2(n);
    Dev<__nv_bfloat16> wb(n);
    dw.up(w); dg.up(g); da.up(a); de.up(e); dw2.up(w); de2.up(e); dm.up(a); dv.up(v);
    T2R(t2r_momentum_step(dw.p, dg.p, da.p, de.p, wb.p, n, nd, 0.01f, 0.9f, 7e-5f, 0.5f, 0.999f, nullptr));
    T2R(t2r_adam_step(dw2.p, dg.p, dm.p, dv.p, de2.p, nullptr, n, nd, 1e-3f, 0.9f, 0.999f, 1e-8f, 3, 7e-5f, 0.5f, 0.999f, nullptr));
    std::vector<float> ms0(n), mom0 = randv(n, 0.05f, false);
    for (int i = 0; i < n; ++i) ms0[i] = 0.5f + fabsf(a[i]);
    Dev<float> dw3(n), dms(n), dmom(n);
    dw3.up(w); dms.up(ms0); dmom.up(mom0);
    T2R(t2r_rmsprop_step(dw3.p, dg.p, dms.p, dmom.p, nullptr, nullptr, n, nd, 0.01f, 0.9f, 0.8f, 1.0f, 7e-5f, 0.5f, 0.f, nullptr));
    T2R(t2r_momentum_step(dw.p + 1, dg.p, da.p, nullptr, nullptr, 4, 0, 0.01f, 0.9f, 0.f, 1.f, 0.f, nullptr) == T2R_ERR_INVALID_ARG ? 0 : 1);
    sync_check("optim");
    std::vector<float> wr(n), er(n), wr2(n), wr3(n);
    for (int i = 0; i < n; ++i) {
      const float gi = g[i] * 0.5f + (i < nd ? 7e-5f * w[i] : 0.f);
      const float msi = 0.9f * ms0[i] + 0.1f * gi * gi;
      const float momi = 0.8f * mom0[i] + 0.01f * gi / sqrtf(msi + 1.0f);
      wr3[i] = w[i] - momi;
    }
    const double lrt = 1e-3 * sqrt(1 - pow(0.999, 3)) / (1 - pow(0.9, 3));
    for (int i = 0; i < n; ++i) {
      float gi = g[i] * 0.5f + (i < nd ? 7e-5f * w[i] : 0.f);
      const float acc = 0.9f * a[i] + gi;
      wr[i] = w[i] - 0.01f * acc;
      er[i] = 0.999f * e[i] + 0.001f * wr[i];
      const float mi = 0.9f * a[i] + 0.1f * gi, vi = 0.999f * v[i] + 0.001f * gi * gi;
      wr2[i] = w[i] - float(lrt) * mi / (sqrtf(vi) + 1e-8f);
    }
    report("momentum w", dw.down(), wr, 1e-6f, 1e-7f);
    report("momentum ema", de.down(), er, 1e-6f, 1e-7f);
    report("momentum w_bf16", from_bf16(wb.down()), wr, 8e-3f, 1e-6f);
    report("adam w", dw2.down(), wr2, 1e-5f, 1e-6f);
    report("rmsprop w", dw3.down(), wr3, 1e-5f, 1e-6f);
  }
  {  // CEM refit + bellman
    const int B = 4, A = 64, D = 10, E = 10;
    std::vector<float> s = randv(size_t(B) * A * D, 1.f, false), val = randv(size_t(B) * A, 1.f, false);
    val[5] = val[9];  // tie
    Dev<float> ds(s.size()), dv(val.size()), dm(B * D), dsd(B * D), bv(B);
    Dev<int> bi(B);
    ds.up(s); dv.up(val);
    T2R(t2r_cem_refit(ds.p, dv.p, dm.p, dsd.p, bv.p, bi.p, B, A, D, E, nullptr));
    sync_check("cem_refit");
    std::vector<float> mr(B * D), sr(B * D), bvr(B);
    for (int b = 0; b < B; ++b) {
      std::vector<int> order(A);
      for (int a = 0; a < A; ++a) order[a] = a;
      std::stable_sort(order.begin(), order.end(), [&](int i, int j) { return val[b * A + i] < val[b * A + j]; });
      for (int d = 0; d < D; ++d) {
        double m = 0; for (int e2 = A - E; e2 < A; ++e2) m += s[(size_t(b) * A + order[e2]) * D + d];
        m /= E;
        double ss = 0; for (int e2 = A - E; e2 < A; ++e2) { const double t = s[(size_t(b) * A + order[e2]) * D + d] - m; ss += t * t; }
        mr[b * D + d] = float(m); sr[b * D + d] = float(sqrt(ss / (E - 1)));
      }
      float best = val[b * A]; for (int a = 1; a < A; ++a) best = fmaxf(best, val[b * A + a]);
      bvr[b] = best;
    }
    report("cem_refit mean", dm.down(), mr, 1e-5f, 1e-6f);
    report("cem_refit std", dsd.down(), sr, 1e-5f, 1e-6f);
    report("cem_refit best", bv.down(), bvr, 0.f, 0.f);
  }
  {  // crop + convert (no distortion) bit-exact, and f32 distort path sanity
    const int N = 2, H = 40, W = 50, h = 31, w = 33;
    std::vector<uint8_t> img(size_t(N) * H * W * 3);
    for (auto& p : img) { rng_state = rng_state * 1664525u + 1013904223u; p = uint8_t(rng_state >> 24); }
    std::vector<T2RDistortParams> pr(N);
    for (int n = 0; n < N; ++n) { memset(&pr[n], 0, sizeof(pr[n])); pr[n].saturation_scale = 1.f; pr[n].contrast_scale = 1.f; pr[n].crop_y = 3 + n; pr[n].crop_x = 7 - n; }
    Dev<uint8_t> dimg(img.size());
    Dev<T2RDistortParams> dpr(N);
    Dev<float> out(size_t(N) * h * w * 3), cm(N * 3);
    dimg.up(img); dpr.up(pr);
    T2R(t2r_crop_convert_distort(dimg.p, out.p, dpr.p, cm.p, N, H, W, h, w, 1, 0, 1, 0, nullptr));
    sync_check("crop_convert");
    std::vector<float> ref(out.n);
    for (int n = 0; n < N; ++n) for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) for (int c = 0; c < 3; ++c)
      ref[((size_t(n) * h + y) * w + x) * 3 + c] = float(img[((size_t(n) * H + y + pr[n].crop_y) * W + x + pr[n].crop_x) * 3 + c]) * (1.0f / 255.0f);
    report("crop_convert f32 exact", out.down(), ref, 0.f, 0.f);
  }
}

// `test_kernels bench N H W Cin Cout K stride [reps] [kinds]`: time fprop / dgrad / wgrad of one
// SAME-padded convolution at full size with CUDA events (no CPU reference).  kinds = any of "fdw".
#define TB(expr) do { int _rc = (expr); if (_rc != 0) { printf("[FAIL] %s rc=%d: %s\n", #expr, _rc, t2r_last_error()); return 1; } } while (0)
static int bench_conv(int argc, char** argv) {
  if (argc < 9) { fprintf(stderr, "usage: bench N H W Cin Cout K stride [reps] [fdw]\n"); return 2; }
  const int N = atoi(argv[2]), H = atoi(argv[3]), W = atoi(argv[4]), Cin = atoi(argv[5]), Cout = atoi(argv[6]),
            K = atoi(argv[7]), stride = atoi(argv[8]);
  const int reps = argc > 9 ? atoi(argv[9]) : 5;
  const char* kinds = argc > 10 ? argv[10] : "fdw";
  const int flags = argc > 11 ? atoi(argv[11]) : 0;  // T2R_EPI_* for fprop; 2 = residual (also dgrad accumulate)
  int32_t Ho, Wo, pt, pl;
  TB(t2r_conv_same_padding(H, K, stride, &Ho, &pt));
  TB(t2r_conv_same_padding(W, K, stride, &Wo, &pl));
  ConvCase c{"bench", N, H, W, Cin, Cout, K, K, stride, pt, pl, Ho, Wo, flags};
  const T2RConvDesc d = mkdesc(c);
  const size_t nx = size_t(N) * H * W * Cin, ny = size_t(N) * Ho * Wo * Cout, nw = size_t(Cout) * K * K * Cin;
  Dev<__nv_bfloat16> x(nx), dx(nx), y(ny), dy(ny), wf(nw), wd(nw), res(flags & T2R_EPI_RESIDUAL ? ny : 1);
  Dev<float> w32(nw), dw(nw), bias(Cout);
  CK(cudaMemset(x.p, 0x3c, nx * 2)); CK(cudaMemset(dy.p, 0x3c, ny * 2));
  w32.up(randv(nw, 0.05f, false));
  TB(t2r_pack_weights(w32.p, wf.p, wd.p, Cout, K * K, Cin, nullptr));
  Dev<char> flush(256u << 20);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double flops = 2.0 * N * Ho * Wo * double(Cout) * K * K * Cin;
  for (const char* k = kinds; *k; ++k) {
    float best = 1e30f, sum = 0;
    for (int r = 0; r < reps + 1; ++r) {
      CK(cudaMemsetAsync(flush.p, r, flush.n));  // evict L2
      CK(cudaEventRecord(e0));
      if (*k == 'f') TB(t2r_conv2d_fprop(&d, x.p, wf.p, bias.p, res.p, y.p, nullptr));
      if (*k == 'd') TB(t2r_conv2d_dgrad(&d, dy.p, wd.p, dx.p, (flags & T2R_EPI_RESIDUAL) ? 1 : 0, nullptr));
      if (*k == 'w') TB(t2r_conv2d_wgrad(&d, x.p, dy.p, dw.p, nullptr));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (r == 0) continue;  // warm-up
      best = fminf(best, ms); sum += ms;
    }
    printf("%c %dx%dx%dx%d->%d k%d s%d  avg %.3f ms  best %.3f ms  %.1f TFLOP/s (avg)\n", *k, N, H, W, Cin, Cout, K,
           stride, sum / reps, best, flops / (sum / reps) * 1e-9);
  }
  return 0;
}

// `test_kernels benchbn rows C [reps]`: GB/s of the batch-norm passes (algorithmic bytes / time).
static int bench_bn(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: benchbn rows C [reps]\n"); return 2; }
  const long long rows = atoll(argv[2]);
  const int C = atoi(argv[3]);
  const int reps = argc > 4 ? atoi(argv[4]) : 5;
  const size_t n = size_t(rows) * C;
  Dev<__nv_bfloat16> x(n), y(n), dy(n), dres(n), dx(n);
  CK(cudaMemset(x.p, 0x3c, n * 2)); CK(cudaMemset(dy.p, 0x3c, n * 2)); CK(cudaMemset(dres.p, 0x3c, n * 2));
  Dev<double> stats(2 * C), red(2 * C);
  Dev<float> gamma(C), beta(C), mm(C), mv(C), mean(C), inv(C), scale(C), shift(C), dg(C), db(C);
  Dev<char> flush(256u << 20);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const char* names[] = {"stats (2 B/el)", "apply+relu (4 B/el)", "backward (10 B/el)", "backward+dres (12 B/el)"};
  const double bytes_per_el[] = {2, 4, 10, 12};
  for (int k = 0; k < 4; ++k) {
    float sum = 0;
    for (int r = 0; r < reps + 1; ++r) {
      CK(cudaMemsetAsync(flush.p, r, flush.n));
      CK(cudaEventRecord(e0));
      if (k == 0) TB(t2r_bn_stats(x.p, rows, C, stats.p, nullptr));
      if (k == 1) TB(t2r_bn_apply(x.p, y.p, rows, C, scale.p, shift.p, nullptr, 1, 1, nullptr));
      if (k >= 2) TB(t2r_bn_backward(dy.p, x.p, k == 3 ? dres.p : nullptr, dx.p, rows, C, gamma.p, mean.p, inv.p, scale.p,
                                     shift.p, 1, red.p, dg.p, db.p, nullptr));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (k == 0 && r == 0)
        TB(t2r_bn_finalize(stats.p, rows, C, nullptr, nullptr, 1e-3f, 0.9f, mm.p, mv.p, mean.p, inv.p, scale.p, shift.p,
                           nullptr));
      if (r > 0) sum += ms;
    }
    printf("bn rows=%lld C=%d %-24s %.3f ms  %.0f GB/s\n", rows, C, names[k], sum / reps,
           bytes_per_el[k] * n / (sum / reps) * 1e-6);
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "bench")) return bench_conv(argc, argv);
  if (argc > 1 && !strcmp(argv[1], "benchbn")) return bench_bn(argc, argv);
  if (argc > 1) g_filter = argv[1];
  int dev_count = 0;
  CK(cudaGetDeviceCount(&dev_count));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs, lib version %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount,
         t2r_version());
  tests_conv();
  tests_bnfuse();
  test_bn(1000, 64, true, 1, false);
  test_bn(777, 256, false, 1, true);
  test_bn(64, 2048, true, 0, false);
  test_bn(5000, 192, true, 1, true);
  test_maxpool(2, 23, 23, 64, 3, 3, true);
  test_maxpool(2, 22, 21, 64, 3, 2, true);
  test_maxpool(2, 23, 22, 64, 3, 2, true);
  test_maxpool(1, 17, 19, 128, 3, 2, false);
  test_maxpool(1, 27, 27, 64, 2, 2, true);
  test_maxpool(1, 12, 12, 128, 2, 2, false);
  test_misc();
  printf("SUMMARY pass=%d fail=%d launches=%lld\n", g_pass, g_fail, (long long)t2r_launch_count());
  return g_fail;
}
