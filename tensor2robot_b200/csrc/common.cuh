// common.cuh — shared host/device helpers for libt2r_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/t2r_b200.h"

// ------------------------------------------------------------------------------------------
// host: error handling + launch accounting
// ------------------------------------------------------------------------------------------
namespace t2r {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define T2R_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::t2r::set_error(__VA_ARGS__);             \
      return T2R_ERR_INVALID_ARG;                \
    }                                            \
  } while (0)

#define T2R_CUDA_OK(expr)                                                          \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      ::t2r::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                       __FILE__, __LINE__);                                        \
      return T2R_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

// Call after every kernel launch: surfaces launch-configuration errors without syncing.
#define T2R_LAUNCH_OK()                                                            \
  do {                                                                             \
    ::t2r::count_launch();                                                         \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess) {                                                       \
      ::t2r::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                       __FILE__, __LINE__);                                        \
      return T2R_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

inline int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Encode a bf16 tensor map with 128B swizzle.  rank 2..4; dims/strides innermost first;
// strides[i] is the byte stride of dim i+1.  Returns 0 on success.
int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box);

}  // namespace t2r

// ------------------------------------------------------------------------------------------
// device: PTX wrappers (mbarrier, TMA, tcgen05).  sm_100a.
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace t2r {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  bool timing = false;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (!timing) {
      t0 = clock64();
      timing = true;
    } else if (clock64() - t0 > 8000000000LL) {  // ~4 s at 2 GHz
      printf("t2r: mbarrier wait timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// TMA tile loads (global -> shared, completes on an mbarrier).
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-converged variants: the WHOLE warp executes the call with warp-uniform operands and one elected
// lane issues.  Keeping the issue loop converged lets the compiler hold descriptors in uniform registers;
// a stream of N = 64 MMAs issues every 48 cycles this way (the shared-memory operand rate) instead of
// 62-100 from inside an `if (lane == 0)` branch (tests/native/exp_mma_issue.cu).
__device__ __forceinline__ void umma_bf16_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp reads lane (base_lane + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA instruction descriptor, kind::f16: bf16 x bf16 -> fp32 (cute::UMMA::InstrDescriptor).
//   [4,6) c_format=1 (F32)  [7,10) a_format=1 (BF16)  [10,13) b_format=1 (BF16)
//   [15] a_major  [16] b_major (0 = K-major, 1 = MN-major)  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major,
                                                       int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn_major) << 15) |
         (uint32_t(b_mn_major) << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// UMMA shared-memory matrix descriptor, 128B swizzle (cute::UMMA::SmemDescriptor):
//   [0,14) addr>>4  [16,30) LBO>>4  [32,46) SBO>>4  [46,48) version=1  [49,52) base_offset
//   [61,64) layout_type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes,
                                                         uint32_t base_offset = 0) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(base_offset & 7) << 49;
  d |= uint64_t(2) << 61;
  return d;
}

__device__ __forceinline__ float bf16_lo(uint32_t packed) {
  return __uint_as_float(packed << 16);
}
__device__ __forceinline__ float bf16_hi(uint32_t packed) {
  return __uint_as_float(packed & 0xFFFF0000u);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// v[j] = value at (row = lane, column j) of a 32 x 32 tile held one row per lane.  Returns, in lane l, the
// sum of column l over the 32 rows: a butterfly that halves the number of live values per step
// (16 + 8 + 4 + 2 + 1 = 31 shuffles), all register indices static.  Destroys v.
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int n = 16; n >= 1; n >>= 1) {
    const bool up = (lane & n) != 0;
#pragma unroll
    for (int j = 0; j < n; ++j) {
      const float send = up ? v[j] : v[j + n];
      const float keep = up ? v[j + n] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, n);
    }
  }
  return v[0];
}

// In-place z = relu(sc * x + sh) on 16-byte pieces of a bf16 tile that TMA wrote with SWIZZLE_128B (row r of 64
// channels at r * 128 B, logical 16-byte chunk c of the row at ((c ^ (r & 7)) * 16)): the batch-norm apply of the
// operand-fused convolutions.  `piece` is the shared address of the caller's first piece, `n` pieces `step` bytes
// apart (the caller picks rows that share r & 7, so the chunk position is the same for all of them); sc / sh are the
// 8 channels of the caller's logical chunk.  Arithmetic and rounding are those of bn_apply_rows_kernel (norm.cu).
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t v) {
  uint32_t r;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(0u));   // rounding is monotonic: relu(rn(z)) == rn(relu(z))
  return r;
}
__device__ __forceinline__ uint4 bnrelu_piece(const uint4 q, const float (&sc)[8], const float (&sh)[8]) {
  uint4 o;
  o.x = relu_bf16x2(pack_bf16(fmaf(bf16_lo(q.x), sc[0], sh[0]), fmaf(bf16_hi(q.x), sc[1], sh[1])));
  o.y = relu_bf16x2(pack_bf16(fmaf(bf16_lo(q.y), sc[2], sh[2]), fmaf(bf16_hi(q.y), sc[3], sh[3])));
  o.z = relu_bf16x2(pack_bf16(fmaf(bf16_lo(q.z), sc[4], sh[4]), fmaf(bf16_hi(q.z), sc[5], sh[5])));
  o.w = relu_bf16x2(pack_bf16(fmaf(bf16_lo(q.w), sc[6], sh[6]), fmaf(bf16_hi(q.w), sc[7], sh[7])));
  return o;
}
// `piece` is the shared address of the caller's first piece, kN pieces `step` bytes apart: all loads are issued
// before the first store so that the eight shared-memory round trips overlap.
template <int kN>
__device__ __forceinline__ void bnrelu_pieces_inplace(uint32_t piece, uint32_t step, const float (&sc)[8],
                                                      const float (&sh)[8]) {
  uint4 q[kN];
#pragma unroll
  for (int i = 0; i < kN; ++i)
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(q[i].x), "=r"(q[i].y), "=r"(q[i].z), "=r"(q[i].w)
                 : "r"(piece + uint32_t(i) * step)
                 : "memory");
#pragma unroll
  for (int i = 0; i < kN; ++i) {
    const uint4 o = bnrelu_piece(q[i], sc, sh);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(piece + uint32_t(i) * step), "r"(o.x), "r"(o.y), "r"(o.z),
                 "r"(o.w)
                 : "memory");
  }
}
__device__ __forceinline__ void load8(const float* __restrict__ p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// generic-proxy writes to shared memory -> visible to the async proxy (TMA stores, tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace t2r
#endif  // __CUDACC__
