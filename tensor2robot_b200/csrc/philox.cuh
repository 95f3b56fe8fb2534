// philox.cuh — Philox4x32-10 counter-based RNG (Salmon et al., SC'11) + Box-Muller.
// The oracle (oracle/philox.py) restates the same generator so samples can be compared
// element-wise; stream = (seed -> key, 128-bit counter = [index lo, index hi, offset lo, offset hi]).
#pragma once
#include <stdint.h>

namespace t2r {

struct Philox4 {
  uint32_t v[4];
};

__host__ __device__ inline Philox4 philox4x32_10(uint64_t seed, uint64_t index, uint64_t offset) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
  uint32_t c0 = uint32_t(index), c1 = uint32_t(index >> 32), c2 = uint32_t(offset),
           c3 = uint32_t(offset >> 32);
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = uint64_t(M0) * c0, p1 = uint64_t(M1) * c2;
    const uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = uint32_t(p1);
    const uint32_t n2 = uint32_t(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = uint32_t(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

// uniform in (0, 1]: (x + 1) * 2^-32 computed in fp32 via the top 24 bits.
__host__ __device__ inline float u01(uint32_t x) { return (float(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }

#ifdef __CUDACC__
// Two standard normals from two uint32 (Box-Muller).
__device__ inline void box_muller(uint32_t a, uint32_t b, float* z0, float* z1) {
  const float u1 = u01(a), u2 = u01(b);
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincosf(6.28318530717958647692f * u2, &s, &c);
  *z0 = r * c;
  *z1 = r * s;
}
#endif

}  // namespace t2r
