// philox.h — Philox4x32-10 and the 53-bit uniform, shared by device kernels and host code.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define LITE_HD __host__ __device__
#else
#define LITE_HD
#endif

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter = global candidate id
// ---------------------------------------------------------------------------
LITE_HD inline void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2,
                                             uint32_t &c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
  uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

LITE_HD inline void philox4x32_10(uint64_t ctr, uint32_t stream, uint64_t seed,
                                              uint32_t out[4]) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = stream, c3 = 0;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

LITE_HD inline double u53(uint32_t hi, uint32_t lo) {
  return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

