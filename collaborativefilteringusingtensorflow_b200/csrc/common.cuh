// Shared device helpers for libcf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "cf_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcf_b200 is written for sm_100a (B200) only"
#endif

void cf_set_error(const char* fmt, ...);
int cf_num_sms();

#define CF_CHECK_ARG(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      cf_set_error(__VA_ARGS__);     \
      return -1;                     \
    }                                \
  } while (0)

#define CF_CUDA_OK(expr)                                                          \
  do {                                                                            \
    cudaError_t e_ = (expr);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      cf_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return -2;                                                                  \
    }                                                                             \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter-based, so batch (seed, epoch, position, lane) is a pure function.
// ---------------------------------------------------------------------------------------------
struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0;
  c[1] = lo1;
  c[2] = n2;
  c[3] = lo0;
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                           uint32_t k0, uint32_t k1) {
  uint32_t c[4] = {c0, c1, c2, c3};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c[0], c[1], c[2], c[3]};
}

// uniform integer in [0, n) from 64 random bits (multiply-shift; bias < n / 2^64)
__host__ __device__ __forceinline__ uint64_t rand_below(uint32_t a, uint32_t b, uint64_t n) {
  const uint64_t r = ((uint64_t)a << 32) | b;
#ifdef __CUDA_ARCH__
  return __umul64hi(r, n);
#else
  return (uint64_t)(((unsigned __int128)r * n) >> 64);
#endif
}

// ---------------------------------------------------------------------------------------------
// Keyed bijection on [0, n): balanced Feistel network over 2*hb bits + cycle walking.
// ---------------------------------------------------------------------------------------------
struct FeistelKey {
  uint32_t rk[6];
  int hb;  // half width in bits, 2*hb >= ceil(log2 n)
};

__host__ __device__ __forceinline__ FeistelKey feistel_key(uint64_t n, uint64_t seed, uint64_t tweak_a, uint32_t tweak_b) {
  FeistelKey k;
  int bits = 2;
  while (bits < 64 && ((uint64_t)1 << bits) < n) ++bits;
  k.hb = (bits + 1) / 2;
  const Philox4 a = philox4x32_10((uint32_t)tweak_a, (uint32_t)(tweak_a >> 32), tweak_b, 0xFE157E1u,
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
  const Philox4 b = philox4x32_10((uint32_t)tweak_a, (uint32_t)(tweak_a >> 32), tweak_b, 0xFE157E2u,
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
  k.rk[0] = a.x; k.rk[1] = a.y; k.rk[2] = a.z; k.rk[3] = a.w; k.rk[4] = b.x; k.rk[5] = b.y;
  return k;
}

__host__ __device__ __forceinline__ uint64_t feistel_perm(uint64_t x, uint64_t n, const FeistelKey& k) {
  const uint32_t mask = k.hb >= 32 ? 0xFFFFFFFFu : ((1u << k.hb) - 1u);
  do {
    uint32_t L = (uint32_t)(x >> k.hb) & mask, R = (uint32_t)x & mask;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      uint32_t f = R * 0xCC9E2D51u + k.rk[r];
      f ^= f >> 15;
      f *= 0x85EBCA6Bu;
      f ^= f >> 13;
      f *= 0xC2B2AE35u;
      f ^= f >> 16;
      const uint32_t nl = R;
      R = (L ^ f) & mask;
      L = nl;
    }
    x = ((uint64_t)L << k.hb) | R;
  } while (x >= n);
  return x;
}

// ---------------------------------------------------------------------------------------------
// L2-coherent vector access (rows are re-written inside the same kernel by other warps -> no .nc, no L1)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stcg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }

// is `x` in the sorted range idx[lo, hi) ?
__device__ __forceinline__ bool csr_contains(const int32_t* __restrict__ idx, int64_t lo, int64_t hi, int32_t x) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int32_t v = __ldg(idx + mid);
    if (v == x) return true;
    if (v < x) lo = mid + 1; else hi = mid;
  }
  return false;
}
