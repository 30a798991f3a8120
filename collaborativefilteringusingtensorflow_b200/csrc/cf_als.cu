// Weighted ALS half-sweep for WRMF (sm_100a): tensor-core Gram G = Y^T Y + batched per-row d x d Cholesky solves.
//
// The reference's wrmf.py is minibatch Adagrad on sampled (u, i, r) rows (reference src/models/basic/models/wrmf.py:52-88;
// that path is cf_train_steps with CF_MODEL_WRMF).  This is the solver of the model the reference's README cites for WRMF
// (README.md:29, Pan et al. / Hu-Koren-Volinsky wALS; the pos/neg-differentiated weighting wrmf.py:61-62 has commented
// out): minimise  sum_{u,i} c_ui (r_ui - x_u.y_i)^2 + reg (|X|^2 + |Y|^2),  c_ui = weight on observed pairs (r = 1), 1
// elsewhere (r = 0).  One half-sweep solves, for every row u of X with observed columns P_u,
//     (Y^T Y + (weight - 1) sum_{i in P_u} y_i y_i^T + reg I) x_u = weight * sum_{i in P_u} y_i          
//   k_als_split    Y fp32 -> transposed bf16 hi / lo planes  Yt[128, n]   (y = hi + lo up to 2^-17: fp32-grade Gram)
//   k_gram_tc      G = Yt Yt^T with tcgen05.mma (hi.hi + hi.lo + lo.hi into one TMEM accumulator), TMA-fed, split over
//                  the item range across CTAs, fp32 atomics into G[128,128]
//   k_als_solve    one CTA per row: A = G + reg I + (weight-1) sum y y^T in shared memory (fp32), in-place Cholesky,
//                  forward/back substitution, x_u written to X.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int ALS_D = 128;   // padded factor dimension handled by the tensor-core Gram (d <= 128)

// ---- Y [n, ld] fp32  ->  hi / lo bf16 planes, transposed: plane[a, i], a < 128 (zero rows for a >= d), i < n_pad
__global__ void __launch_bounds__(256) k_als_split(const float* __restrict__ Y, long long n, long long n_pad, int d, int ld,
                                                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  __shared__ float tile[32][33];
  const long long i0 = (long long)blockIdx.x * 32;
  const int a0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const long long i = i0 + r;
    const int a = a0 + tx;
    tile[r][tx] = (i < n && a < d) ? Y[i * ld + a] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int a = a0 + r;
    const long long i = i0 + tx;
    if (i < n_pad) {
      const float v = tile[tx][r];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[(long long)a * n_pad + i] = h;
      lo[(long long)a * n_pad + i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

// ---- G += Yt[:, chunk range] Yt[:, chunk range]^T on the tensor cores
__global__ void __launch_bounds__(192, 1)
k_gram_tc(const __grid_constant__ CUtensorMap tmHi, const __grid_constant__ CUtensorMap tmLo, int n_chunks, float* __restrict__ G) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sHi = smem;
  uint8_t* sLo = smem + CHUNK_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * CHUNK_BYTES);
  uint64_t* full = bars;
  uint64_t* mma_done = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (n_chunks + gridDim.x - 1) / gridDim.x;
  const int c_lo = blockIdx.x * per, c_hi = min(n_chunks, c_lo + per);

  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(mma_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {   // one thread drives TMA + MMA (the Gram is a tiny share of a half-sweep)
    const uint32_t idesc = umma_idesc_bf16(ROWS, ALS_D);
    uint32_t first = 1u;
    for (int c = c_lo; c < c_hi; ++c) {
      const uint32_t ph = (uint32_t)(c - c_lo) & 1u;
      if (c > c_lo) mbar_wait(mma_done, ph ^ 1u);   // the previous chunk's MMAs have finished reading shared memory
      mbar_arrive_expect_tx(full, 2u * CHUNK_BYTES);
      tma_load_2d(&tmHi, full, sHi, c * KCH, 0);
      tma_load_2d(&tmLo, full, sLo, c * KCH, 0);
      mbar_wait(full, ph);
      tc_fence_after();
      const uint32_t h = smem_u32(sHi), l = smem_u32(sLo);
#pragma unroll
      for (int k = 0; k < KCH / 16; ++k) {
        const uint64_t dh = umma_desc_sw128(h + k * 32), dl = umma_desc_sw128(l + k * 32);
        tc_mma_bf16(tmem_base, dh, dh, idesc, first ? 0u : 1u);   // hi . hi
        first = 0u;
        tc_mma_bf16(tmem_base, dh, dl, idesc, 1u);                // hi . lo
        tc_mma_bf16(tmem_base, dl, dh, idesc, 1u);                // lo . hi
      }
      tc_commit(mma_done);
    }
    if (c_hi > c_lo) mbar_wait(mma_done, (uint32_t)(c_hi - c_lo - 1) & 1u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp >= 2 && c_hi > c_lo) {   // warps 2..5: TMEM lane quadrant = warp % 4; one thread per row of G
    const int q = warp & 3;
    const int row = q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < ALS_D / 32; ++c) {
      uint32_t r[32];
      tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) atomicAdd(G + row * ALS_D + c * 32 + j, __uint_as_float(r[j]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

struct SolveParams {
  float* X;             // [n_x, ldx] rows to solve
  const float* Y;       // [n_y, ldy]
  const float* G;       // [128, 128] = Y^T Y
  const long long* indptr;
  const int32_t* indices;
  long long n_x;
  int d, ldx, ldy;
  float weight, reg;
};

// one CTA per row of X; A (d x d, fp32, leading dimension d + 1) lives in shared memory
__global__ void __launch_bounds__(256) k_als_solve(const __grid_constant__ SolveParams P) {
  extern __shared__ float sm[];
  const int d = P.d, lda = d + 1;
  float* A = sm;                    // [d, lda]
  float* b = A + d * lda;           // [d]
  float* rows = b + d;              // [8, d] gathered y rows
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 thread grid over (row i, column j): no integer divisions in the loops
  for (long long u = blockIdx.x; u < P.n_x; u += gridDim.x) {
    // A = G + reg I (lower triangle), b = 0
    for (int i = ty; i < d; i += 16)
      for (int j = tx; j <= i; j += 16) A[i * lda + j] = P.G[i * ALS_D + j] + (i == j ? P.reg : 0.f);
    for (int e = tid; e < d; e += blockDim.x) b[e] = 0.f;
    const long long lo = P.indptr[u], hi = P.indptr[u + 1];
    const float wm1 = P.weight - 1.f;
    for (long long e0 = lo; e0 < hi; e0 += 8) {
      const int nb = (int)min(8ll, hi - e0);
      __syncthreads();
      for (int r = 0; r < nb; ++r) {
        const float* yr = P.Y + (long long)P.indices[e0 + r] * P.ldy;
        for (int k = tid; k < d; k += blockDim.x) rows[r * d + k] = yr[k];
      }
      __syncthreads();
      if (wm1 != 0.f) {   // A += (weight - 1) sum_r y_r y_r^T   (lower triangle)
        for (int i = ty; i < d; i += 16)
          for (int j = tx; j <= i; j += 16) {
            float acc = 0.f;
            for (int r = 0; r < nb; ++r) acc = fmaf(rows[r * d + i], rows[r * d + j], acc);
            A[i * lda + j] = fmaf(wm1, acc, A[i * lda + j]);
          }
      }
      for (int k = tid; k < d; k += blockDim.x) {   // b += weight * sum_r y_r   (r_ui = 1 on observed pairs)
        float acc = 0.f;
        for (int r = 0; r < nb; ++r) acc += rows[r * d + k];
        b[k] = fmaf(P.weight, acc, b[k]);
      }
    }
    __syncthreads();
    // in-place Cholesky A = L L^T (lower), right-looking
    for (int k = 0; k < d; ++k) {
      if (tid == 0) A[k * lda + k] = sqrtf(fmaxf(A[k * lda + k], 1e-30f));
      __syncthreads();
      const float inv = 1.f / A[k * lda + k];
      for (int i = k + 1 + tid; i < d; i += blockDim.x) A[i * lda + k] *= inv;
      __syncthreads();
      for (int i = k + 1 + ty; i < d; i += 16) {
        const float lik = A[i * lda + k];
        for (int j = k + 1 + tx; j <= i; j += 16) A[i * lda + j] = fmaf(-lik, A[j * lda + k], A[i * lda + j]);
      }
      __syncthreads();
    }
    // L z = b, L^T x = z (one warp; dot products by shuffle reduction)
    if (tid < 32) {
      for (int k = 0; k < d; ++k) {
        float acc = 0.f;
        for (int j = tid; j < k; j += 32) acc = fmaf(A[k * lda + j], b[j], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (tid == 0) b[k] = (b[k] - acc) / A[k * lda + k];
        __syncwarp();
      }
      for (int k = d - 1; k >= 0; --k) {
        float acc = 0.f;
        for (int j = k + 1 + tid; j < d; j += 32) acc = fmaf(A[j * lda + k], b[j], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (tid == 0) b[k] = (b[k] - acc) / A[k * lda + k];
        __syncwarp();
      }
    }
    __syncthreads();
    for (int k = tid; k < d; k += blockDim.x) P.X[u * P.ldx + k] = b[k];
    __syncthreads();
  }
}

// ---- register-blocked variant (the one that is launched): the 128 x 128 system is tiled 16 x 16, thread (ty, tx) owns the
// 8 x 8 block (rows 8ty.., columns 8tx..) of the lower triangle in REGISTERS.  Building A costs 64 FMAs per 16 shared
// loads per gathered row (vs 1 FMA per 2 loads element-wise), and the right-looking blocked Cholesky needs 3 barriers
// per block column (48 per matrix, vs 384 column-wise).  Dimensions d < 128 are padded: the padding block of A is reg * I,
// which decouples and yields zeros.
constexpr int BS = 8, NBK = ALS_D / BS;

__global__ void __launch_bounds__(256, 2) k_als_solve_blocked(const __grid_constant__ SolveParams P) {
  extern __shared__ float sm[];
  constexpr int lda = ALS_D + 1;
  float* L = sm;                          // [128, 129] the factor, written once for the triangular solves
  float* panel = L + ALS_D * lda;         // [16][64] current block column  L_ik (row-major 8 x 8 each)
  float* rows = panel + NBK * 64;         // [8][128] gathered y rows
  float* b = rows + 8 * ALS_D;            // [128]
  float* diag = b + ALS_D;                // [64] factored diagonal block
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const bool lower = ty >= tx;
  const int d = P.d;
  for (long long u = blockIdx.x; u < P.n_x; u += gridDim.x) {
    float a[BS][BS];
    if (lower) {
#pragma unroll
      for (int i = 0; i < BS; ++i)
#pragma unroll
        for (int j = 0; j < BS; ++j) {
          const int gi = BS * ty + i, gj = BS * tx + j;
          a[i][j] = ((gi < d && gj < d) ? __ldg(P.G + gi * ALS_D + gj) : 0.f) + (gi == gj ? P.reg : 0.f);
        }
    }
    if (tid < ALS_D) b[tid] = 0.f;
    const long long lo = P.indptr[u], hi = P.indptr[u + 1];
    const float wm1 = P.weight - 1.f;
    for (long long e0 = lo; e0 < hi; e0 += 8) {
      const int nb = (int)min(8ll, hi - e0);
      __syncthreads();
      for (int r = 0; r < nb; ++r) {
        const float* yr = P.Y + (long long)P.indices[e0 + r] * P.ldy;
        if (tid < ALS_D) rows[r * ALS_D + tid] = tid < d ? yr[tid] : 0.f;
      }
      __syncthreads();
      if (lower && wm1 != 0.f) {
        for (int r = 0; r < nb; ++r) {
          float yi[BS], yj[BS];
#pragma unroll
          for (int i = 0; i < BS; ++i) {
            yi[i] = wm1 * rows[r * ALS_D + BS * ty + i];
            yj[i] = rows[r * ALS_D + BS * tx + i];
          }
#pragma unroll
          for (int i = 0; i < BS; ++i)
#pragma unroll
            for (int j = 0; j < BS; ++j) a[i][j] = fmaf(yi[i], yj[j], a[i][j]);
        }
      }
      if (tid < ALS_D) {
        float acc = 0.f;
        for (int r = 0; r < nb; ++r) acc += rows[r * ALS_D + tid];
        b[tid] = fmaf(P.weight, acc, b[tid]);
      }
    }
    // ---- blocked right-looking Cholesky
    for (int kb = 0; kb < NBK; ++kb) {
      if (ty == kb && tx == kb) {   // factor the diagonal block in registers
#pragma unroll
        for (int k = 0; k < BS; ++k) {
          a[k][k] = sqrtf(fmaxf(a[k][k], 1e-30f));
          const float inv = 1.f / a[k][k];
#pragma unroll
          for (int i = k + 1; i < BS; ++i) a[i][k] *= inv;
#pragma unroll
          for (int i = k + 1; i < BS; ++i)
#pragma unroll
            for (int j = k + 1; j <= i; ++j) a[i][j] = fmaf(-a[i][k], a[j][k], a[i][j]);
        }
#pragma unroll
        for (int i = 0; i < BS; ++i)
#pragma unroll
          for (int j = 0; j < BS; ++j) {
            diag[i * BS + j] = j <= i ? a[i][j] : 0.f;
            L[(BS * kb + i) * lda + BS * kb + j] = j <= i ? a[i][j] : 0.f;
          }
      }
      __syncthreads();
      if (tx == kb && ty > kb) {    // L_ik = A_ik L_kk^-T  (each row of the block: forward substitution with L_kk)
#pragma unroll
        for (int i = 0; i < BS; ++i) {
#pragma unroll
          for (int j = 0; j < BS; ++j) {
            float v = a[i][j];
#pragma unroll
            for (int q = 0; q < j; ++q) v = fmaf(-a[i][q], diag[j * BS + q], v);
            a[i][j] = v / diag[j * BS + j];
          }
        }
#pragma unroll
        for (int i = 0; i < BS; ++i)
#pragma unroll
          for (int j = 0; j < BS; ++j) {
            panel[ty * 64 + i * BS + j] = a[i][j];
            L[(BS * ty + i) * lda + BS * kb + j] = a[i][j];
          }
      }
      __syncthreads();
      if (lower && tx > kb) {       // A_ij -= L_ik L_jk^T
        const float* li = panel + ty * 64;
        const float* lj = panel + tx * 64;
#pragma unroll
        for (int q = 0; q < BS; ++q) {
          float ci[BS], cj[BS];
#pragma unroll
          for (int i = 0; i < BS; ++i) {
            ci[i] = li[i * BS + q];
            cj[i] = lj[i * BS + q];
          }
#pragma unroll
          for (int i = 0; i < BS; ++i)
#pragma unroll
            for (int j = 0; j < BS; ++j) a[i][j] = fmaf(-ci[i], cj[j], a[i][j]);
        }
      }
      __syncthreads();
    }
    // ---- blocked substitutions: L z = b then L^T x = z.  Per block: thread 0 solves the 8 x 8 triangle, then 128
    // threads (one per row) eliminate the solved block from the remaining right-hand side.
    for (int kb = 0; kb < NBK; ++kb) {
      if (tid == 0) {
#pragma unroll
        for (int k = 0; k < BS; ++k) {
          float v = b[BS * kb + k];
#pragma unroll
          for (int q = 0; q < k; ++q) v = fmaf(-L[(BS * kb + k) * lda + BS * kb + q], b[BS * kb + q], v);
          b[BS * kb + k] = v / L[(BS * kb + k) * lda + BS * kb + k];
        }
      }
      __syncthreads();
      if (tid < ALS_D && tid >= BS * (kb + 1)) {
        float v = b[tid];
#pragma unroll
        for (int q = 0; q < BS; ++q) v = fmaf(-L[tid * lda + BS * kb + q], b[BS * kb + q], v);
        b[tid] = v;
      }
      __syncthreads();
    }
    for (int kb = NBK - 1; kb >= 0; --kb) {
      if (tid == 0) {
#pragma unroll
        for (int k = BS - 1; k >= 0; --k) {
          float v = b[BS * kb + k];
#pragma unroll
          for (int q = k + 1; q < BS; ++q) v = fmaf(-L[(BS * kb + q) * lda + BS * kb + k], b[BS * kb + q], v);
          b[BS * kb + k] = v / L[(BS * kb + k) * lda + BS * kb + k];
        }
      }
      __syncthreads();
      if (tid < BS * kb) {       // x_tid -= sum_q L[kb-block row q][tid] * x_q   (L^T)
        float v = b[tid];
#pragma unroll
        for (int q = 0; q < BS; ++q) v = fmaf(-L[(BS * kb + q) * lda + tid], b[BS * kb + q], v);
        b[tid] = v;
      }
      __syncthreads();
    }
    __syncthreads();
    if (tid < d) P.X[u * P.ldx + tid] = b[tid];
    __syncthreads();
  }
}

// ---- the whitened path (weight > 1).  With G0 = Y^T Y + reg I = L L^T and W = Y L^-T (w_a = L^-1 y_a, one dense pass per
// half-sweep) the normal equations of row u with observed set P_u (n = |P_u|, c = weight - 1) become, for x = L^-T xt,
//     (I + c W_u^T W_u) xt = weight * W_u^T 1                                                     (128 x 128, n > 128)
//     xt = W_u^T (weight * 1 - t),   (I / c + W_u W_u^T) t = weight * (W_u W_u^T) 1              (n x n,  n <= 128)
// Both matrices are the identity plus a small positive term (sum_a |w_a|^2 over the WHOLE catalogue is <= 128), so fp32
// Cholesky is comfortable, and a row costs an n x n solve instead of a 128 x 128 one when it is short -- which is most
// rows (configs[3]: median 21 observed columns).  Three kernels share the rows by length:
//   k_als_small<16>, <32>   n <= 16 / 17..32: one WARP per row, the n x n system in registers (one row per lane), shuffles
//   k_als_rows_tc           n > 32: one CTA per row; the Gram (W_u W_u^T for n <= 128, W_u^T W_u in 64-row chunks beyond)
//                           runs on tcgen05 (bf16 hi/lo split, 3 MMAs per k-step, fp32 accumulator in TMEM) from operand
//                           tiles the CTA writes itself in the SWIZZLE_128B K-major layout; then a register-blocked
//                           Cholesky (8 x 8 blocks, two barriers per block column, right-hand side carried as an extra
//                           row so the forward substitution is free) and a block back-substitution out of registers.
// X = Xt L^-1 is one more dense pass (k_als_mul).
constexpr int SM_LD = 132;    // row stride of a gathered fp32 row tile (16-byte aligned rows, conflict-free 128-bit access)

// G0 = G + reg I (identity on the padding) -> L (fp64, in shared memory) -> L^-1 and its transpose in fp32
__global__ void __launch_bounds__(1024) k_als_prep(const float* __restrict__ G, float reg, int d, float* __restrict__ Linv,
                                                    float* __restrict__ LinvT) {
  extern __shared__ double sL[];                       // [128][129]
  constexpr int LD = ALS_D + 1;
  const int tid = threadIdx.x;
  for (int e = tid; e < ALS_D * ALS_D; e += blockDim.x) {
    const int i = e >> 7, j = e & 127;
    sL[i * LD + j] = (i < d && j < d) ? (double)G[e] + (i == j ? (double)reg : 0.0) : (i == j ? 1.0 : 0.0);
  }
  __syncthreads();
  for (int k = 0; k < ALS_D; ++k) {                    // right-looking Cholesky, lower triangle
    if (tid == 0) sL[k * LD + k] = sqrt(fmax(sL[k * LD + k], 1e-300));
    __syncthreads();
    const double inv = 1.0 / sL[k * LD + k];
    for (int i = k + 1 + tid; i < ALS_D; i += blockDim.x) sL[i * LD + k] *= inv;
    __syncthreads();
    const int m = ALS_D - 1 - k;
    for (int e = tid; e < m * m; e += blockDim.x) {
      const int i = k + 1 + e / m, j = k + 1 + e % m;
      if (j <= i) sL[i * LD + j] -= sL[i * LD + k] * sL[j * LD + k];
    }
    __syncthreads();
  }
  // in-place inverse of the lower triangle, last column first: X[j+1:, j] = -X[j+1:, j+1:] L[j+1:, j] / L[j][j]
  for (int j = ALS_D - 1; j >= 0; --j) {
    double x = 0.0;
    const int i = tid;
    if (i > j && i < ALS_D)
      for (int k = j + 1; k <= i; ++k) x += sL[i * LD + k] * sL[k * LD + j];
    __syncthreads();
    const double dj = 1.0 / sL[j * LD + j];
    __syncthreads();
    if (i > j && i < ALS_D) sL[i * LD + j] = -x * dj;
    if (i == j) sL[j * LD + j] = dj;
    __syncthreads();
  }
  for (int e = tid; e < ALS_D * ALS_D; e += blockDim.x) {
    const int i = e >> 7, j = e & 127;
    const float v = j <= i ? (float)sL[i * LD + j] : 0.f;
    Linv[i * ALS_D + j] = v;
    LinvT[j * ALS_D + i] = v;
  }
}

// out[r, :ncols] = in[r, :d] M  (M [128][128] fp32 resident in shared memory; 32 rows per block and pass; in == out is fine:
// a block reads its 32 rows completely before it writes them)
__global__ void __launch_bounds__(256) k_als_mul(const float* in, long long n, int d, int ld_in, const float* __restrict__ M,
                                                 float* out, int ld_out, int ncols) {
  extern __shared__ float sm[];
  float* Ms = sm;                     // [128][128]
  float* Ys = Ms + ALS_D * ALS_D;     // [32][SM_LD]
  for (int e = threadIdx.x; e < ALS_D * ALS_D; e += blockDim.x) Ms[e] = M[e];
  const int col = threadIdx.x & 127, half = threadIdx.x >> 7;     // thread: one column, 16 of the 32 rows
  const int d4 = (d + 3) & ~3;
  for (long long r0 = (long long)blockIdx.x * 32; r0 < n; r0 += (long long)gridDim.x * 32) {
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * ALS_D; e += blockDim.x) {
      const long long r = r0 + (e >> 7);
      const int k = e & 127;
      Ys[(e >> 7) * SM_LD + k] = (r < n && k < d) ? in[r * ld_in + k] : 0.f;
    }
    __syncthreads();
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int k = 0; k < d4; k += 4) {
      const float g0 = Ms[k * ALS_D + col], g1 = Ms[(k + 1) * ALS_D + col], g2 = Ms[(k + 2) * ALS_D + col], g3 = Ms[(k + 3) * ALS_D + col];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 y = *reinterpret_cast<const float4*>(Ys + (half * 16 + i) * SM_LD + k);
        acc[i] = fmaf(y.x, g0, acc[i]);
        acc[i] = fmaf(y.y, g1, acc[i]);
        acc[i] = fmaf(y.z, g2, acc[i]);
        acc[i] = fmaf(y.w, g3, acc[i]);
      }
    }
    __syncthreads();
    if (col < ncols) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const long long r = r0 + half * 16 + i;
        if (r < n) out[r * ld_out + col] = acc[i];
      }
    }
  }
}

struct WhiteParams {
  float* X;                 // [n_x, ldx]: receives xt (the first d columns); k_als_mul turns it into x afterwards
  const float* W;           // [n_y, 128] = Y L^-T
  const long long* indptr;
  const int32_t* indices;
  long long n_x;
  int d, ldx;
  float weight;
  int* cursor;              // k_als_rows_tc: next chunk of rows (zeroed before the launch)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- rows with N_LO < n <= NMAX observed columns (and, in the <16> instance, the empty rows): one warp per row
template <int NMAX, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_als_small(const __grid_constant__ WhiteParams P) {
  constexpr int N_LO = NMAX == 16 ? 0 : 16;
  constexpr int SLD = NMAX + 1;
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* Wt = sm + (size_t)warp * (NMAX * SM_LD + NMAX * SLD);   // [NMAX][SM_LD] gathered rows of W
  float* Ss = Wt + NMAX * SM_LD;                                  // [NMAX][SLD]  W_u W_u^T
  const float c = P.weight - 1.f, invc = 1.f / c;
  const long long nw = (long long)gridDim.x * WARPS;
  for (long long u = (long long)blockIdx.x * WARPS + warp; u < P.n_x; u += nw) {
    const long long lo = P.indptr[u];
    const int n = (int)min(P.indptr[u + 1] - lo, (long long)(NMAX + 1));
    if (n == 0 && N_LO == 0) {                         // nothing observed: the right-hand side is 0, so x = 0
      for (int k = lane; k < P.d; k += 32) P.X[u * P.ldx + k] = 0.f;
      continue;
    }
    if (n <= N_LO || n > NMAX) continue;
    __syncwarp();
    // gather: one 512-byte row per instruction; rows n.. are zero (the padded system decouples)
    const int idx = lane < n ? P.indices[lo + lane] : 0;
#pragma unroll 4
    for (int a = 0; a < NMAX; ++a) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a < n) v = __ldg(reinterpret_cast<const float4*>(P.W + (long long)__shfl_sync(0xffffffffu, idx, a) * ALS_D) + lane);
      *reinterpret_cast<float4*>(Wt + a * SM_LD + 4 * lane) = v;
    }
    __syncwarp();
    // S0 = W_u W_u^T: lane a keeps 64 columns of its own row in registers; the other rows arrive as broadcast 128-bit loads,
    // four at a time (four independent accumulators)
    const int arow = lane < NMAX ? lane : 0;
    const int ng = (n + 3) >> 2;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      float own[64];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(Wt + arow * SM_LD + 64 * pass + 4 * j);
        own[4 * j] = v.x; own[4 * j + 1] = v.y; own[4 * j + 2] = v.z; own[4 * j + 3] = v.w;
      }
#pragma unroll 1
      for (int g = 0; g < ng; ++g) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* wb = Wt + (4 * g) * SM_LD + 64 * pass;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 b0 = *reinterpret_cast<const float4*>(wb + 4 * j);
          const float4 b1 = *reinterpret_cast<const float4*>(wb + SM_LD + 4 * j);
          const float4 b2 = *reinterpret_cast<const float4*>(wb + 2 * SM_LD + 4 * j);
          const float4 b3 = *reinterpret_cast<const float4*>(wb + 3 * SM_LD + 4 * j);
          a0 = fmaf(own[4 * j], b0.x, a0); a0 = fmaf(own[4 * j + 1], b0.y, a0); a0 = fmaf(own[4 * j + 2], b0.z, a0); a0 = fmaf(own[4 * j + 3], b0.w, a0);
          a1 = fmaf(own[4 * j], b1.x, a1); a1 = fmaf(own[4 * j + 1], b1.y, a1); a1 = fmaf(own[4 * j + 2], b1.z, a1); a1 = fmaf(own[4 * j + 3], b1.w, a1);
          a2 = fmaf(own[4 * j], b2.x, a2); a2 = fmaf(own[4 * j + 1], b2.y, a2); a2 = fmaf(own[4 * j + 2], b2.z, a2); a2 = fmaf(own[4 * j + 3], b2.w, a2);
          a3 = fmaf(own[4 * j], b3.x, a3); a3 = fmaf(own[4 * j + 1], b3.y, a3); a3 = fmaf(own[4 * j + 2], b3.z, a3); a3 = fmaf(own[4 * j + 3], b3.w, a3);
        }
        if (lane < NMAX) {
          float* s = Ss + lane * SLD + 4 * g;
          if (pass == 0) { s[0] = a0; s[1] = a1; s[2] = a2; s[3] = a3; }
          else { s[0] += a0; s[1] += a1; s[2] += a2; s[3] += a3; }
        }
      }
    }
    __syncwarp();
    // lane a <- row a of S0 (columns >= 4 ng were never written: they are zero by construction of the padding)
    float S[NMAX];
    float r0 = 0.f;
#pragma unroll
    for (int b = 0; b < NMAX; ++b) {
      S[b] = (lane < n && b < 4 * ng) ? Ss[arow * SLD + b] : 0.f;
      if (b >= n) S[b] = 0.f;
      r0 += S[b];
      if (b == lane) S[b] += invc;
    }
    // Cholesky of I / c + S0, one row per lane; column k is broadcast by shuffles
#pragma unroll
    for (int k = 0; k < NMAX; ++k) {
      if (k >= n) break;
      const float dkk = __shfl_sync(0xffffffffu, S[k], k);
      const float Lk = lane >= k ? S[k] * rsqrtf(fmaxf(dkk, 1e-30f)) : 0.f;   // lane k: sqrt(dkk); lanes > k: L[a][k]
      S[k] = Lk;
#pragma unroll
      for (int j = k + 1; j < NMAX; ++j) S[j] = fmaf(-Lk, __shfl_sync(0xffffffffu, Lk, j), S[j]);
    }
    float dinv = 1.f;
#pragma unroll
    for (int b = 0; b < NMAX; ++b)
      if (b == lane) dinv = 1.f / S[b];
    // L z = weight * (S0 1)
    float z = P.weight * r0;
#pragma unroll
    for (int k = 0; k < NMAX; ++k) {
      if (k >= n) break;
      const float zk = __shfl_sync(0xffffffffu, z * dinv, k);
      if (lane == k) z = zk;
      else if (lane > k) z = fmaf(-S[k], zk, z);
    }
    // L^T t = z
    float t = 0.f;
#pragma unroll
    for (int a = NMAX - 1; a >= 0; --a) {
      if (a >= n) continue;
      const float part = (lane > a && lane < n) ? S[a] * t : 0.f;
      const float sum = warp_sum(part);
      if (lane == a) t = (z - sum) * dinv;
    }
    const float coef = lane < n ? P.weight - t : 0.f;
    // xt = W_u^T (weight - t)
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = 0; a < n; ++a) {
      const float ca = __shfl_sync(0xffffffffu, coef, a);
      const float4 v = *reinterpret_cast<const float4*>(Wt + a * SM_LD + 4 * lane);
      acc.x = fmaf(ca, v.x, acc.x); acc.y = fmaf(ca, v.y, acc.y); acc.z = fmaf(ca, v.z, acc.z); acc.w = fmaf(ca, v.w, acc.w);
    }
    float* xr = P.X + u * P.ldx + 4 * lane;
    if (4 * lane + 0 < P.d) xr[0] = acc.x;
    if (4 * lane + 1 < P.d) xr[1] = acc.y;
    if (4 * lane + 2 < P.d) xr[2] = acc.z;
    if (4 * lane + 3 < P.d) xr[3] = acc.w;
  }
}

// ---- rows with more than 32 observed columns: one CTA per row, tcgen05 Gram + register-blocked Cholesky
constexpr int RT_TILE = 16384;                 // one operand tile: 128 rows x 64 bf16, SWIZZLE_128B K-major (what a TMA box {64, 128} writes)
constexpr int RT_A_BYTES = ALS_D * (ALS_D + 1) * 4;   // the fp32 system, aliased onto the four operand tiles once the MMAs are done
constexpr int RT_WARPS = 5, RT_THREADS = RT_WARPS * 32;
constexpr int RT_TRI = NBK * (NBK + 1) / 2;            // 136 blocks of the lower triangle

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& h, __nv_bfloat16& l) {
  h = __float2bfloat16_rn(v);
  l = __float2bfloat16_rn(v - __bfloat162float(h));
}

__global__ void __launch_bounds__(RT_THREADS, 3) k_als_rows_tc(const __grid_constant__ WhiteParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;                                             // 4 x 16 KB operand tiles ...
  float* A = reinterpret_cast<float*>(smem);                         // ... later the fp32 system [128][129]
  // once the register blocks are loaded the system's shared-memory copy is dead: the factorisation's scratch lives there
  float* panel = reinterpret_cast<float*>(smem);   // [16][64] block column L_ik
  float* zrow = panel + NBK * 64;              // [8]   the right-hand side's block of the current block column
  float* diagA = zrow + 8;                     // [64]  the diagonal block about to be factored
  float* dinvs = diagA + 64;                   // [16][64] inverses of the factored diagonal blocks (row-major, lower)
  float* zvec = dinvs + NBK * 64;              // [128] z, then the solution
  float* svec = reinterpret_cast<float*>(smem + ((RT_A_BYTES + 127) & ~127));   // [5][128] per-warp partial sums / [128] row sums
  int* rlist = reinterpret_cast<int*>(svec + RT_WARPS * ALS_D);      // [160] rows of this chunk that belong here
  int* ridx = rlist + RT_THREADS;                                    // [128] the observed columns of a short row
  uint64_t* bars = reinterpret_cast<uint64_t*>(ridx + ALS_D);        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  int* wcnt = reinterpret_cast<int*>(tmem_slot + 1);                 // [5] members per warp, [5] = total, [6] = chunk id

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // 136 threads own the 8 x 8 blocks of the lower triangle (thread t <-> block (ty, tx), tx <= ty, row by row), 16 more carry
  // the right-hand side as an extra block row (one block column each), 8 idle: 5 warps instead of the 8 a square 16 x 16
  // thread grid needs -- three CTAs per SM fit (registers), and the factorisation is bound by its barrier chain, not by issue
  const bool lower = tid < RT_TRI;
  int ty = -1, tx = -2;                        // (no role matches these)
  if (lower) {
    ty = (int)((sqrtf(8.f * (float)tid + 1.f) - 1.f) * 0.5f);
    while (ty * (ty + 1) / 2 > tid) --ty;
    while ((ty + 1) * (ty + 2) / 2 <= tid) ++ty;
    tx = tid - ty * (ty + 1) / 2;
  }
  const bool is_aug = tid >= RT_TRI && tid < RT_TRI + NBK;
  const int aug_col = tid - RT_TRI;
  const float c = P.weight - 1.f, invc = 1.f / c;

  if (tid == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(ROWS, ALS_D);
  uint32_t nwait0 = 0u, nwait1 = 0u;     // completed phases of the two mbarriers (every thread keeps the same count)

  const long long n_chunks = (P.n_x + RT_THREADS - 1) / RT_THREADS;
  for (;;) {
    // ---- take the next chunk of 160 rows (long rows cost 10x a short one: static striding leaves SMs idle at the end)
    __syncthreads();                       // (the previous chunk's list and counters are no longer read)
    if (tid == 0) wcnt[RT_WARPS + 1] = atomicAdd(P.cursor, 1);
    __syncthreads();
    const long long chunk0 = (long long)wcnt[RT_WARPS + 1] * RT_THREADS;
    if (wcnt[RT_WARPS + 1] >= n_chunks) break;
    {
      const long long u = chunk0 + tid;
      const bool mine = u < P.n_x && (P.indptr[u + 1] - P.indptr[u]) > 32;
      const unsigned m = __ballot_sync(0xffffffffu, mine);
      if (lane == 0) wcnt[warp] = __popc(m);
      __syncthreads();
      int base = 0;
      for (int w = 0; w < warp; ++w) base += wcnt[w];
      if (mine) rlist[base + __popc(m & ((1u << lane) - 1u))] = tid;
      if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < RT_WARPS; ++w) tot += wcnt[w];
        wcnt[RT_WARPS] = tot;
      }
      __syncthreads();
    }
    const int nmine = wcnt[RT_WARPS];
    for (int ri = 0; ri < nmine; ++ri) {
      const long long u = chunk0 + rlist[ri];
      const long long lo = P.indptr[u];
      const long long nl = P.indptr[u + 1] - lo;
      const bool wide = nl > ALS_D;         // 128 x 128 system (I + c W^T W); otherwise n x n (I / c + W W^T)
      const int n = wide ? ALS_D : (int)nl; // order of the system
      // =========================================================== Gram on the tensor cores
      if (!wide) {
        // tiles: hi[kc 0], hi[kc 1], lo[kc 0], lo[kc 1]; row a of the tile = gathered row a, zero beyond n
        // (warp w takes the tile rows w, w + 5, ...: 26 at most, in two batches of 13 whose loads are all issued before the
        // first conversion -- two round trips, not twenty-six)
        const int my_a = warp + RT_WARPS * lane;
        const int my_idx = (lane < 26 && my_a < n) ? P.indices[lo + my_a] : -1;
        if (lane < 26 && my_a < n) ridx[my_a] = my_idx;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          float4 v[13];
#pragma unroll
          for (int r = 0; r < 13; ++r) {
            const int ia = __shfl_sync(0xffffffffu, my_idx, 13 * half + r);
            v[r] = ia >= 0 ? __ldg(reinterpret_cast<const float4*>(P.W + (long long)ia * ALS_D) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int r = 0; r < 13; ++r) {
            const int a = warp + RT_WARPS * (13 * half + r);
            if (a < ALS_D) {
              __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
              split_bf16(v[r].x, h0, l0); split_bf16(v[r].y, h1, l1); split_bf16(v[r].z, h2, l2); split_bf16(v[r].w, h3, l3);
              const int off = (lane >> 4) * RT_TILE + a * 128 + ((((lane & 15) >> 1) ^ (a & 7)) << 4) + (lane & 1) * 8;
              *reinterpret_cast<uint2*>(tiles + off) = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
              *reinterpret_cast<uint2*>(tiles + 2 * RT_TILE + off) = make_uint2(pack_bf16(l0, l1), pack_bf16(l2, l3));
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
          tc_fence_after();
          const uint32_t hb = smem_u32(tiles), lb = smem_u32(tiles + 2 * RT_TILE);
#pragma unroll
          for (int kc = 0; kc < 2; ++kc)
#pragma unroll
            for (int k = 0; k < KCH / 16; ++k) {
              const uint64_t dh = umma_desc_sw128(hb + kc * RT_TILE + k * 32), dl = umma_desc_sw128(lb + kc * RT_TILE + k * 32);
              tc_mma_bf16(tmem_base, dh, dh, idesc, (kc | k) ? 1u : 0u);
              tc_mma_bf16(tmem_base, dh, dl, idesc, 1u);
              tc_mma_bf16(tmem_base, dl, dh, idesc, 1u);
            }
          tc_commit(bars);
        }
        mbar_wait(bars, nwait0 & 1u);
        ++nwait0;
      } else {
        // chunks of 64 gathered rows; tile[i][a - 64 ch] = W[a][i] (the contraction runs over the gathered rows), two buffers
        const int nch = (int)((nl + 63) >> 6);
        float sacc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int ch = 0; ch < nch; ++ch) {
          const int buf = ch & 1;
          if (ch >= 2) {                      // the MMAs of chunk ch - 2 have finished reading this buffer
            if (buf == 0) { mbar_wait(bars, nwait0 & 1u); ++nwait0; }
            else { mbar_wait(bars + 1, nwait1 & 1u); ++nwait1; }
          }
          uint8_t* th = tiles + buf * 2 * RT_TILE;
#pragma unroll 1
          for (int sc = warp; sc < 8; sc += RT_WARPS) {   // eight 16-byte column chunks of the tile over five warps
          const long long a0 = (long long)ch * 64 + sc * 8;
          long long src[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) src[q] = (a0 + q < nl) ? (long long)P.indices[lo + a0 + q] * ALS_D : -1;
#pragma unroll
          for (int ig = 0; ig < 4; ++ig) {
            const int i = lane + 32 * ig;
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = src[q] >= 0 ? __ldg(P.W + src[q] + i) : 0.f;
            uint32_t hp[4], lp[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              __nv_bfloat16 ha, hb2, la, lb2;
              split_bf16(v[2 * q], ha, la);
              split_bf16(v[2 * q + 1], hb2, lb2);
              hp[q] = pack_bf16(ha, hb2);
              lp[q] = pack_bf16(la, lb2);
              sacc[ig] += v[2 * q] + v[2 * q + 1];
            }
            const int off = i * 128 + ((sc ^ (i & 7)) << 4);
            *reinterpret_cast<uint4*>(th + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
            *reinterpret_cast<uint4*>(th + RT_TILE + off) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
          }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncthreads();
          if (tid == 0) {
            tc_fence_after();
            const uint32_t hb = smem_u32(th), lb = smem_u32(th + RT_TILE);
#pragma unroll
            for (int k = 0; k < KCH / 16; ++k) {
              const uint64_t dh = umma_desc_sw128(hb + k * 32), dl = umma_desc_sw128(lb + k * 32);
              tc_mma_bf16(tmem_base, dh, dh, idesc, (ch | k) ? 1u : 0u);
              tc_mma_bf16(tmem_base, dh, dl, idesc, 1u);
              tc_mma_bf16(tmem_base, dl, dh, idesc, 1u);
            }
            tc_commit(bars + buf);
          }
        }
        // drain: the last one or two chunks
        if (nch >= 2) {
          if (((nch - 2) & 1) == 0) { mbar_wait(bars, nwait0 & 1u); ++nwait0; }
          else { mbar_wait(bars + 1, nwait1 & 1u); ++nwait1; }
        }
        if (((nch - 1) & 1) == 0) { mbar_wait(bars, nwait0 & 1u); ++nwait0; }
        else { mbar_wait(bars + 1, nwait1 & 1u); ++nwait1; }
#pragma unroll
        for (int ig = 0; ig < 4; ++ig) svec[warp * ALS_D + lane + 32 * ig] = sacc[ig];
      }
      tc_fence_after();
      // =========================================================== accumulator -> the fp32 system in shared memory
      if (warp < 4) {                        // warp q reads TMEM lane quadrant q: one matrix row per thread
        const int row = warp * 32 + lane;
        float rs = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t r[32];
          tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cc * 32), r);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = cc * 32 + j;
            const float g = __uint_as_float(r[j]);
            rs += g;
            A[row * (ALS_D + 1) + col] = wide ? fmaf(c, g, row == col ? 1.f : 0.f) : g + (row == col ? invc : 0.f);
          }
        }
        if (!wide) svec[row] = rs;
      }
      tc_fence_before();
      __syncthreads();
      // =========================================================== register blocks + right-hand side
      float a[BS][BS];
      float g[BS];
      if (lower) {
#pragma unroll
        for (int i = 0; i < BS; ++i)
#pragma unroll
          for (int j = 0; j < BS; ++j) a[i][j] = A[(BS * ty + i) * (ALS_D + 1) + BS * tx + j];
      }
      if (is_aug) {
#pragma unroll
        for (int j = 0; j < BS; ++j) {
          const int e = BS * aug_col + j;
          float s = 0.f;
          if (wide) {
#pragma unroll
            for (int w = 0; w < RT_WARPS; ++w) s += svec[w * ALS_D + e];
          } else {
            s = e < n ? svec[e] : 0.f;
          }
          g[j] = P.weight * s;
        }
      }
      __syncthreads();                       // every block is in registers: the system's shared-memory copy becomes scratch
      if (ty == 0 && tx == 0) {
#pragma unroll
        for (int i = 0; i < BS; ++i)
#pragma unroll
          for (int j = 0; j < BS; ++j) diagA[i * BS + j] = a[i][j];
      }
      __syncthreads();
      const int nb = (n + BS - 1) / BS;
      // =========================================================== blocked Cholesky, two barriers per block column
      for (int kb = 0; kb < nb; ++kb) {
        const bool in_panel = tx == kb && ty >= kb && ty < nb;
        const bool aug_here = is_aug && aug_col == kb;
        if (in_panel || aug_here) {
          float dg[BS][BS], inv[BS];
#pragma unroll
          for (int i = 0; i < BS; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) dg[i][j] = diagA[i * BS + j];
#pragma unroll
          for (int k = 0; k < BS; ++k) {
            inv[k] = rsqrtf(fmaxf(dg[k][k], 1e-30f));
            dg[k][k] *= inv[k];
#pragma unroll
            for (int i = k + 1; i < BS; ++i) dg[i][k] *= inv[k];
#pragma unroll
            for (int i = k + 1; i < BS; ++i)
#pragma unroll
              for (int j = k + 1; j <= i; ++j) dg[i][j] = fmaf(-dg[i][k], dg[j][k], dg[i][j]);
          }
          if (aug_here) {                    // z_kb = g L_kk^-T
#pragma unroll
            for (int j = 0; j < BS; ++j) {
              float v = g[j];
#pragma unroll
              for (int q = 0; q < j; ++q) v = fmaf(-g[q], dg[j][q], v);
              g[j] = v * inv[j];
            }
#pragma unroll
            for (int j = 0; j < BS; ++j) {
              zrow[j] = g[j];
              zvec[BS * kb + j] = g[j];
            }
          } else if (ty == kb) {             // the diagonal thread publishes L_kk^-1 for the back-substitution (off the critical path:
                                             // the panel threads are busy with their 8 x 8 solves meanwhile)
#pragma unroll
            for (int j = 0; j < BS; ++j) {
              float xc[BS];
              xc[j] = inv[j];
#pragma unroll
              for (int i = j + 1; i < BS; ++i) {
                float v = 0.f;
#pragma unroll
                for (int q = j; q < i; ++q) v = fmaf(dg[i][q], xc[q], v);
                xc[i] = -v * inv[i];
              }
#pragma unroll
              for (int i = 0; i < BS; ++i) dinvs[kb * 64 + i * BS + j] = i >= j ? xc[i] : 0.f;
            }
          } else {                           // L_ik = A_ik L_kk^-T
#pragma unroll
            for (int i = 0; i < BS; ++i)
#pragma unroll
              for (int j = 0; j < BS; ++j) {
                float v = a[i][j];
#pragma unroll
                for (int q = 0; q < j; ++q) v = fmaf(-a[i][q], dg[j][q], v);
                a[i][j] = v * inv[j];
              }
#pragma unroll
            for (int i = 0; i < BS; ++i)
#pragma unroll
              for (int j = 0; j < BS; ++j) panel[ty * 64 + i * BS + j] = a[i][j];
          }
        }
        __syncthreads();
        if (lower && tx > kb && ty < nb) {   // A_ij -= L_ik L_jk^T
          const float* li = panel + ty * 64;
          const float* lj = panel + tx * 64;
#pragma unroll
          for (int q = 0; q < BS; ++q) {
            float ci[BS], cj[BS];
#pragma unroll
            for (int i = 0; i < BS; ++i) {
              ci[i] = li[i * BS + q];
              cj[i] = lj[i * BS + q];
            }
#pragma unroll
            for (int i = 0; i < BS; ++i)
#pragma unroll
              for (int j = 0; j < BS; ++j) a[i][j] = fmaf(-ci[i], cj[j], a[i][j]);
          }
          if (ty == kb + 1 && tx == kb + 1) {
#pragma unroll
            for (int i = 0; i < BS; ++i)
#pragma unroll
              for (int j = 0; j < BS; ++j) diagA[i * BS + j] = a[i][j];
          }
        }
        if (is_aug && aug_col > kb && aug_col < nb) {   // g_j -= z_kb L_jk^T
          const float* lj = panel + aug_col * 64;
#pragma unroll
          for (int j = 0; j < BS; ++j) {
            float v = g[j];
#pragma unroll
            for (int q = 0; q < BS; ++q) v = fmaf(-zrow[q], lj[j * BS + q], v);
            g[j] = v;
          }
        }
        __syncthreads();
      }
      // =========================================================== L^T x = z: x_kb = L_kk^-T z_kb with the published inverse
      // blocks (a mat-vec, no division chain); block row kb then removes x_kb from every z_j, j < kb, out of its register
      // blocks, and the thread of block (kb, kb - 1) -- the last one to touch z_(kb-1) -- finishes x_(kb-1) right away: one
      // barrier per block column
      if (ty == nb - 1 && tx == nb - 1) {
        const float* di = dinvs + (nb - 1) * 64;
        float xb[BS];
#pragma unroll
        for (int j = 0; j < BS; ++j) {
          float v = 0.f;
#pragma unroll
          for (int i = j; i < BS; ++i) v = fmaf(di[i * BS + j], zvec[BS * (nb - 1) + i], v);
          xb[j] = v;
        }
#pragma unroll
        for (int j = 0; j < BS; ++j) zvec[BS * (nb - 1) + j] = xb[j];
      }
      __syncthreads();
      for (int kb = nb - 1; kb >= 1; --kb) {
        if (ty == kb && tx < kb) {           // z_tx -= L_(kb,tx)^T x_kb
          float xb[BS], zj[BS];
#pragma unroll
          for (int i = 0; i < BS; ++i) xb[i] = zvec[BS * kb + i];
#pragma unroll
          for (int j = 0; j < BS; ++j) {
            float v = zvec[BS * tx + j];
#pragma unroll
            for (int i = 0; i < BS; ++i) v = fmaf(-a[i][j], xb[i], v);
            zj[j] = v;
          }
          if (tx == kb - 1) {                // z_(kb-1) is complete: x_(kb-1) = L^-T z_(kb-1)
            const float* di = dinvs + (kb - 1) * 64;
#pragma unroll
            for (int j = 0; j < BS; ++j) {
              float v = 0.f;
#pragma unroll
              for (int i = j; i < BS; ++i) v = fmaf(di[i * BS + j], zj[i], v);
              zvec[BS * tx + j] = v;
            }
          } else {
#pragma unroll
            for (int j = 0; j < BS; ++j) zvec[BS * tx + j] = zj[j];
          }
        }
        __syncthreads();
      }
      // =========================================================== xt
      if (wide) {
        if (tid < P.d) P.X[u * P.ldx + tid] = zvec[tid];
      } else {
        // xt = W_u^T (weight - t): warp w takes the gathered rows w, w + 5, ... (13 loads in flight at a time), a lane four
        // columns; the five partial sums meet in shared memory
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          float4 v[13];
#pragma unroll
          for (int r = 0; r < 13; ++r) {
            const int e = warp + RT_WARPS * (13 * half + r);
            v[r] = e < n ? __ldg(reinterpret_cast<const float4*>(P.W + (long long)ridx[e] * ALS_D) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int r = 0; r < 13; ++r) {
            const int e = warp + RT_WARPS * (13 * half + r);
            const float cf = e < n ? P.weight - zvec[e] : 0.f;
            acc.x = fmaf(cf, v[r].x, acc.x); acc.y = fmaf(cf, v[r].y, acc.y); acc.z = fmaf(cf, v[r].z, acc.z); acc.w = fmaf(cf, v[r].w, acc.w);
          }
        }
        *reinterpret_cast<float4*>(svec + warp * ALS_D + 4 * lane) = acc;
        __syncthreads();
        if (tid < P.d) {
          float x = 0.f;
#pragma unroll
          for (int w = 0; w < RT_WARPS; ++w) x += svec[w * ALS_D + tid];
          P.X[u * P.ldx + tid] = x;
        }
      }
      __syncthreads();                       // zvec / svec / A are rewritten by the next row
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

typedef CUresult (*encode_tiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_plane_map(CUtensorMap* tm, void* base, long long n_pad) {
  static encode_tiled_t enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<encode_tiled_t>(p);
  }
  CF_CHECK_ARG(enc != nullptr, "cf_als_half_sweep: cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)n_pad, (cuuint64_t)ALS_D};
  const cuuint64_t strides[1] = {(cuuint64_t)n_pad * 2};
  const cuuint32_t box[2] = {KCH, ROWS};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CF_CHECK_ARG(r == CUDA_SUCCESS, "cf_als_half_sweep: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

}  // namespace

// workspace layout: [hi | lo bf16 planes of Y^T] [G fp32 128x128] [L^-1 fp32 128x128] [L^-T fp32 128x128] [cursor] [W fp32 n_y x 128]
static size_t ws_off_g(long long n_pad) { return (((size_t)2 * ALS_D * n_pad * 2 + 1023) / 1024) * 1024; }
static size_t ws_off_linv(long long n_pad) { return ws_off_g(n_pad) + (size_t)ALS_D * ALS_D * 4; }
static size_t ws_off_linvt(long long n_pad) { return ws_off_linv(n_pad) + (size_t)ALS_D * ALS_D * 4; }
static size_t ws_off_cursor(long long n_pad) { return ws_off_linvt(n_pad) + (size_t)ALS_D * ALS_D * 4; }
static size_t ws_off_w(long long n_pad) { return ws_off_cursor(n_pad) + 1024; }

extern "C" int64_t cf_als_workspace_bytes(int64_t n_y) {
  const long long n_pad = (n_y + KCH - 1) / KCH * KCH;
  return (int64_t)(ws_off_w(n_pad) + (size_t)n_y * ALS_D * 4 + 2048);
}

// Gram stage: G += Y^T Y over the given rows (bf16 hi/lo planes + tcgen05); G is NOT zeroed here
static int als_gram(const float* Y, long long n_y, int d, int ldy, float* G, uint8_t* ws, cudaStream_t stream) {
  const long long n_pad = (n_y + KCH - 1) / KCH * KCH;
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* lo = hi + (size_t)ALS_D * n_pad;
  dim3 sgrid((unsigned)((n_pad + 31) / 32), ALS_D / 32);
  k_als_split<<<sgrid, 256, 0, stream>>>(Y, n_y, n_pad, d, ldy, hi, lo);
  CUtensorMap tmHi, tmLo;
  if (int rc = make_plane_map(&tmHi, hi, n_pad)) return rc;
  if (int rc = make_plane_map(&tmLo, lo, n_pad)) return rc;
  const int n_chunks = (int)(n_pad / KCH);
  int ggrid = cf_num_sms();
  if (ggrid > n_chunks) ggrid = n_chunks;
  const size_t gsmem = 2 * CHUNK_BYTES + 64 + 1024;
  CF_CUDA_OK(cudaFuncSetAttribute(k_gram_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
  k_gram_tc<<<ggrid, 192, gsmem, stream>>>(tmHi, tmLo, n_chunks, G);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

constexpr int SMALL16_WARPS = 8, SMALL32_WARPS = 5;
static size_t small_smem(int nmax, int warps) { return (size_t)warps * (nmax * SM_LD + nmax * (nmax + 1)) * 4; }
static size_t rows_tc_smem() {
  return (size_t)((RT_A_BYTES + 127) & ~127) + (size_t)RT_WARPS * ALS_D * 4 + (RT_THREADS + ALS_D) * 4 + 16 + 4 + (RT_WARPS + 2) * 4 + 64 + 1024;
}

// Solve stage: every row of X from the (complete) Gram G and its observed rows of Y.  With a workspace and weight > 1 the
// whitened path above; otherwise (or with CF_ALS_DIRECT=1) the direct 128 x 128 solve of every row.
static int als_solve(const cf_als_args* a, const float* G, cudaStream_t stream, uint8_t* ws) {
  const int sms = cf_num_sms();
  if (ws != nullptr && a->weight > 1.f && !getenv("CF_ALS_DIRECT")) {
    const long long n_pad = (a->n_y + KCH - 1) / KCH * KCH;
    float* Linv = reinterpret_cast<float*>(ws + ws_off_linv(n_pad));
    float* LinvT = reinterpret_cast<float*>(ws + ws_off_linvt(n_pad));
    int* cursor = reinterpret_cast<int*>(ws + ws_off_cursor(n_pad));
    float* W = reinterpret_cast<float*>(ws + ws_off_w(n_pad));
    const size_t psmem = (size_t)ALS_D * (ALS_D + 1) * 8;
    CF_CUDA_OK(cudaFuncSetAttribute(k_als_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
    k_als_prep<<<1, 1024, psmem, stream>>>(G, a->reg, a->d, Linv, LinvT);
    const size_t msmem = ((size_t)ALS_D * ALS_D + 32 * SM_LD) * 4;
    CF_CUDA_OK(cudaFuncSetAttribute(k_als_mul, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    long long mg = (a->n_y + 31) / 32;
    if (mg > (long long)sms * 2) mg = (long long)sms * 2;
    k_als_mul<<<(unsigned)mg, 256, msmem, stream>>>(a->Y, a->n_y, a->d, a->ldy, LinvT, W, ALS_D, ALS_D);   // W = Y L^-T
    CF_CUDA_OK(cudaMemsetAsync(cursor, 0, 4, stream));
    WhiteParams P;
    P.X = a->X; P.W = W; P.indptr = (const long long*)a->indptr; P.indices = a->indices; P.n_x = a->n_x; P.d = a->d;
    P.ldx = a->ldx; P.weight = a->weight; P.cursor = cursor;
    const size_t s16 = small_smem(16, SMALL16_WARPS), s32 = small_smem(32, SMALL32_WARPS), srt = rows_tc_smem();
    CF_CUDA_OK(cudaFuncSetAttribute(k_als_small<16, SMALL16_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s16));
    CF_CUDA_OK(cudaFuncSetAttribute(k_als_small<32, SMALL32_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s32));
    CF_CUDA_OK(cudaFuncSetAttribute(k_als_rows_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)srt));
    long long g16 = (a->n_x + SMALL16_WARPS - 1) / SMALL16_WARPS, g32 = (a->n_x + SMALL32_WARPS - 1) / SMALL32_WARPS;
    if (g16 > (long long)sms * 2) g16 = (long long)sms * 2;
    if (g32 > (long long)sms * 2) g32 = (long long)sms * 2;
    long long grt = (a->n_x + RT_THREADS - 1) / RT_THREADS;
    if (grt > (long long)sms * 3) grt = (long long)sms * 3;
    k_als_rows_tc<<<(unsigned)grt, RT_THREADS, srt, stream>>>(P);      // the long rows first: they are the tail
    k_als_small<32, SMALL32_WARPS><<<(unsigned)g32, SMALL32_WARPS * 32, s32, stream>>>(P);
    k_als_small<16, SMALL16_WARPS><<<(unsigned)g16, SMALL16_WARPS * 32, s16, stream>>>(P);
    long long xg = (a->n_x + 31) / 32;
    if (xg > (long long)sms * 2) xg = (long long)sms * 2;
    k_als_mul<<<(unsigned)xg, 256, msmem, stream>>>(a->X, a->n_x, a->d, a->ldx, Linv, a->X, a->ldx, a->d);   // X = Xt L^-1
    CF_CUDA_OK(cudaGetLastError());
    return 0;
  }
  SolveParams S;
  S.X = a->X; S.Y = a->Y; S.G = G; S.indptr = (const long long*)a->indptr; S.indices = a->indices;
  S.n_x = a->n_x; S.d = a->d; S.ldx = a->ldx; S.ldy = a->ldy; S.weight = a->weight; S.reg = a->reg;
  long long sg = a->n_x;
  const long long cap = (long long)sms * 8;
  if (sg > cap) sg = cap;
  if (getenv("CF_ALS_COLUMNWISE")) {   // the first, column-wise kernel (kept for A/B measurements)
    const size_t ssmem = ((size_t)a->d * (a->d + 1) + a->d + 8 * a->d) * 4;
    CF_CUDA_OK(cudaFuncSetAttribute(k_als_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));
    k_als_solve<<<(unsigned)sg, 256, ssmem, stream>>>(S);
  } else {
    const size_t ssmem = ((size_t)ALS_D * (ALS_D + 1) + NBK * 64 + 8 * ALS_D + ALS_D + 64) * 4;
    CF_CUDA_OK(cudaFuncSetAttribute(k_als_solve_blocked, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));
    k_als_solve_blocked<<<(unsigned)sg, 256, ssmem, stream>>>(S);
  }
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

static int als_check(const cf_als_args* a, const char* who, bool need_ws) {
  CF_CHECK_ARG(a != nullptr, "%s: args is NULL", who);
  CF_CHECK_ARG(a->X && a->Y && a->indptr && a->indices, "%s: NULL pointer", who);
  CF_CHECK_ARG(a->d > 0 && a->d <= ALS_D && a->ldx >= a->d && a->ldy >= a->d, "%s: n_factors up to %d are supported (d=%d)", who, ALS_D, a->d);
  CF_CHECK_ARG(a->n_x > 0 && a->n_y > 0, "%s: empty factor matrix", who);
  CF_CHECK_ARG(a->reg > 0.f, "%s: reg must be positive (it keeps the normal equations SPD)", who);
  if (need_ws)
    CF_CHECK_ARG(a->workspace && ((uintptr_t)a->workspace % 1024) == 0 && a->workspace_bytes >= cf_als_workspace_bytes(a->n_y), "%s: workspace too small or not 1024-byte aligned", who);
  return 0;
}

extern "C" int cf_als_half_sweep(const cf_als_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = als_check(a, "cf_als_half_sweep", true)) return rc;
  const long long n_pad = (a->n_y + KCH - 1) / KCH * KCH;
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  float* G = reinterpret_cast<float*>(ws + ws_off_g(n_pad));
  CF_CUDA_OK(cudaMemsetAsync(G, 0, ALS_D * ALS_D * 4, stream));
  if (int rc = als_gram(a->Y, a->n_y, a->d, a->ldy, G, ws, stream)) return rc;
  return als_solve(a, G, stream, ws);
}

extern "C" int cf_als_gram(const float* Y, int64_t n_y, int32_t d, int32_t ldy, float* G, void* workspace,
                           int64_t workspace_bytes, void* stream_) {
  CF_CHECK_ARG(Y && G && workspace, "cf_als_gram: NULL pointer");
  CF_CHECK_ARG(d > 0 && d <= ALS_D && ldy >= d && n_y > 0, "cf_als_gram: bad shape (d=%d, n_y=%lld)", d, (long long)n_y);
  CF_CHECK_ARG(((uintptr_t)workspace % 1024) == 0 && workspace_bytes >= cf_als_workspace_bytes(n_y), "cf_als_gram: workspace too small or not 1024-byte aligned");
  return als_gram(Y, n_y, d, ldy, G, reinterpret_cast<uint8_t*>(workspace), (cudaStream_t)stream_);
}

extern "C" int cf_als_solve_rows(const cf_als_args* a, const float* G, void* stream_) {
  if (int rc = als_check(a, "cf_als_solve_rows", false)) return rc;
  CF_CHECK_ARG(G != nullptr, "cf_als_solve_rows: G is NULL");
  uint8_t* ws = nullptr;      // optional: with a workspace of cf_als_workspace_bytes(n_y) bytes the short rows take the low-rank path
  if (a->workspace != nullptr) {
    CF_CHECK_ARG(((uintptr_t)a->workspace % 1024) == 0 && a->workspace_bytes >= cf_als_workspace_bytes(a->n_y), "cf_als_solve_rows: workspace too small or not 1024-byte aligned");
    ws = reinterpret_cast<uint8_t*>(a->workspace);
  }
  return als_solve(a, G, (cudaStream_t)stream_, ws);
}
