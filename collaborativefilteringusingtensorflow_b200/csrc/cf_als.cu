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
  int skip_le;          // rows with at most this many observed columns are solved by k_als_woodbury instead (-1: none)
  const float* Z;       // [n_y, 128] = Y (G + reg I)^-1 (Woodbury path)
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
    if (P.indptr[u + 1] - P.indptr[u] <= P.skip_le) continue;   // block-uniform: a short row, solved by the low-rank kernel
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

// ---- low-rank (Woodbury) path for rows with few observed columns.  Measured on a configs[3]-shaped slice (1 M rows, 50 M
// interactions, median 21 per row): the register-blocked solve above spends ~260 k cycles per row, most of them in the
// 128 x 128 Cholesky whatever the row's length.  But A_u = G0 + c Y_u^T Y_u (G0 = Y^T Y + reg I, c = weight - 1) is a rank-n_u
// update of a matrix that is the SAME for every row, so with Z = Y G0^-1 (one GEMM per half-sweep)
//     x_u = p - Z_u^T (I / c + Y_u Z_u^T)^-1 (Y_u p),   p = weight * sum_a z_a
// needs an n_u x n_u Cholesky instead: 9 k flops at the median row instead of 700 k.
constexpr int WB_N = 64;      // longest row solved this way (shared memory: two [64][129] row tiles + S[64][65])

// G0^-1 in fp64 by Gauss-Jordan without pivoting (G0 is SPD); one block, the [128][256] tableau lives in global scratch
__global__ void __launch_bounds__(1024) k_als_inverse(const float* __restrict__ G, float reg, int d, double* __restrict__ W,
                                                       float* __restrict__ Ginv) {
  const int tid = threadIdx.x;
  for (int e = tid; e < ALS_D * 2 * ALS_D; e += blockDim.x) {
    const int i = e / (2 * ALS_D), j = e % (2 * ALS_D);
    double v;
    if (j < ALS_D) v = (i < d && j < d) ? (double)G[i * ALS_D + j] + (i == j ? (double)reg : 0.0) : (i == j ? 1.0 : 0.0);
    else v = (j - ALS_D == i) ? 1.0 : 0.0;
    W[e] = v;
  }
  __syncthreads();
  __shared__ double s_col[ALS_D];
  __shared__ double s_piv;
  for (int k = 0; k < ALS_D; ++k) {
    if (tid == 0) s_piv = 1.0 / W[k * 2 * ALS_D + k];
    __syncthreads();
    for (int j = tid; j < 2 * ALS_D; j += blockDim.x) W[k * 2 * ALS_D + j] *= s_piv;
    if (tid < ALS_D) s_col[tid] = W[tid * 2 * ALS_D + k];
    __syncthreads();
    for (int e = tid; e < ALS_D * 2 * ALS_D; e += blockDim.x) {
      const int i = e / (2 * ALS_D), j = e % (2 * ALS_D);
      if (i != k) W[e] -= s_col[i] * W[k * 2 * ALS_D + j];
    }
    __syncthreads();
  }
  for (int e = tid; e < ALS_D * ALS_D; e += blockDim.x) {
    const int i = e / ALS_D, j = e % ALS_D;
    Ginv[e] = (i < d && j < d) ? (float)W[i * 2 * ALS_D + ALS_D + j] : 0.f;
  }
}

// Z[n, 128] = Y[n, :d] Ginv[:d, :128]  (fp32; 32 rows per block, Ginv resident in shared memory)
__global__ void __launch_bounds__(256) k_als_z(const float* __restrict__ Y, long long n, int d, int ldy, const float* __restrict__ Ginv,
                                               float* __restrict__ Z) {
  extern __shared__ float sm[];
  float* Gs = sm;                    // [128][128]
  float* Ys = Gs + ALS_D * ALS_D;    // [32][128]
  for (int e = threadIdx.x; e < ALS_D * ALS_D; e += blockDim.x) Gs[e] = Ginv[e];
  const int col = threadIdx.x & 127, half = threadIdx.x >> 7;     // thread: one column, 16 of the 32 rows
  for (long long r0 = (long long)blockIdx.x * 32; r0 < n; r0 += (long long)gridDim.x * 32) {
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * ALS_D; e += blockDim.x) {
      const long long r = r0 + e / ALS_D;
      const int k = e % ALS_D;
      Ys[e] = (r < n && k < d) ? Y[r * ldy + k] : 0.f;
    }
    __syncthreads();
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int k = 0; k < d; ++k) {
      const float g = Gs[k * ALS_D + col];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(Ys[(half * 16 + i) * ALS_D + k], g, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const long long r = r0 + half * 16 + i;
      if (r < n) Z[r * ALS_D + col] = acc[i];
    }
  }
}

__global__ void __launch_bounds__(128) k_als_woodbury(const __grid_constant__ SolveParams P) {
  extern __shared__ float sm[];
  constexpr int LD = ALS_D + 1;
  float* Ys = sm;                      // [WB_N][129]
  float* Zs = Ys + WB_N * LD;          // [WB_N][129]
  float* S = Zs + WB_N * LD;           // [WB_N][WB_N + 1]
  float* p = S + WB_N * (WB_N + 1);    // [128]
  float* r = p + ALS_D;                // [WB_N]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float c = P.weight - 1.f;
  for (long long u = blockIdx.x; u < P.n_x; u += gridDim.x) {
    const long long lo = P.indptr[u];
    const int n = (int)(P.indptr[u + 1] - lo);
    if (n > P.skip_le) continue;       // block-uniform: a long row, solved by k_als_solve_blocked
    __syncthreads();
    if (n == 0) {                      // nothing observed: b = 0, so x = 0
      if (tid < P.d) P.X[u * P.ldx + tid] = 0.f;
      continue;
    }
    for (int a = warp; a < n; a += 4) {          // gather the observed rows of Y and Z
      const long long i = P.indices[lo + a];
      for (int k = lane; k < ALS_D; k += 32) {
        Ys[a * LD + k] = k < P.d ? P.Y[i * P.ldy + k] : 0.f;
        Zs[a * LD + k] = P.Z[i * ALS_D + k];
      }
    }
    __syncthreads();
    {                                            // p = weight * sum_a z_a
      float acc = 0.f;
      for (int a = 0; a < n; ++a) acc += Zs[a * LD + tid];
      p[tid] = P.weight * acc;
    }
    __syncthreads();
    // S = I / c + Y_u Z_u^T (lower triangle) and r = Y_u p: one dot product of length 128 per (a, b <= a) / per a, by warps
    const int npair = n * (n + 1) / 2;
    for (int e = warp; e < npair + n; e += 4) {
      int a, b;
      const float* rhs;
      if (e < npair) {
        a = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
        while (a * (a + 1) / 2 > e) --a;
        while ((a + 1) * (a + 2) / 2 <= e) ++a;
        b = e - a * (a + 1) / 2;
        rhs = Zs + b * LD;
      } else {
        a = e - npair;
        b = -1;
        rhs = p;
      }
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < ALS_D; k += 32) acc = fmaf(Ys[a * LD + k + lane], rhs[k + lane], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        if (b >= 0) S[a * (WB_N + 1) + b] = acc + (a == b ? 1.f / c : 0.f);
        else r[a] = acc;
      }
    }
    __syncthreads();
    // Cholesky of S (n x n, right-looking by columns), then L z = r, L^T t = z
    for (int k = 0; k < n; ++k) {
      if (tid == 0) S[k * (WB_N + 1) + k] = sqrtf(fmaxf(S[k * (WB_N + 1) + k], 1e-30f));
      __syncthreads();
      const float inv = 1.f / S[k * (WB_N + 1) + k];
      for (int i = k + 1 + tid; i < n; i += blockDim.x) S[i * (WB_N + 1) + k] *= inv;
      __syncthreads();
      const int m = n - k - 1;
      for (int e = tid; e < m * m; e += blockDim.x) {
        const int i = k + 1 + e / m, j = k + 1 + e % m;
        if (j <= i) S[i * (WB_N + 1) + j] = fmaf(-S[i * (WB_N + 1) + k], S[j * (WB_N + 1) + k], S[i * (WB_N + 1) + j]);
      }
      __syncthreads();
    }
    if (warp == 0) {
      for (int k = 0; k < n; ++k) {
        float acc = 0.f;
        for (int j = lane; j < k; j += 32) acc = fmaf(S[k * (WB_N + 1) + j], r[j], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) r[k] = (r[k] - acc) / S[k * (WB_N + 1) + k];
        __syncwarp();
      }
      for (int k = n - 1; k >= 0; --k) {
        float acc = 0.f;
        for (int j = k + 1 + lane; j < n; j += 32) acc = fmaf(S[j * (WB_N + 1) + k], r[j], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) r[k] = (r[k] - acc) / S[k * (WB_N + 1) + k];
        __syncwarp();
      }
    }
    __syncthreads();
    {                                            // x = p - Z_u^T t
      float acc = p[tid];
      for (int a = 0; a < n; ++a) acc = fmaf(-Zs[a * LD + tid], r[a], acc);
      if (tid < P.d) P.X[u * P.ldx + tid] = acc;
    }
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

// workspace layout: [hi | lo bf16 planes of Y^T] [G fp32 128x128] [Ginv fp32 128x128] [Gauss-Jordan tableau fp64 128x256] [Z fp32 n_y x 128]
static size_t ws_off_g(long long n_pad) { return (((size_t)2 * ALS_D * n_pad * 2 + 1023) / 1024) * 1024; }
static size_t ws_off_ginv(long long n_pad) { return ws_off_g(n_pad) + (size_t)ALS_D * ALS_D * 4; }
static size_t ws_off_w(long long n_pad) { return ws_off_ginv(n_pad) + (size_t)ALS_D * ALS_D * 4; }
static size_t ws_off_z(long long n_pad) { return ws_off_w(n_pad) + (size_t)ALS_D * 2 * ALS_D * 8; }

extern "C" int64_t cf_als_workspace_bytes(int64_t n_y) {
  const long long n_pad = (n_y + KCH - 1) / KCH * KCH;
  return (int64_t)(ws_off_z(n_pad) + (size_t)n_y * ALS_D * 4 + 2048);
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

// Solve stage: every row of X from the (complete) Gram G and its observed rows of Y.  With a workspace (and weight > 1) the
// rows with at most WB_N observed columns take the low-rank path, the others the full 128 x 128 solve.
static int als_solve(const cf_als_args* a, const float* G, cudaStream_t stream, uint8_t* ws) {
  SolveParams S;
  S.X = a->X; S.Y = a->Y; S.G = G; S.indptr = (const long long*)a->indptr; S.indices = a->indices;
  S.n_x = a->n_x; S.d = a->d; S.ldx = a->ldx; S.ldy = a->ldy; S.weight = a->weight; S.reg = a->reg;
  S.skip_le = -1; S.Z = nullptr;
  if (ws != nullptr && a->weight > 1.f && !getenv("CF_ALS_DIRECT")) {
    const long long n_pad = (a->n_y + KCH - 1) / KCH * KCH;
    float* Ginv = reinterpret_cast<float*>(ws + ws_off_ginv(n_pad));
    double* W = reinterpret_cast<double*>(ws + ws_off_w(n_pad));
    float* Z = reinterpret_cast<float*>(ws + ws_off_z(n_pad));
    k_als_inverse<<<1, 1024, 0, stream>>>(G, a->reg, a->d, W, Ginv);
    const size_t zsmem = ((size_t)ALS_D * ALS_D + 32 * ALS_D) * 4;
    CF_CUDA_OK(cudaFuncSetAttribute(k_als_z, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsmem));
    long long zg = (a->n_y + 31) / 32;
    if (zg > (long long)cf_num_sms() * 2) zg = (long long)cf_num_sms() * 2;
    k_als_z<<<(unsigned)zg, 256, zsmem, stream>>>(a->Y, a->n_y, a->d, a->ldy, Ginv, Z);
    S.skip_le = WB_N; S.Z = Z;
    const size_t wsmem = ((size_t)2 * WB_N * (ALS_D + 1) + WB_N * (WB_N + 1) + ALS_D + WB_N) * 4;
    CF_CUDA_OK(cudaFuncSetAttribute(k_als_woodbury, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
    long long wg = a->n_x;
    if (wg > (long long)cf_num_sms() * 16) wg = (long long)cf_num_sms() * 16;
    k_als_woodbury<<<(unsigned)wg, 128, wsmem, stream>>>(S);
  }
  long long sg = a->n_x;
  const long long cap = (long long)cf_num_sms() * 8;
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
