// Neighbourhood models and the user-similarity preprocessing of the pairwise family.
//
// Replaces the numpy loops of the reference's
//   models/basic/models/itemcf.py:19-40  (__calsim__: R^T R, cosine normalisation; __topk__: K best neighbours per item)
//   models/basic/models/usercf.py:19-42  (__calsim__: R R^T; __predict__: the K most similar users' rows, weighted)
//   models/pl/models/prigp.py:60-81, cplr_u.py:64-87 (the same user-user cosine + top-K before their samplers)
// and the scoring / recommendation loops itemcf.py:42-66, usercf.py:31-67.
//
// cf_neighbors       one block per entity a: co-occurrence counts with every entity sharing a feature (sparse R^T R row in a
//                    dense L2-resident accumulator), sim = (c / |lo|) / |hi| in IEEE fp32 with lo < hi the two indices --
//                    the order the reference's in-place row / column divisions produce (itemcf.py:21-26) -- zero diagonal,
//                    then the K best by (similarity desc, index order given by the caller).
// cf_neighbor_scores one block per query user: the reference's accumulation of fp32 products into a float64 vector.  Every
//                    addend is an fp32 value and the exponents span a few bits, so the fp64 sums are EXACT and do not
//                    depend on the order of the atomics.
// cf_topk_dense      masked top-N of a dense float64 score row (np.argsort(...)[-maxsz-topN:][::-1] + the filter loop).
#include <math.h>

#include "common.cuh"

namespace {

struct NbrDev {
  const long long* r_indptr; const int32_t* r_indices; const float* r_values;   // entity -> features
  const long long* c_indptr; const int32_t* c_indices; const float* c_values;   // feature -> entities
  long long n;
  int K, tie_high;
  int32_t* out_idx; float* out_sim;
  float* norms;
  float* acc;        // [grid, n] zero at rest
  float* csim;       // [grid, n]
  int32_t* cand;     // [grid, n]
};

__global__ void __launch_bounds__(256) k_row_norms(const __grid_constant__ NbrDev P) {
  for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a < P.n; a += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (long long e = P.r_indptr[a]; e < P.r_indptr[a + 1]; ++e) {
      const float v = P.r_values ? P.r_values[e] : 1.f;
      s = __fadd_rn(s, __fmul_rn(v, v));
    }
    P.norms[a] = sqrtf(s);      // np.linalg.norm of the float32 row / column (exact for a binarised matrix)
  }
}

// (similarity, index) ordering of the selection: larger similarity first; among equals the higher (or lower) index
__device__ __forceinline__ bool nbr_before(float sa, int ia, float sb, int ib, int tie_high) {
  return sa > sb || (sa == sb && (tie_high ? ia > ib : ia < ib));
}

__global__ void __launch_bounds__(256) k_neighbors(const __grid_constant__ NbrDev P) {
  __shared__ int s_n;
  __shared__ float s_bs[256];
  __shared__ int s_bi[256], s_bp[256];
  float* acc = P.acc + (long long)blockIdx.x * P.n;
  float* csim = P.csim + (long long)blockIdx.x * P.n;
  int32_t* cand = P.cand + (long long)blockIdx.x * P.n;
  for (long long a = blockIdx.x; a < P.n; a += gridDim.x) {
    const long long lo = P.r_indptr[a], hi = P.r_indptr[a + 1];
    // ---- co-occurrence counts of a with every entity that shares a feature
    for (long long e = lo; e < hi; ++e) {
      const int f = P.r_indices[e];
      const float va = P.r_values ? P.r_values[e] : 1.f;
      for (long long e2 = P.c_indptr[f] + threadIdx.x; e2 < P.c_indptr[f + 1]; e2 += blockDim.x)
        atomicAdd(acc + P.c_indices[e2], va * (P.c_values ? P.c_values[e2] : 1.f));
    }
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // ---- similarities of the touched entities (each taken once: the exchange returns the accumulator to zero)
    const float da = P.norms[a];
    for (long long e = lo; e < hi; ++e) {
      const int f = P.r_indices[e];
      for (long long e2 = P.c_indptr[f] + threadIdx.x; e2 < P.c_indptr[f + 1]; e2 += blockDim.x) {
        const int b = P.c_indices[e2];
        const float c = atomicExch(acc + b, 0.f);
        if (c != 0.f && b != a) {       // the diagonal is set to zero (itemcf.py:27)
          const float db = P.norms[b];
          const float s = b > a ? __fdiv_rn(__fdiv_rn(c, da), db) : __fdiv_rn(__fdiv_rn(c, db), da);
          if (s != 0.f) {
            const int pos = atomicAdd(&s_n, 1);
            cand[pos] = b;
            csim[pos] = s;
          }
        }
      }
    }
    __syncthreads();
    const int nc = s_n;
    // ---- the K best: K rounds of a block-wide arg-best over the candidates that are still in
    for (int k = 0; k < P.K; ++k) {
      float bs = -INFINITY;
      int bi = -1, bp = -1;
      for (int p = threadIdx.x; p < nc; p += blockDim.x) {
        const float s = csim[p];
        if (s > 0.f || s < 0.f) {   // (taken entries are marked with 0)
          const int i = cand[p];
          if (bp < 0 || nbr_before(s, i, bs, bi, P.tie_high)) { bs = s; bi = i; bp = p; }
        }
      }
      s_bs[threadIdx.x] = bs; s_bi[threadIdx.x] = bi; s_bp[threadIdx.x] = bp;
      __syncthreads();
      for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
          const int q = threadIdx.x + o;
          if (s_bp[q] >= 0 && (s_bp[threadIdx.x] < 0 || nbr_before(s_bs[q], s_bi[q], s_bs[threadIdx.x], s_bi[threadIdx.x], P.tie_high))) {
            s_bs[threadIdx.x] = s_bs[q]; s_bi[threadIdx.x] = s_bi[q]; s_bp[threadIdx.x] = s_bp[q];
          }
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        const bool ok = s_bp[0] >= 0;
        P.out_idx[a * P.K + k] = ok ? s_bi[0] : -1;
        P.out_sim[a * P.K + k] = ok ? s_bs[0] : 0.f;
        if (ok) csim[s_bp[0]] = 0.f;
      }
      __syncthreads();
    }
  }
}

struct ScoreDev {
  const int32_t* users; int T;
  const long long* t_indptr; const int32_t* t_indices; const float* t_values;   // training CSR user -> items
  const int32_t* nbr_idx; const float* nbr_sim; int K;
  long long n_items;
  int mode;            // 0: item neighbourhoods (itemcf.py:42-50), 1: user neighbourhoods (usercf.py:31-44)
  double* out;         // [T, n_items] zeroed by the caller
};

__global__ void __launch_bounds__(256) k_neighbor_scores(const __grid_constant__ ScoreDev P) {
  for (int t = blockIdx.x; t < P.T; t += gridDim.x) {
    const long long u = P.users ? P.users[t] : t;
    double* out = P.out + (long long)t * P.n_items;
    if (P.mode == 0) {
      const long long lo = P.t_indptr[u], hi = P.t_indptr[u + 1];
      const long long work = (hi - lo) * P.K;
      for (long long w = threadIdx.x; w < work; w += blockDim.x) {
        const long long e = lo + w / P.K;
        const int k = (int)(w % P.K);
        const long long i = P.t_indices[e];
        const int j = P.nbr_idx[i * P.K + k];
        if (j >= 0) atomicAdd(out + j, (double)__fmul_rn(P.nbr_sim[i * P.K + k], P.t_values ? P.t_values[e] : 1.f));
      }
    } else {
      for (int k = 0; k < P.K; ++k) {
        const int v = P.nbr_idx[u * P.K + k];
        const float s = P.nbr_sim[u * P.K + k];
        if (v < 0 || !(s > 0.f)) continue;       // usercf.py:40: only positive similarities count
        for (long long e = P.t_indptr[v] + threadIdx.x; e < P.t_indptr[v + 1]; e += blockDim.x)
          atomicAdd(out + P.t_indices[e], (double)__fmul_rn(s, P.t_values ? P.t_values[e] : 1.f));
      }
    }
  }
}

struct DenseTopDev {
  double* scores; long long n_items; int T, N, tie_high;
  const int32_t* users;
  const long long* m_indptr; const int32_t* m_indices;
  int32_t* out_idx; double* out_val;
};

__device__ __forceinline__ bool dense_before(double sa, int ia, double sb, int ib, int tie_high) {
  return sa > sb || (sa == sb && (tie_high ? ia > ib : ia < ib));
}

__global__ void __launch_bounds__(256) k_topk_dense(const __grid_constant__ DenseTopDev P) {
  __shared__ double s_v[256];
  __shared__ int s_i[256];
  for (int t = blockIdx.x; t < P.T; t += gridDim.x) {
    double* sc = P.scores + (long long)t * P.n_items;
    if (P.m_indptr) {       // the user's training items never appear (itemcf.py:61-62)
      const long long u = P.users ? P.users[t] : t;
      for (long long e = P.m_indptr[u] + threadIdx.x; e < P.m_indptr[u + 1]; e += blockDim.x) sc[P.m_indices[e]] = -INFINITY;
    }
    __syncthreads();
    for (int k = 0; k < P.N; ++k) {
      double bv = -INFINITY;
      int bi = -1;
      for (long long j = threadIdx.x; j < P.n_items; j += blockDim.x) {
        const double v = sc[j];
        if (v > -INFINITY && (bi < 0 || dense_before(v, (int)j, bv, bi, P.tie_high))) { bv = v; bi = (int)j; }
      }
      s_v[threadIdx.x] = bv; s_i[threadIdx.x] = bi;
      __syncthreads();
      for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
          const int q = threadIdx.x + o;
          if (s_i[q] >= 0 && (s_i[threadIdx.x] < 0 || dense_before(s_v[q], s_i[q], s_v[threadIdx.x], s_i[threadIdx.x], P.tie_high))) {
            s_v[threadIdx.x] = s_v[q]; s_i[threadIdx.x] = s_i[q];
          }
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        P.out_idx[(long long)t * P.N + k] = s_i[0];
        if (P.out_val) P.out_val[(long long)t * P.N + k] = s_i[0] >= 0 ? s_v[0] : -INFINITY;
        if (s_i[0] >= 0) sc[s_i[0]] = -INFINITY;
      }
      __syncthreads();
    }
  }
}

}  // namespace

extern "C" int64_t cf_neighbors_concurrent_rows(void) { return (int64_t)cf_num_sms() * 2; }

extern "C" int cf_neighbors(const cf_neighbor_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CF_CHECK_ARG(a != nullptr, "cf_neighbors: args is NULL");
  CF_CHECK_ARG(a->rows.indptr && a->rows.indices && a->cols.indptr && a->cols.indices, "cf_neighbors: both CSRs are required");
  CF_CHECK_ARG(a->rows.n_rows > 0 && a->rows.n_rows == a->cols.n_cols && a->rows.n_cols == a->cols.n_rows && a->rows.nnz == a->cols.nnz,
               "cf_neighbors: cols must be the transpose of rows");
  CF_CHECK_ARG(a->K > 0 && a->K <= 4096, "cf_neighbors: K must be in [1, 4096]");
  CF_CHECK_ARG(a->out_idx && a->out_sim && a->norms && a->scratch && a->cand, "cf_neighbors: NULL output / scratch");
  CF_CHECK_ARG(a->grid_rows > 0, "cf_neighbors: grid_rows must be positive");
  CF_CHECK_ARG((a->rows.values == nullptr) == (a->cols.values == nullptr), "cf_neighbors: values on both CSRs or on neither");
  NbrDev P = {};
  P.r_indptr = (const long long*)a->rows.indptr; P.r_indices = a->rows.indices; P.r_values = a->rows.values;
  P.c_indptr = (const long long*)a->cols.indptr; P.c_indices = a->cols.indices; P.c_values = a->cols.values;
  P.n = a->rows.n_rows; P.K = a->K; P.tie_high = a->tie_high_index_first;
  P.out_idx = a->out_idx; P.out_sim = a->out_sim; P.norms = a->norms;
  P.acc = a->scratch; P.csim = a->scratch + a->grid_rows * P.n; P.cand = a->cand;
  long long g = (P.n + 255) / 256;
  if (g > 1024) g = 1024;
  k_row_norms<<<(unsigned)g, 256, 0, stream>>>(P);
  long long grid = a->grid_rows < P.n ? a->grid_rows : P.n;
  k_neighbors<<<(unsigned)grid, 256, 0, stream>>>(P);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cf_neighbor_scores(const cf_neighbor_score_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CF_CHECK_ARG(a != nullptr, "cf_neighbor_scores: args is NULL");
  CF_CHECK_ARG(a->train.indptr && a->train.indices && a->nbr_idx && a->nbr_sim && a->out_scores, "cf_neighbor_scores: NULL pointer");
  CF_CHECK_ARG(a->T > 0 && a->K > 0 && a->train.n_cols > 0, "cf_neighbor_scores: T, K and n_items must be positive");
  CF_CHECK_ARG(a->mode == 0 || a->mode == 1, "cf_neighbor_scores: mode must be 0 (item neighbourhoods) or 1 (user neighbourhoods)");
  ScoreDev P = {};
  P.users = a->users; P.T = a->T;
  P.t_indptr = (const long long*)a->train.indptr; P.t_indices = a->train.indices; P.t_values = a->train.values;
  P.nbr_idx = a->nbr_idx; P.nbr_sim = a->nbr_sim; P.K = a->K; P.n_items = a->train.n_cols; P.mode = a->mode; P.out = a->out_scores;
  int grid = a->T < cf_num_sms() * 8 ? a->T : cf_num_sms() * 8;
  k_neighbor_scores<<<grid, 256, 0, stream>>>(P);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cf_topk_dense(double* scores, int64_t n_items, int32_t T, int32_t N, int32_t tie_high_index_first,
                             const int32_t* users, const cf_csr* mask, int32_t* out_idx, double* out_val, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CF_CHECK_ARG(scores && out_idx && n_items > 0 && T > 0 && N > 0, "cf_topk_dense: scores, out_idx and positive sizes are required");
  DenseTopDev P = {};
  P.scores = scores; P.n_items = n_items; P.T = T; P.N = N; P.tie_high = tie_high_index_first; P.users = users;
  P.m_indptr = mask ? (const long long*)mask->indptr : nullptr; P.m_indices = mask ? mask->indices : nullptr;
  P.out_idx = out_idx; P.out_val = out_val;
  int grid = T < cf_num_sms() * 8 ? T : cf_num_sms() * 8;
  k_topk_dense<<<grid, 256, 0, stream>>>(P);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}
