// Exact full-catalog scoring + masked streaming top-K, top-K merge and ranking metrics (sm_100a, CUDA cores).
//
// Replaces (reference src/models/pl/models/bprmf.py:77-103 and the same lines of cml.py / gbprmf.py / wrmf.py):
//   predicts = matmul(U[test_users], V^T) (+ bias | -squared distance)       [T, n_items] in host-visible memory
//   top_k(predicts, max|train(u)| + topN) then a Python loop dropping training items
// with one pass that never materialises the score matrix: a block owns one query user, streams over the items,
// masks the user's training items with a per-chunk shared-memory bitmap built from its sorted CSR row, and keeps
// candidates that beat the running K-th best in a shared buffer that is bitonic-sorted and truncated when full.
// Order: (score desc, item id asc) == tf.nn.top_k's tie rule.  Scores: fp32 inputs, fp64 sequential-k accumulation
// (bit-identical to oracle/scoring.py, independent of FMA contraction).
#include <math.h>

#include "common.cuh"

namespace {

constexpr int TK_THREADS = 256;
constexpr int TK_CAP = 2048;    // candidate buffer entries (K <= 1024)
constexpr int TK_CHUNK = 8192;  // items per mask bitmap

__device__ __forceinline__ bool key_before(double va, int ia, double vb, int ib) {
  return va > vb || (va == vb && ia < ib);
}

// in-place bitonic sort of (val, idx)[0..n) (n a power of two) by (val desc, idx asc); all threads of the block call it
__device__ void bitonic_sort(double* val, int* idx, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int p = t ^ j;
        if (p > t) {
          const bool up = (t & k) == 0;
          const double va = val[t], vb = val[p];
          const int ia = idx[t], ib = idx[p];
          const bool swap = up ? key_before(vb, ib, va, ia) : key_before(va, ia, vb, ib);
          if (swap) {
            val[t] = vb; idx[t] = ib;
            val[p] = va; idx[p] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
}

struct TopkDev {
  const float *U, *V, *b;
  long long n_items;
  int ld, nvec;
  const int32_t* users;
  int T, K, kind;
  const long long* tr_indptr;
  const int32_t* tr_indices;
  int32_t* out_idx;
  double* out_val;
  long long item_lo, item_hi;
  const int32_t* only_if_flag;   // when set, only query rows t with only_if_flag[t] != 0 are computed
};


// fp32 inputs, fp64 sequential-k accumulation (bit-identical to oracle/scoring.py)
__device__ __forceinline__ double score_item(const TopkDev& P, const float* s_u, long long item) {
  const float4* vp = reinterpret_cast<const float4*>(P.V + item * P.ld);
  double s = 0.0;
  if (P.kind == CF_SCORE_NEG_SQDIST) {  // cml.py:116
    for (int k4 = 0; k4 < P.nvec; ++k4) {
      const float4 v = __ldg(vp + k4);
      const float4 q = *reinterpret_cast<const float4*>(s_u + 4 * k4);
      double df = (double)q.x - (double)v.x; s = __dadd_rn(s, __dmul_rn(df, df));
      df = (double)q.y - (double)v.y; s = __dadd_rn(s, __dmul_rn(df, df));
      df = (double)q.z - (double)v.z; s = __dadd_rn(s, __dmul_rn(df, df));
      df = (double)q.w - (double)v.w; s = __dadd_rn(s, __dmul_rn(df, df));
    }
    s = -s;
  } else {  // bprmf.py:80 / gbprmf.py:98 / wrmf.py:80
    for (int k4 = 0; k4 < P.nvec; ++k4) {
      const float4 v = __ldg(vp + k4);
      const float4 q = *reinterpret_cast<const float4*>(s_u + 4 * k4);
      s = fma((double)q.x, (double)v.x, s);
      s = fma((double)q.y, (double)v.y, s);
      s = fma((double)q.z, (double)v.z, s);
      s = fma((double)q.w, (double)v.w, s);
    }
    if (P.kind == CF_SCORE_DOT_BIAS) s = __dadd_rn(s, (double)__ldg(P.b + item));
  }
  return s;
}

// the reference's __predict__ as a dense [T, n_items] fp64 matrix (small inputs / debugging; top-K never uses it)
__global__ void __launch_bounds__(TK_THREADS) k_scores(const __grid_constant__ TopkDev P, double* __restrict__ out) {
  __shared__ __align__(16) float s_u[512];
  for (int t = blockIdx.x; t < P.T; t += gridDim.x) {
    const long long u = P.users ? P.users[t] : t;
    for (int k = threadIdx.x; k < P.ld; k += blockDim.x) s_u[k] = P.U[u * P.ld + k];
    __syncthreads();
    for (long long item = threadIdx.x; item < P.n_items; item += blockDim.x)
      out[(long long)t * P.n_items + item] = score_item(P, s_u, item);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(TK_THREADS) k_topk_exact(const __grid_constant__ TopkDev P) {
  __shared__ double s_val[TK_CAP];
  __shared__ int s_idx[TK_CAP];
  __shared__ __align__(16) float s_u[512];
  __shared__ unsigned s_mask[TK_CHUNK / 32];
  __shared__ int s_count;
  __shared__ double s_thv;
  __shared__ int s_thi;

  for (int t = blockIdx.x; t < P.T; t += gridDim.x) {
    if (P.only_if_flag && !P.only_if_flag[t]) continue;   // block-uniform
    const long long u = P.users ? P.users[t] : t;
    for (int k = threadIdx.x; k < P.ld; k += blockDim.x) s_u[k] = P.U[u * P.ld + k];
    if (threadIdx.x == 0) {
      s_count = 0;
      s_thv = -INFINITY;
      s_thi = 0x7fffffff;
    }
    long long tlo = 0, thi = 0;
    if (P.tr_indptr) {
      tlo = P.tr_indptr[u];
      thi = P.tr_indptr[u + 1];
    }
    __syncthreads();

    for (long long c0 = P.item_lo; c0 < P.item_hi; c0 += TK_CHUNK) {
      const long long c1 = min(c0 + (long long)TK_CHUNK, P.item_hi);
      for (int w = threadIdx.x; w < TK_CHUNK / 32; w += blockDim.x) s_mask[w] = 0u;
      __syncthreads();
      if (thi > tlo) {  // mark this user's training items that fall into [c0, c1)
        long long lo = tlo, hi = thi;
        while (lo < hi) {  // first entry >= c0
          const long long mid = (lo + hi) >> 1;
          if (P.tr_indices[mid] < c0) lo = mid + 1; else hi = mid;
        }
        for (long long e = lo + threadIdx.x; e < thi; e += blockDim.x) {
          const long long x = P.tr_indices[e];
          if (x >= c1) break;
          atomicOr(&s_mask[(x - c0) >> 5], 1u << ((x - c0) & 31));
        }
      }
      __syncthreads();
      for (long long s0 = c0; s0 < c1; s0 += TK_THREADS) {
        const long long item = s0 + threadIdx.x;
        if (item < c1 && !((s_mask[(item - c0) >> 5] >> ((item - c0) & 31)) & 1u)) {
          const double s = score_item(P, s_u, item);
          if (key_before(s, (int)item, s_thv, s_thi)) {
            const int pos = atomicAdd(&s_count, 1);
            s_val[pos] = s;
            s_idx[pos] = (int)item;
          }
        }
        __syncthreads();
        if (s_count > TK_CAP - TK_THREADS) {  // block-uniform: compact to the K best, raise the threshold
          const int cnt = s_count;
          for (int e = cnt + threadIdx.x; e < TK_CAP; e += blockDim.x) {
            s_val[e] = -INFINITY;
            s_idx[e] = 0x7fffffff;
          }
          __syncthreads();
          bitonic_sort(s_val, s_idx, TK_CAP);
          if (threadIdx.x == 0) {
            s_count = min(cnt, P.K);
            if (cnt >= P.K) {
              s_thv = s_val[P.K - 1];
              s_thi = s_idx[P.K - 1];
            }
          }
          __syncthreads();
        }
      }
    }
    const int cnt = s_count;
    for (int e = cnt + threadIdx.x; e < TK_CAP; e += blockDim.x) {
      s_val[e] = -INFINITY;
      s_idx[e] = 0x7fffffff;
    }
    __syncthreads();
    bitonic_sort(s_val, s_idx, TK_CAP);
    for (int k = threadIdx.x; k < P.K; k += blockDim.x) {
      const bool ok = k < cnt;
      P.out_idx[(long long)t * P.K + k] = ok ? s_idx[k] : -1;
      if (P.out_val) P.out_val[(long long)t * P.K + k] = ok ? s_val[k] : -INFINITY;
    }
    __syncthreads();
  }
}

// merge P shard lists per user: dynamic smem holds n2 = pow2 >= P*K entries
__global__ void __launch_bounds__(TK_THREADS) k_topk_merge(const int32_t* idx, const double* val, int P, int T, int K, int n2,
                                                           int32_t* out_idx, double* out_val) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* s_val = reinterpret_cast<double*>(smem);
  int* s_idx = reinterpret_cast<int*>(s_val + n2);
  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    for (int e = threadIdx.x; e < n2; e += blockDim.x) {
      double v = -INFINITY;
      int i = 0x7fffffff;
      if (e < P * K) {
        const int p = e / K, k = e - p * K;
        const long long off = ((long long)p * T + t) * K + k;
        const int ii = idx[off];
        if (ii >= 0) {
          i = ii;
          v = val[off];
        }
      }
      s_val[e] = v;
      s_idx[e] = i;
    }
    __syncthreads();
    bitonic_sort(s_val, s_idx, n2);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const bool ok = s_idx[k] != 0x7fffffff;
      out_idx[(long long)t * K + k] = ok ? s_idx[k] : -1;
      if (out_val) out_val[(long long)t * K + k] = ok ? s_val[k] : -INFINITY;
    }
    __syncthreads();
  }
}

// metrics/ranking.py:11-91, one thread per user; out[t, 8] = {pre, recall, ndcg, map, mrr, hit, rr, n_pred}
__global__ void __launch_bounds__(128) k_rank_metrics(const int32_t* __restrict__ pred, int T, int ldp, int k,
                                                      const long long* __restrict__ tptr, const int32_t* __restrict__ tidx,
                                                      double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int32_t* p = pred + (long long)t * ldp;
  const long long lo = tptr[t], hi = tptr[t + 1];
  const int n_true = (int)(hi - lo);
  const int kk = min(k, ldp);
  int n_pred = 0, n_hit_pos = 0, n_common = 0, first_hit = -1, truth0_pos = -1;
  double dcg = 0.0, ap = 0.0;
  const int32_t truth0 = n_true > 0 ? tidx[lo] : -1;
  for (int i = 0; i < kk; ++i) {
    const int32_t x = p[i];
    if (x < 0) break;
    ++n_pred;
    if (x == truth0 && truth0_pos < 0) truth0_pos = i;
    if (csr_contains(tidx, lo, hi, x)) {
      ++n_hit_pos;                                   // label_list[i] = True           (ranking.py:34)
      if (first_hit < 0) first_hit = i;
      dcg += 1.0 / log2((double)i + 2.0);            // (2**1 - 1) / log2(i + 2)       (:36)
      ap += (double)n_hit_pos / ((double)i + 1.0);   // rank / (i + 1)                 (:51-52)
      bool dup = false;                              // set(pred[:k]) & truth          (:16,:25)
      for (int q = 0; q < i; ++q) dup |= (p[q] == x);
      n_common += !dup;
    }
  }
  double ideal = 0.0;
  for (int i = 0; i < n_hit_pos; ++i) ideal += 1.0 / log2((double)i + 2.0);  // sorted(label_list, reverse=True)   (:37-38)
  double* o = out + (long long)t * 8;
  o[0] = (double)n_common / (double)k;
  o[1] = (double)n_common / fmax((double)n_true, 1.0);
  o[2] = dcg / fmax(ideal, 1.0);
  o[3] = n_true > 0 ? ap / (double)n_true : 0.0;
  o[4] = first_hit >= 0 ? 1.0 / ((double)first_hit + 1.0) : 0.0;
  o[5] = truth0_pos >= 0 ? 1.0 : 0.0;                                        // hr  (:80)
  o[6] = truth0_pos >= 0 ? 1.0 / ((double)truth0_pos + 1.0) : 0.0;           // arhr (:89-90)
  o[7] = (double)n_pred;
}

}  // namespace

int cf_topk_exact_flagged(const cf_topk_args* a, const int32_t* only_if_flag, cudaStream_t stream) {
  CF_CHECK_ARG(a != nullptr, "cf_topk_exact: args is NULL");
  CF_CHECK_ARG(a->U && a->V && a->out_idx, "cf_topk_exact: U, V and out_idx are required");
  CF_CHECK_ARG(a->d > 0 && a->ld >= a->d && a->ld % 4 == 0 && a->ld <= 512, "cf_topk_exact: need 0 < d <= ld <= 512, ld %% 4 == 0");
  CF_CHECK_ARG(a->T > 0, "cf_topk_exact: T must be positive");
  CF_CHECK_ARG(a->K > 0 && a->K <= TK_CAP / 2, "cf_topk_exact: K must be in [1, %d] (got %d)", TK_CAP / 2, a->K);
  CF_CHECK_ARG(a->kind >= CF_SCORE_DOT && a->kind <= CF_SCORE_NEG_SQDIST, "cf_topk_exact: unknown scoring kind %d", a->kind);
  CF_CHECK_ARG(a->kind != CF_SCORE_DOT_BIAS || a->b, "cf_topk_exact: DOT_BIAS needs the bias vector");
  CF_CHECK_ARG(a->n_items > 0 && a->n_items < (1ll << 31), "cf_topk_exact: n_items must fit int32");
  TopkDev P;
  P.U = a->U; P.V = a->V; P.b = a->b; P.n_items = a->n_items; P.ld = a->ld; P.nvec = a->ld / 4;
  P.users = a->users; P.T = a->T; P.K = a->K; P.kind = a->kind;
  P.tr_indptr = (const long long*)a->train.indptr; P.tr_indices = a->train.indices;
  P.out_idx = a->out_idx; P.out_val = a->out_val;
  P.item_lo = a->item_lo; P.item_hi = a->item_hi; P.only_if_flag = only_if_flag;
  if (P.item_lo == 0 && P.item_hi == 0) P.item_hi = a->n_items;
  CF_CHECK_ARG(P.item_lo >= 0 && P.item_hi <= a->n_items && P.item_lo <= P.item_hi, "cf_topk_exact: bad item range");
  int grid = a->T;
  const int cap = cf_num_sms() * 4;
  if (grid > cap) grid = cap;
  k_topk_exact<<<grid, TK_THREADS, 0, stream>>>(P);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cf_topk_exact(const cf_topk_args* a, void* stream_) {
  return cf_topk_exact_flagged(a, nullptr, (cudaStream_t)stream_);
}

extern "C" int cf_scores(const cf_topk_args* a, double* out_scores, void* stream_) {
  CF_CHECK_ARG(a != nullptr && out_scores != nullptr, "cf_scores: NULL argument");
  CF_CHECK_ARG(a->U && a->V, "cf_scores: U and V are required");
  CF_CHECK_ARG(a->d > 0 && a->ld >= a->d && a->ld % 4 == 0 && a->ld <= 512, "cf_scores: need 0 < d <= ld <= 512, ld %% 4 == 0");
  CF_CHECK_ARG(a->T > 0 && a->n_items > 0, "cf_scores: T and n_items must be positive");
  CF_CHECK_ARG(a->kind >= CF_SCORE_DOT && a->kind <= CF_SCORE_NEG_SQDIST, "cf_scores: unknown scoring kind %d", a->kind);
  CF_CHECK_ARG(a->kind != CF_SCORE_DOT_BIAS || a->b, "cf_scores: DOT_BIAS needs the bias vector");
  TopkDev P;
  P.U = a->U; P.V = a->V; P.b = a->b; P.n_items = a->n_items; P.ld = a->ld; P.nvec = a->ld / 4;
  P.users = a->users; P.T = a->T; P.K = 0; P.kind = a->kind;
  P.tr_indptr = nullptr; P.tr_indices = nullptr; P.out_idx = nullptr; P.out_val = nullptr;
  P.item_lo = 0; P.item_hi = a->n_items; P.only_if_flag = nullptr;
  int grid = a->T;
  const int cap = cf_num_sms() * 4;
  if (grid > cap) grid = cap;
  k_scores<<<grid, TK_THREADS, 0, (cudaStream_t)stream_>>>(P, out_scores);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cf_topk_merge(const int32_t* idx, const double* val, int32_t P, int32_t T, int32_t K, int32_t* out_idx,
                             double* out_val, void* stream_) {
  CF_CHECK_ARG(idx && val && out_idx && P > 0 && T > 0 && K > 0, "cf_topk_merge: bad arguments");
  int n2 = 1;
  while (n2 < P * K) n2 <<= 1;
  if (n2 < 32) n2 = 32;
  const size_t smem = (size_t)n2 * (sizeof(double) + sizeof(int));
  CF_CHECK_ARG(smem <= 200 * 1024, "cf_topk_merge: P*K = %d too large for one block", P * K);
  CF_CUDA_OK(cudaFuncSetAttribute(k_topk_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = T;
  const int cap = cf_num_sms() * 4;
  if (grid > cap) grid = cap;
  k_topk_merge<<<grid, TK_THREADS, smem, (cudaStream_t)stream_>>>(idx, val, P, T, K, n2, out_idx, out_val);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cf_rank_metrics(const int32_t* pred, int32_t T, int32_t ldp, int32_t k, const int64_t* truth_indptr,
                               const int32_t* truth_indices, double* out, void* stream_) {
  CF_CHECK_ARG(pred && truth_indptr && truth_indices && out, "cf_rank_metrics: NULL argument");
  CF_CHECK_ARG(T > 0 && ldp > 0 && k > 0, "cf_rank_metrics: len(yss_true) != len(yss_pred) or len(yss_true)==0 or k<=0!");
  k_rank_metrics<<<(T + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(pred, T, ldp, k, (const long long*)truth_indptr,
                                                                      truth_indices, out);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}
