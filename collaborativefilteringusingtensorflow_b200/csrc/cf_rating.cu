// Rating path: predictions of explicit (user, item) pairs and the rating metrics.
//
// Replaces, for reference src/models/basic/models/mf.py:
//   :66-72  __predict          rating_ = reduce_sum(user_embed * items_embed, 1) for the fed (user, item) rows
//   :80-83  __eval             clip_by_value(predict, range) -> metrics/rating.py evaluate
// and reference src/metrics/rating.py:4-17 (MAE / MSE / RMSE as sums over the test tuples; the caller divides by n and
// takes the root exactly where rating.py does).
// The dot product uses the same fp64 sequential-k accumulation of the fp32 rows as the scoring kernels (cf_topk.cu), so
// predict_pairs(u, i) is the float32 rounding of cf_scores' entry (u, i).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_predict_pairs(const float* __restrict__ U, const float* __restrict__ V,
                                                       const float* __restrict__ b, long long n_users, long long n_items,
                                                       int ld, int kind, const int32_t* __restrict__ pairs, long long n,
                                                       float* __restrict__ out, int32_t* counters) {
  const int nvec = ld / 4;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    const long long u = __ldg(pairs + 2 * k), i = __ldg(pairs + 2 * k + 1);
    if (u < 0 || u >= n_users || i < 0 || i >= n_items) {   // TF: InvalidArgumentError in the gather
      atomicOr(counters + 1, CF_FLAG_INDEX_RANGE);
      out[k] = nanf("");
      continue;
    }
    const float4* up = reinterpret_cast<const float4*>(U + u * ld);
    const float4* vp = reinterpret_cast<const float4*>(V + i * ld);
    double s = 0.0;
    if (kind == CF_SCORE_NEG_SQDIST) {
      for (int k4 = 0; k4 < nvec; ++k4) {
        const float4 q = __ldg(up + k4), v = __ldg(vp + k4);
        double df = (double)q.x - (double)v.x; s = __dadd_rn(s, __dmul_rn(df, df));
        df = (double)q.y - (double)v.y; s = __dadd_rn(s, __dmul_rn(df, df));
        df = (double)q.z - (double)v.z; s = __dadd_rn(s, __dmul_rn(df, df));
        df = (double)q.w - (double)v.w; s = __dadd_rn(s, __dmul_rn(df, df));
      }
      s = -s;
    } else {
      for (int k4 = 0; k4 < nvec; ++k4) {
        const float4 q = __ldg(up + k4), v = __ldg(vp + k4);
        s = fma((double)q.x, (double)v.x, s);
        s = fma((double)q.y, (double)v.y, s);
        s = fma((double)q.z, (double)v.z, s);
        s = fma((double)q.w, (double)v.w, s);
      }
      if (kind == CF_SCORE_DOT_BIAS) s = __dadd_rn(s, (double)__ldg(b + i));
    }
    out[k] = (float)s;
  }
}

// sums[0] += sum |t - clip(p)|, sums[1] += sum (t - clip(p))^2 in fp64 (rating.py:4-17 computes in float64)
template <typename PredT>
__global__ void __launch_bounds__(256) k_rating_metrics(const PredT* __restrict__ pred, const double* __restrict__ truth,
                                                        long long n, double lo, double hi, double* sums) {
  double a = 0.0, q = 0.0;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    double p = (double)pred[k];
    p = p < lo ? lo : (p > hi ? hi : p);   // tf.clip_by_value (mf.py:81); NaN stays NaN, like TF
    const double e = truth[k] - p;
    a += fabs(e);
    q += e * e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  __shared__ double sa[8], sq[8];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sa[w] = a; sq[w] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tq = 0.0;
    for (int k = 0; k < 8; ++k) { ta += sa[k]; tq += sq[k]; }
    atomicAdd(sums, ta);
    atomicAdd(sums + 1, tq);
  }
}

}  // namespace

extern "C" int cf_predict_pairs(const float* U, const float* V, const float* b, int64_t n_users, int64_t n_items, int32_t d,
                                int32_t ld, int32_t score, const int32_t* pairs, int64_t n, float* out, int32_t* counters,
                                void* stream_) {
  CF_CHECK_ARG(U && V && pairs && out && counters, "cf_predict_pairs: NULL pointer");
  CF_CHECK_ARG(d > 0 && ld >= d && ld % 4 == 0, "cf_predict_pairs: ld must be a multiple of 4 and >= d (ld=%d, d=%d)", ld, d);
  CF_CHECK_ARG(((uintptr_t)U % 16 == 0) && ((uintptr_t)V % 16 == 0), "cf_predict_pairs: tables must be 16-byte aligned");
  CF_CHECK_ARG(score == CF_SCORE_DOT || score == CF_SCORE_NEG_SQDIST || (score == CF_SCORE_DOT_BIAS && b != nullptr),
               "cf_predict_pairs: bad score kind %d (DOT_BIAS needs b)", score);
  CF_CHECK_ARG(n_users > 0 && n_items > 0 && n >= 0, "cf_predict_pairs: bad sizes");
  if (n == 0) return 0;
  long long grid = (n + 255) / 256;
  const long long cap = (long long)cf_num_sms() * 8;
  if (grid > cap) grid = cap;
  k_predict_pairs<<<(unsigned)grid, 256, 0, (cudaStream_t)stream_>>>(U, V, b, n_users, n_items, ld, score, pairs, n, out, counters);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cf_rating_metrics(const void* pred, int32_t pred_is_f64, const double* truth, int64_t n, double lo, double hi,
                                 double* sums, void* stream_) {
  CF_CHECK_ARG(pred && truth && sums, "cf_rating_metrics: NULL pointer");
  CF_CHECK_ARG(n > 0, "cf_rating_metrics: no ratings (rating.py divides by ys_true.shape[0])");
  CF_CHECK_ARG(lo <= hi, "cf_rating_metrics: empty clip range");
  long long grid = (n + 255) / 256;
  const long long cap = (long long)cf_num_sms() * 4;
  if (grid > cap) grid = cap;
  if (pred_is_f64)
    k_rating_metrics<double><<<(unsigned)grid, 256, 0, (cudaStream_t)stream_>>>((const double*)pred, truth, n, lo, hi, sums);
  else
    k_rating_metrics<float><<<(unsigned)grid, 256, 0, (cudaStream_t)stream_>>>((const float*)pred, truth, n, lo, hi, sums);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}
