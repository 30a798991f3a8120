// SVD rating model: reference src/models/basic/models/svd.py.
//
//   :66-72  __predict      rating_ = reduce_sum(matmul(user_embed, kernel) * items_embed, 1)
//   :52-64  loss           l2_loss(predict - rating) + reg * (l2_loss(U_u) + l2_loss(V_i))      (no L2 on the kernel)
//   :74-80  __optimize__   Adagrad on user_embed, kernel and item_embed
// The kernel matrix makes every minibatch touch a dense d x d parameter, so the step is done in the gradient-only
// form of the replicated mode: cf_svd_grads red.adds every pair's row gradients into dense tables gU / gV and the
// block-local sum of e * U_u (x) V_i into gK; cf_apply_dense then applies the three tables (rows with an all-zero
// gradient are skipped = TF's sparse apply on the gathered rows; K's gradient is dense like in TF).
// One block per pair (grid-stride), K and the block's partial gK live in shared memory.
#include "common.cuh"

namespace {

struct SvdDev {
  const float *U, *V, *K;
  long long n_users, n_items;
  int d, ld, ldk, lds;      // lds = odd row stride of the shared-memory copies of K / gK (bank-conflict free)
  const int32_t* pairs;
  const float* ratings;
  long long B;
  float reg;
  float *gU, *gV, *gK;
  double* loss;
  int32_t* counters;
};

constexpr int SVD_THREADS = 128;

__device__ __forceinline__ float block_sum(float x, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  __syncthreads();                       // red may still be read from the previous call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < SVD_THREADS / 32; ++w) s += red[w];
  return s;
}

__global__ void __launch_bounds__(SVD_THREADS) k_svd_grads(const __grid_constant__ SvdDev P) {
  extern __shared__ float sm[];
  const int d = P.d, lds = P.lds;
  float* Ks = sm;                     // [d][lds]
  float* Gs = Ks + d * lds;           // [d][lds] this block's partial gradient of K
  float* su = Gs + d * lds;           // [d]
  float* sv = su + d;                 // [d]
  float* st = sv + d;                 // [d]  t = K v
  float* ss = st + d;                 // [d]  s = u K
  float* red = ss + d;                // [4]
  const int tid = threadIdx.x;
  for (int k = tid; k < d * d; k += SVD_THREADS) {
    const int a = k / d, c = k - a * d;
    Ks[a * lds + c] = P.K[a * P.ldk + c];
    Gs[a * lds + c] = 0.f;
  }
  double loss_acc = 0.0;
  __syncthreads();
  for (long long b = blockIdx.x; b < P.B; b += gridDim.x) {
    const long long u = __ldg(P.pairs + 2 * b), i = __ldg(P.pairs + 2 * b + 1);
    if (u < 0 || u >= P.n_users || i < 0 || i >= P.n_items) {      // block-uniform
      if (tid == 0) atomicOr(P.counters + 1, CF_FLAG_INDEX_RANGE);
      continue;
    }
    __syncthreads();                   // the previous pair's su / sv / st / ss are no longer read
    for (int a = tid; a < d; a += SVD_THREADS) {
      su[a] = P.U[u * P.ld + a];
      sv[a] = P.V[i * P.ld + a];
    }
    __syncthreads();
    float part = 0.f, sq = 0.f;
    for (int a = tid; a < d; a += SVD_THREADS) {
      float t = 0.f, s = 0.f;
      for (int c = 0; c < d; ++c) {
        t = fmaf(Ks[a * lds + c], sv[c], t);     // (K v)_a
        s = fmaf(su[c], Ks[c * lds + a], s);     // (u K)_a
      }
      st[a] = t;
      ss[a] = s;
      part = fmaf(su[a], t, part);
      sq += su[a] * su[a] + sv[a] * sv[a];
    }
    const float pred = block_sum(part, red);
    const float e = pred - __ldg(P.ratings + b);
    if (P.loss) {
      const float regsq = block_sum(sq, red);
      if (tid == 0) loss_acc += (double)(0.5f * e * e + 0.5f * P.reg * regsq);
    }
    for (int a = tid; a < d; a += SVD_THREADS) {
      atomicAdd(P.gU + u * P.ld + a, fmaf(e, st[a], P.reg * su[a]));
      atomicAdd(P.gV + i * P.ld + a, fmaf(e, ss[a], P.reg * sv[a]));
    }
    for (int k = tid; k < d * d; k += SVD_THREADS) {   // every entry of Gs is owned by one thread: no atomics
      const int a = k / d, c = k - a * d;
      Gs[a * lds + c] = fmaf(e * su[a], sv[c], Gs[a * lds + c]);
    }
  }
  __syncthreads();
  for (int k = tid; k < d * d; k += SVD_THREADS) {
    const int a = k / d, c = k - a * d;
    const float g = Gs[a * lds + c];
    if (g != 0.f) atomicAdd(P.gK + a * P.ldk + c, g);
  }
  if (P.loss && tid == 0 && loss_acc != 0.0) atomicAdd(P.loss, loss_acc);
}

// predictions of explicit (user, item) rows: sum_a sum_c U_u[a] K[a][c] V_i[c], fp64 accumulation
__global__ void __launch_bounds__(SVD_THREADS) k_svd_predict(const __grid_constant__ SvdDev P, long long n, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int d = P.d, lds = P.lds;
  float* Ks = sm;
  for (int k = threadIdx.x; k < d * d; k += SVD_THREADS) {
    const int a = k / d, c = k - a * d;
    Ks[a * lds + c] = P.K[a * P.ldk + c];
  }
  __syncthreads();
  for (long long k = (long long)blockIdx.x * SVD_THREADS + threadIdx.x; k < n; k += (long long)gridDim.x * SVD_THREADS) {
    const long long u = __ldg(P.pairs + 2 * k), i = __ldg(P.pairs + 2 * k + 1);
    if (u < 0 || u >= P.n_users || i < 0 || i >= P.n_items) {
      atomicOr(P.counters + 1, CF_FLAG_INDEX_RANGE);
      out[k] = nanf("");
      continue;
    }
    const float* up = P.U + u * P.ld;
    const float* vp = P.V + i * P.ld;
    double acc = 0.0;
    for (int c = 0; c < d; ++c) {
      double s = 0.0;                               // (u K)_c
      for (int a = 0; a < d; ++a) s = fma((double)__ldg(up + a), (double)Ks[a * lds + c], s);
      acc = fma(s, (double)__ldg(vp + c), acc);
    }
    out[k] = (float)acc;
  }
}

int fill(SvdDev& P, const cf_svd_args* a, const char* who) {
  CF_CHECK_ARG(a != nullptr, "%s: args is NULL", who);
  CF_CHECK_ARG(a->U && a->V && a->K && a->pairs && a->counters, "%s: NULL pointer", who);
  CF_CHECK_ARG(a->d > 0 && a->d <= 128 && a->ld >= a->d && a->ldk >= a->d, "%s: n_factors up to 128 are supported (d=%d)", who, a->d);
  CF_CHECK_ARG(a->n_users > 0 && a->n_items > 0 && a->B >= 0, "%s: bad sizes", who);
  P.U = a->U; P.V = a->V; P.K = a->K; P.n_users = a->n_users; P.n_items = a->n_items;
  P.d = a->d; P.ld = a->ld; P.ldk = a->ldk; P.lds = a->d | 1;
  P.pairs = a->pairs; P.ratings = a->ratings; P.B = a->B; P.reg = a->reg;
  P.gU = a->gradU; P.gV = a->gradV; P.gK = a->gradK; P.loss = a->loss; P.counters = a->counters;
  return 0;
}

}  // namespace

extern "C" int cf_svd_grads(const cf_svd_args* a, void* stream_) {
  SvdDev P;
  if (int rc = fill(P, a, "cf_svd_grads")) return rc;
  CF_CHECK_ARG(a->ratings && a->gradU && a->gradV && a->gradK, "cf_svd_grads: ratings and the gradient tables are required");
  if (a->B == 0) return 0;
  const size_t smem = ((size_t)2 * P.d * P.lds + 4 * P.d + 8) * sizeof(float);
  CF_CUDA_OK(cudaFuncSetAttribute(k_svd_grads, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = a->B;
  const long long cap = (long long)cf_num_sms();
  if (grid > cap) grid = cap;
  k_svd_grads<<<(unsigned)grid, SVD_THREADS, smem, (cudaStream_t)stream_>>>(P);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cf_svd_predict_pairs(const cf_svd_args* a, float* out, void* stream_) {
  SvdDev P;
  if (int rc = fill(P, a, "cf_svd_predict_pairs")) return rc;
  CF_CHECK_ARG(out != nullptr || a->B == 0, "cf_svd_predict_pairs: out is NULL");
  if (a->B == 0) return 0;
  const size_t smem = ((size_t)P.d * P.lds) * sizeof(float);
  CF_CUDA_OK(cudaFuncSetAttribute(k_svd_predict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = (a->B + SVD_THREADS - 1) / SVD_THREADS;
  const long long cap = (long long)cf_num_sms() * 2;
  if (grid > cap) grid = cap;
  k_svd_predict<<<(unsigned)grid, SVD_THREADS, smem, (cudaStream_t)stream_>>>(P, a->B, out);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}
