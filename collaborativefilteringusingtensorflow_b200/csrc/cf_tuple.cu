// The remaining pairwise family: PRIGP 5-tuples and CPLR coefficient-weighted 4-tuples.
//
// Replaces `sess.run(train_op)` of the reference's
//   models/pl/models/prigp.py:92-137   L = sum -log s(x_ui - x_uj) + alpha sum -log s(x_ut - x_uk)
//                                          + reg (l2(U_u) + l2(V[i,j,t,k]) + l2(b[i,j,t,k])),  x_um = <U_u, V_m> + b_m;
//                                      AdagradOptimizer on user_embed and item_embed ONLY (var_list, :134) -- the bias is not trained
//   models/pl/models/cplr_u.py:99-144  c_ij = coef_ui + 1, c_tj = coef_ut + 1, c_it = c_ij / c_tj;
//                                      L = alpha sum -log s(c_it (x_ui - x_ut)) + beta sum -log s(c_tj (x_ut - x_uj))
//                                          + gamma sum -log s(c_ij (x_ui - x_uj)) + reg (l2(U_u) + l2(V[i,t,j]) + l2(b[i,t,j]));
//                                      Adagrad on user_embed, item_embed and item_bias (:141)
// and their samplers
//   samplers/sampler_prigp.py:22-52     (u, i) from the epoch's shuffled positives, j a uniform non-positive; t a uniform item of
//                                      the user's coefficient row, k a uniform item outside it; with probability Phi(nnz_coef / n_items)
//                                      (the reference compares a standard NORMAL draw, :44) k is re-drawn inside the row with another
//                                      coefficient than t's and the pair is ordered by coefficient
//   samplers/sampler_uitj_ranking.py:22-38   u uniform over the users that have positives, collaborative items and room for a
//                                      negative; i a uniform positive, t a uniform collaborative item (coefficient row minus
//                                      positives), j a uniform item in neither; coefs = (coef[u, i], coef[u, t])
// Both models share the form  x_um = <U_u, V_m> + b_m  over up to four item slots m with a per-slot coefficient g_m = dL/dx_um:
//   dU_u = sum_m g_m V_m + reg U_u,   dV_m = g_m U_u + reg V_m,   db_m = g_m + reg b_m.
// cf_tuple_grads is the gradient-only step (like cf_svd_grads): one warp per tuple red.adds the row gradients into dense
// gradient tables, cf_apply_dense applies them (rows with an all-zero gradient are skipped = TF's sparse apply).
#include <math.h>

#include "common.cuh"

namespace {

struct TupleDev {
  const float *U, *V, *b;
  long long n_users, n_items, B;
  int d, ld, model;
  const int32_t* tuples;     // PRIGP [B, 5] (u, i, j, t, k); CPLR [B, 4] (u, i, t, j)
  const float* coefs;        // CPLR [B, 2] (coef[u, i], coef[u, t])
  float alpha, beta, gamma, reg;
  float *gU, *gV, *gb;
  double* loss;
  int32_t* counters;
};

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ float softplus_neg_t(float x) { return x > 0.f ? log1pf(__expf(-x)) : (-x + log1pf(__expf(x))); }
__device__ __forceinline__ float sigm1_t(float x) { return -1.f / (1.f + expf(x)); }   // sigmoid(x) - 1

__global__ void __launch_bounds__(256) k_tuple_grads(const __grid_constant__ TupleDev P) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * blockDim.x / 32;
  const int ns = P.model == CF_TUPLE_PRIGP ? 4 : 3;       // item slots
  const int width = ns + 1;
  double loss_acc = 0.0;
  for (long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32; t < P.B; t += nw) {
    const long long u = __ldg(P.tuples + t * width);
    long long m[4] = {0, 0, 0, 0};
    bool ok = u >= 0 && u < P.n_users;
    for (int s = 0; s < ns; ++s) {
      m[s] = __ldg(P.tuples + t * width + 1 + s);
      ok = ok && m[s] >= 0 && m[s] < P.n_items;
    }
    if (!ok) {
      if (lane == 0) atomicOr(P.counters + 1, CF_FLAG_INDEX_RANGE);
      continue;
    }
    const float* up = P.U + u * P.ld;
    float x[4] = {0.f, 0.f, 0.f, 0.f}, sq = 0.f;
    for (int a = lane; a < P.d; a += 32) {
      const float uu = up[a];
      sq = fmaf(uu, uu, sq);
      for (int s = 0; s < ns; ++s) {
        const float vv = P.V[m[s] * P.ld + a];
        x[s] = fmaf(uu, vv, x[s]);
        sq = fmaf(vv, vv, sq);
      }
    }
    sq = warp_sum(sq);
    float bm[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < ns; ++s) {
      bm[s] = __ldg(P.b + m[s]);
      x[s] = warp_sum(x[s]) + bm[s];
      sq = fmaf(bm[s], bm[s], sq);
    }
    float lv;
    if (P.model == CF_TUPLE_PRIGP) {        // slots (i, j, t, k)
      const float x1 = x[0] - x[1], x2 = x[2] - x[3];
      const float s1 = sigm1_t(x1), s2 = P.alpha * sigm1_t(x2);
      g[0] = s1; g[1] = -s1; g[2] = s2; g[3] = -s2;
      lv = softplus_neg_t(x1) + P.alpha * softplus_neg_t(x2);
    } else {                                // slots (i, t, j)
      const float cij = __ldg(P.coefs + 2 * t) + 1.f, ctj = __ldg(P.coefs + 2 * t + 1) + 1.f;
      const float cit = cij / ctj;
      const float z1 = cit * (x[0] - x[1]), z2 = ctj * (x[1] - x[2]), z3 = cij * (x[0] - x[2]);
      const float a1 = P.alpha * cit * sigm1_t(z1), a2 = P.beta * ctj * sigm1_t(z2), a3 = P.gamma * cij * sigm1_t(z3);
      g[0] = a1 + a3; g[1] = a2 - a1; g[2] = -a2 - a3;
      lv = P.alpha * softplus_neg_t(z1) + P.beta * softplus_neg_t(z2) + P.gamma * softplus_neg_t(z3);
    }
    if (P.loss && lane == 0) loss_acc += (double)(lv + 0.5f * P.reg * sq);
    for (int a = lane; a < P.d; a += 32) {
      const float uu = up[a];
      float gu = P.reg * uu;
      for (int s = 0; s < ns; ++s) {
        const float vv = P.V[m[s] * P.ld + a];
        gu = fmaf(g[s], vv, gu);
        atomicAdd(P.gV + m[s] * P.ld + a, fmaf(g[s], uu, P.reg * vv));
      }
      atomicAdd(P.gU + u * P.ld + a, gu);
    }
    if (P.gb != nullptr && lane < ns) atomicAdd(P.gb + m[lane], fmaf(P.reg, bm[lane], g[lane]));
  }
  if (P.loss) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    if (lane == 0 && loss_acc != 0.0) atomicAdd(P.loss, loss_acc);
  }
}

// ------------------------------------------------------------------------------------------------ samplers
struct TSampDev {
  const long long* p_indptr; const int32_t* p_indices; const int32_t* p_rows;      // training positives (CSR + COO rows)
  const long long* c_indptr; const int32_t* c_indices; const float* c_values;       // coefficient rows (sorted columns)
  const long long* t_indptr; const int32_t* t_indices;                               // CPLR: collaborative items (coef minus positives)
  const int32_t* eligible; long long n_eligible;                                     // CPLR: users that can be drawn
  long long n_users, n_items, nnz;
  unsigned long long seed;
  long long epoch, batch0;
  int n_batches, B, model;
  int32_t* out_tuples;
  float* out_coefs;
  int32_t* flags;
};

struct Rng {
  uint32_t c0, c1, c2, k0, k1;
  uint32_t n;
  Philox4 cur;
  int have;
  __device__ Rng(unsigned long long seed, long long pos, long long epoch, uint32_t stream)
      : c0((uint32_t)pos), c1((uint32_t)(pos >> 32)), c2((uint32_t)epoch ^ (stream << 24)), k0((uint32_t)seed), k1((uint32_t)(seed >> 32)),
        n(0), have(0) {}
  __device__ uint32_t next() {
    if (have == 0) {
      cur = philox4x32_10(c0, c1, c2, n++, k0, k1);
      have = 4;
    }
    const uint32_t v = have == 4 ? cur.x : have == 3 ? cur.y : have == 2 ? cur.z : cur.w;
    --have;
    return v;
  }
  __device__ long long below(long long m) {
    const uint32_t a = next(), b = next();
    return (long long)rand_below(a, b, (uint64_t)m);
  }
  __device__ float uniform() { return ((float)(next() >> 8) + 0.5f) * (1.f / 16777216.f); }
  __device__ float normal() {      // Box-Muller
    const float u1 = uniform(), u2 = uniform();
    return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
  }
};

__device__ __forceinline__ float coef_at(const TSampDev& P, long long u, int item) {
  long long lo = P.c_indptr[u], hi = P.c_indptr[u + 1];
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    const int v = P.c_indices[mid];
    if (v == item) return P.c_values[mid];
    if (v < item) lo = mid + 1; else hi = mid;
  }
  return 0.f;
}

__global__ void __launch_bounds__(256) k_sample_tuples(const __grid_constant__ TSampDev P) {
  const long long total = (long long)P.n_batches * P.B;
  const FeistelKey key = feistel_key((uint64_t)P.nnz, P.seed, (uint64_t)P.epoch, 0x5052u);
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const long long pos = P.batch0 * P.B + q;            // position inside the epoch
    Rng rng(P.seed, pos, P.epoch, P.model == CF_TUPLE_PRIGP ? 3u : 4u);
    bool gave_up = false;
    if (P.model == CF_TUPLE_PRIGP) {
      const long long e = (long long)feistel_perm((uint64_t)pos, (uint64_t)P.nnz, key);   // the epoch's shuffle (sampler_prigp.py:24)
      const long long u = P.p_rows[e];
      const int i = P.p_indices[e];
      int j = 0, tries = 0;
      do { j = (int)rng.below(P.n_items); } while (csr_contains(P.p_indices, P.p_indptr[u], P.p_indptr[u + 1], j) && ++tries < CF_SAMPLER_MAX_TRIES);
      gave_up = gave_up || tries >= CF_SAMPLER_MAX_TRIES;
      int t = i, k = j;
      const long long clo = P.c_indptr[u], chi = P.c_indptr[u + 1];
      const long long nc = chi - clo;
      if (nc > 0) {                                       // sampler_prigp.py:37-41
        t = P.c_indices[clo + rng.below(nc)];
        tries = 0;
        do { k = (int)rng.below(P.n_items); } while (csr_contains(P.c_indices, clo, chi, k) && ++tries < CF_SAMPLER_MAX_TRIES);
        gave_up = gave_up || tries >= CF_SAMPLER_MAX_TRIES;
        // more than one distinct coefficient value in the row?  (user_coefItemset_vals, :16)
        bool varied = false;
        const float c0 = P.c_values[clo];
        for (long long z = clo + 1; z < chi && !varied; ++z) varied = P.c_values[z] != c0;
        if (varied && rng.normal() < (float)nc / (float)P.n_items) {      // :43 (a standard normal draw, as written there)
          const float ct = coef_at(P, u, t);
          float ck;
          tries = 0;
          do { k = P.c_indices[clo + rng.below(nc)]; ck = coef_at(P, u, k); } while (ck == ct && ++tries < 4 * CF_SAMPLER_MAX_TRIES);
          gave_up = gave_up || ck == ct;
          if (ct < ck) { const int s = t; t = k; k = s; }
        }
      }
      int32_t* o = P.out_tuples + q * 5;
      o[0] = (int)u; o[1] = i; o[2] = j; o[3] = t; o[4] = k;
    } else {
      const long long u = P.eligible[rng.below(P.n_eligible)];             // sampler_uitj_ranking.py:26-28, without the rejection loop
      const long long plo = P.p_indptr[u], tlo = P.t_indptr[u];
      const int i = P.p_indices[plo + rng.below(P.p_indptr[u + 1] - plo)];
      const int t = P.t_indices[tlo + rng.below(P.t_indptr[u + 1] - tlo)];
      int j = 0, tries = 0;
      do {
        j = (int)rng.below(P.n_items);
      } while ((csr_contains(P.p_indices, plo, P.p_indptr[u + 1], j) || csr_contains(P.t_indices, tlo, P.t_indptr[u + 1], j)) &&
               ++tries < CF_SAMPLER_MAX_TRIES);
      gave_up = gave_up || tries >= CF_SAMPLER_MAX_TRIES;
      int32_t* o = P.out_tuples + q * 4;
      o[0] = (int)u; o[1] = i; o[2] = t; o[3] = j;
      P.out_coefs[2 * q] = coef_at(P, u, i);
      P.out_coefs[2 * q + 1] = coef_at(P, u, t);
    }
    if (gave_up) atomicOr(P.flags, CF_FLAG_SAMPLER_GAVEUP);
  }
}

}  // namespace

extern "C" int cf_tuple_grads(const cf_tuple_args* a, void* stream_) {
  CF_CHECK_ARG(a != nullptr, "cf_tuple_grads: args is NULL");
  CF_CHECK_ARG(a->model == CF_TUPLE_PRIGP || a->model == CF_TUPLE_CPLR, "cf_tuple_grads: unknown model %d", a->model);
  CF_CHECK_ARG(a->U && a->V && a->b && a->tuples && a->gradU && a->gradV && a->counters, "cf_tuple_grads: NULL pointer");
  CF_CHECK_ARG(a->model != CF_TUPLE_CPLR || (a->coefs && a->gradb), "cf_tuple_grads: CPLR needs coefs and gradb");
  CF_CHECK_ARG(a->d > 0 && a->ld >= a->d && a->n_users > 0 && a->n_items > 0 && a->B >= 0, "cf_tuple_grads: bad sizes");
  if (a->B == 0) return 0;
  TupleDev P = {};
  P.U = a->U; P.V = a->V; P.b = a->b; P.n_users = a->n_users; P.n_items = a->n_items; P.B = a->B;
  P.d = a->d; P.ld = a->ld; P.model = a->model; P.tuples = a->tuples; P.coefs = a->coefs;
  P.alpha = a->alpha; P.beta = a->beta; P.gamma = a->gamma; P.reg = a->reg;
  P.gU = a->gradU; P.gV = a->gradV; P.gb = a->model == CF_TUPLE_CPLR ? a->gradb : nullptr;   // prigp.py:134: the bias is not trained
  P.loss = a->loss; P.counters = a->counters;
  long long grid = (a->B + 7) / 8;
  const long long cap = (long long)cf_num_sms() * 8;
  if (grid > cap) grid = cap;
  k_tuple_grads<<<(unsigned)grid, 256, 0, (cudaStream_t)stream_>>>(P);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cf_sample_tuples(const cf_tuple_sample_args* a, void* stream_) {
  CF_CHECK_ARG(a != nullptr, "cf_sample_tuples: args is NULL");
  CF_CHECK_ARG(a->model == CF_TUPLE_PRIGP || a->model == CF_TUPLE_CPLR, "cf_sample_tuples: unknown model %d", a->model);
  CF_CHECK_ARG(a->train.indptr && a->train.indices && a->coef.indptr && a->coef.indices && a->coef.values && a->out_tuples && a->flags,
               "cf_sample_tuples: NULL pointer");
  CF_CHECK_ARG(a->B > 0 && a->n_batches >= 0 && a->train.n_cols > 0, "cf_sample_tuples: bad sizes");
  if (a->model == CF_TUPLE_PRIGP) {
    CF_CHECK_ARG(a->train.rows != nullptr && a->train.nnz > 0, "cf_sample_tuples: PRIGP needs the COO rows of the training CSR");
    CF_CHECK_ARG((a->batch0 + a->n_batches) * (int64_t)a->B <= a->train.nnz, "cf_sample_tuples: batches beyond the epoch (int(nnz / B) per epoch)");
  } else {
    CF_CHECK_ARG(a->collab.indptr && a->collab.indices && a->eligible && a->n_eligible > 0 && a->out_coefs,
                 "cf_sample_tuples: CPLR needs the collaborative rows, the eligible users and out_coefs");
  }
  if (a->n_batches == 0) return 0;
  TSampDev P = {};
  P.p_indptr = (const long long*)a->train.indptr; P.p_indices = a->train.indices; P.p_rows = a->train.rows;
  P.c_indptr = (const long long*)a->coef.indptr; P.c_indices = a->coef.indices; P.c_values = a->coef.values;
  P.t_indptr = (const long long*)a->collab.indptr; P.t_indices = a->collab.indices;
  P.eligible = a->eligible; P.n_eligible = a->n_eligible;
  P.n_users = a->train.n_rows; P.n_items = a->train.n_cols; P.nnz = a->train.nnz;
  P.seed = a->seed; P.epoch = a->epoch; P.batch0 = a->batch0; P.n_batches = a->n_batches; P.B = a->B; P.model = a->model;
  P.out_tuples = a->out_tuples; P.out_coefs = a->out_coefs; P.flags = a->flags;
  const long long total = (long long)a->n_batches * a->B;
  long long grid = (total + 255) / 256;
  const long long cap = (long long)cf_num_sms() * 8;
  if (grid > cap) grid = cap;
  k_sample_tuples<<<(unsigned)grid, 256, 0, (cudaStream_t)stream_>>>(P);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}
