// Fused minibatch update step for BPRMF / CML / GBPRMF / WRMF on sm_100a.
//
// Replaces, per minibatch, the whole TF1 train_op of the reference:
//   gathers + loss + autodiff + Adagrad sparse apply (+ CML clip)
//   bprmf.py:52-88, cml.py:55-129, gbprmf.py:58-106, basic/models/wrmf.py:52-88   (reference src/models/...)
// Math: SURVEY.md Appendix A.  One group of LPG lanes (8/16/32, by row width) owns one (user, item) pair:
// it gathers the u / i / j_1..W (/ g_1..G) rows with 128-bit L2-coherent loads, forms the gradients in
// registers and commits every row exactly once per occurrence.
//
// Minibatch-synchronous semantics (CF_UPDATE_SYNC, the reference's): a counting kernel first records, per
// table row, how many times it occurs in the minibatch (meta word, low 32 bits) and gives rows that occur
// more than once a slot in an L2-resident staging buffer.  In the fused kernel a row that occurs once is
// updated straight from registers (read param + acc, write param + acc: the algorithmic minimum); a row that
// occurs T > 1 times gets its T gradients red.add-ed into its staging slot and the LAST arriver (meta word,
// high 32 bits) applies the summed gradient once.  A row is only ever written after every pair that reads it
// has finished reading, so all gradients are evaluated at pre-update parameters, like TF.
#include <math.h>

#include "common.cuh"

#include "cf_step_impl.cuh"

using namespace cfstep;

// shapes: 0: LPG 8 (ld <= 32)  1: LPG 16 (ld <= 64)  2: LPG 32 (ld <= 128)  3: LPG 32 x2 (ld <= 256)  4: LPG 32 x4 (ld <= 512)
#define CF_DECL_MODEL(M) CF_STEP_PICK_DECL(M, 0); CF_STEP_PICK_DECL(M, 1); CF_STEP_PICK_DECL(M, 2); CF_STEP_PICK_DECL(M, 3); CF_STEP_PICK_DECL(M, 4);
CF_DECL_MODEL(0) CF_DECL_MODEL(1) CF_DECL_MODEL(2) CF_DECL_MODEL(3)

namespace {

// ---- occurrence counting + staging-slot assignment (SYNC mode), one thread per (pair, role)
__global__ void __launch_bounds__(256) k_count(const __grid_constant__ StepDev P) {
  const int R = (P.model == CF_MODEL_WRMF) ? 2 : 2 + P.W + P.G;
  const long long total = (long long)P.B * R;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long b = t / R;
    const int k = (int)(t - b * R);
    int tab;
    long long r;
    if (k == 0) { tab = 0; r = __ldg(P.pairs + 2 * b); }
    else if (k == 1) { tab = 1; r = __ldg(P.pairs + 2 * b + 1); }
    else if (k < 2 + P.W) { tab = 1; r = __ldg(P.negs + b * P.W + (k - 2)); }
    else { tab = 0; r = __ldg(P.group + b * P.G + (k - 2 - P.W)); }
    if (!in_range(r, tab ? P.n_items : P.n_users)) {
      atomicOr(P.counters + 1, CF_FLAG_INDEX_RANGE);
      continue;
    }
    const unsigned long long old = atomicAdd((tab ? P.metaV : P.metaU) + r, 1ull);
    if ((unsigned)old == 1u) {  // second occurrence: this row needs a staging slot
      const int s = atomicAdd(P.counters, 1);
      if (s < P.staging_rows) (tab ? P.slotV : P.slotU)[r] = s;
      else atomicOr(P.counters + 1, CF_FLAG_STAGING_FULL);
    }
  }
}

template <int LPG, int NV>
__global__ void __launch_bounds__(256) k_clip(float* tab, long long n_rows, int ld, int nvec, float clip) {
  const int lane = threadIdx.x & 31, gl = lane & (LPG - 1), leader = lane & ~(LPG - 1);
  const unsigned gmask = LPG == 32 ? 0xffffffffu : (((1u << LPG) - 1u) << leader);
  const long long ngroups = (long long)gridDim.x * blockDim.x / LPG;
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPG; r < n_rows; r += ngroups) {
    Row<NV> p = load_row<LPG, NV>(tab, r, ld, nvec, gl);
    const float nrm = sqrtf(group_sum<LPG>(dotp<NV>(p, p), gmask));
    if (nrm > clip) {
      const float den = fmaxf(nrm, clip);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        p.v[k].x = (p.v[k].x * clip) / den;
        p.v[k].y = (p.v[k].y * clip) / den;
        p.v[k].z = (p.v[k].z * clip) / den;
        p.v[k].w = (p.v[k].w * clip) / den;
      }
      store_row<LPG, NV>(tab, r, ld, nvec, gl, p);
    }
  }
}


step_kernel_t pick_kernel(int model, int nvec, int W, int* lpg) {
  const int shape = nvec <= 8 ? 0 : nvec <= 16 ? 1 : nvec <= 32 ? 2 : nvec <= 64 ? 3 : 4;
  *lpg = shape == 0 ? 8 : shape == 1 ? 16 : 32;
#define CF_CASE(M)                                   \
  case M:                                            \
    switch (shape) {                                 \
      case 0: return cf_step_pick_##M##_0(W);        \
      case 1: return cf_step_pick_##M##_1(W);        \
      case 2: return cf_step_pick_##M##_2(W);        \
      case 3: return cf_step_pick_##M##_3(W);        \
      default: return cf_step_pick_##M##_4(W);       \
    }
  switch (model) {
    CF_CASE(0)
    CF_CASE(1)
    CF_CASE(2)
    default:
    CF_CASE(3)
  }
#undef CF_CASE
}

}  // namespace


extern "C" int64_t cf_step_staging_rows(int32_t model, int32_t B, int32_t W, int32_t G) {
  const int64_t R = (model == CF_MODEL_WRMF) ? 2 : 2 + (int64_t)W + G;
  return (int64_t)B * R / 2 + 1;  // a row needs a slot only if it occurs at least twice
}

extern "C" int32_t cf_step_launches_per_batch(void) { return 2; }

extern "C" int cf_train_steps(const cf_step_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CF_CHECK_ARG(a != nullptr, "cf_train_steps: args is NULL");
  CF_CHECK_ARG(a->model >= CF_MODEL_BPR && a->model <= CF_MODEL_WRMF, "cf_train_steps: unknown model %d", a->model);
  CF_CHECK_ARG(a->optimizer == CF_OPT_ADAGRAD || a->optimizer == CF_OPT_SGD, "cf_train_steps: unknown optimizer %d", a->optimizer);
  CF_CHECK_ARG(a->update == CF_UPDATE_SYNC || a->update == CF_UPDATE_HOGWILD, "cf_train_steps: unknown update mode %d", a->update);
  CF_CHECK_ARG(a->U && a->V && a->pairs, "cf_train_steps: U, V and pairs are required");
  CF_CHECK_ARG(a->d > 0 && a->ld >= a->d && a->ld % 4 == 0 && a->ld <= 512, "cf_train_steps: need 0 < d <= ld <= 512, ld %% 4 == 0 (d=%d ld=%d)", a->d, a->ld);
  CF_CHECK_ARG(((uintptr_t)a->U % 16 == 0) && ((uintptr_t)a->V % 16 == 0), "cf_train_steps: tables must be 16-byte aligned");
  CF_CHECK_ARG(a->B > 0 && a->n_batches >= 0, "cf_train_steps: B must be positive");
  CF_CHECK_ARG(a->n_users > 0 && a->n_items > 0 && a->n_users < (1ll << 31) && a->n_items < (1ll << 31), "cf_train_steps: table sizes must fit int32 ids");
  if (a->optimizer == CF_OPT_ADAGRAD) CF_CHECK_ARG(a->accU && a->accV, "cf_train_steps: Adagrad needs accU/accV");
  int W = a->W, G = a->G;
  if (a->model == CF_MODEL_WRMF) {
    CF_CHECK_ARG(a->ratings != nullptr, "cf_train_steps: WRMF needs ratings");
    W = 0; G = 0;
  } else {
    CF_CHECK_ARG(W >= 1 && a->negs != nullptr, "cf_train_steps: ranking models need W >= 1 negatives");
  }
  if (a->model == CF_MODEL_GBPR) {
    CF_CHECK_ARG(G >= 1 && a->group != nullptr, "cf_train_steps: GBPR needs a group of G >= 1 users");
    CF_CHECK_ARG(a->b != nullptr && (a->optimizer != CF_OPT_ADAGRAD || a->accb != nullptr), "cf_train_steps: GBPR needs the item bias (and its accumulator)");
  } else {
    G = 0;
  }
  if (a->model == CF_MODEL_CML) CF_CHECK_ARG(a->clip_norm > 0.f, "cf_train_steps: CML needs clip_norm > 0");
  if (a->update == CF_UPDATE_SYNC) {
    CF_CHECK_ARG(a->metaU && a->metaV && a->slotU && a->slotV && a->staging && a->counters, "cf_train_steps: SYNC mode needs the workspace");
    CF_CHECK_ARG(a->staging_rows >= cf_step_staging_rows(a->model, a->B, W, G), "cf_train_steps: staging_rows %lld < required %lld",
                 (long long)a->staging_rows, (long long)cf_step_staging_rows(a->model, a->B, W, G));
  } else {
    CF_CHECK_ARG(a->counters != nullptr, "cf_train_steps: counters (flags) are required");
  }

  StepDev P;
  P.U = a->U; P.V = a->V; P.b = (a->model == CF_MODEL_GBPR) ? a->b : nullptr;
  P.accU = a->accU; P.accV = a->accV; P.accb = a->accb;
  P.n_users = a->n_users; P.n_items = a->n_items;
  P.d = a->d; P.ld = a->ld; P.nvec = a->ld / 4;
  P.B = a->B; P.W = W; P.G = G;
  P.model = a->model; P.optimizer = a->optimizer; P.update = a->update; P.use_rank_weight = a->use_rank_weight;
  P.lr = a->lr; P.reg = a->reg; P.margin = a->margin; P.clip = a->clip_norm; P.rho = a->rho; P.weight = a->weight;
  P.metaU = (unsigned long long*)a->metaU; P.metaV = (unsigned long long*)a->metaV;
  P.slotU = a->slotU; P.slotV = a->slotV; P.staging = a->staging; P.staging_rows = a->staging_rows;
  P.lds = a->ld + 4; P.counters = a->counters;

  int lpg = 32;
  step_kernel_t kern = pick_kernel(a->model, P.nvec, W, &lpg);
  static int sms = 0;
  if (!sms) sms = cf_num_sms();
  int occ = 0;
  CF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
  if (occ < 1) occ = 1;
  const long long groups_per_block = 256 / lpg;
  long long grid = (a->B + groups_per_block - 1) / groups_per_block;
  const long long cap = (long long)sms * occ;
  if (grid > cap) grid = cap;
  const long long R = (a->model == CF_MODEL_WRMF) ? 2 : 2 + W + G;
  long long cgrid = ((long long)a->B * R + 255) / 256;
  if (cgrid > (long long)sms * 8) cgrid = (long long)sms * 8;

  for (int nb = 0; nb < a->n_batches; ++nb) {
    const long long off = (long long)nb * a->B;
    P.pairs = a->pairs + 2 * off;
    P.negs = a->negs ? a->negs + off * W : nullptr;
    P.group = (G && a->group) ? a->group + off * G : nullptr;
    P.ratings = a->ratings ? a->ratings + off : nullptr;
    P.loss = a->loss ? a->loss + nb : nullptr;
    if (a->update == CF_UPDATE_SYNC) k_count<<<(unsigned)cgrid, 256, 0, stream>>>(P);
    kern<<<(unsigned)grid, 256, 0, stream>>>(P);
  }
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cf_clip_rows(float* table, int64_t n_rows, int32_t d, int32_t ld, float clip_norm, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CF_CHECK_ARG(table && n_rows > 0 && ld >= d && ld % 4 == 0 && ld <= 512 && clip_norm > 0.f, "cf_clip_rows: bad arguments");
  const int nvec = ld / 4;
  static int sms = 0;
  if (!sms) sms = cf_num_sms();
  const int lpg = nvec <= 8 ? 8 : (nvec <= 16 ? 16 : 32);
  long long grid = (n_rows + (256 / lpg) - 1) / (256 / lpg);
  if (grid > (long long)sms * 8) grid = (long long)sms * 8;
  if (nvec <= 8) k_clip<8, 1><<<(unsigned)grid, 256, 0, stream>>>(table, n_rows, ld, nvec, clip_norm);
  else if (nvec <= 16) k_clip<16, 1><<<(unsigned)grid, 256, 0, stream>>>(table, n_rows, ld, nvec, clip_norm);
  else if (nvec <= 32) k_clip<32, 1><<<(unsigned)grid, 256, 0, stream>>>(table, n_rows, ld, nvec, clip_norm);
  else if (nvec <= 64) k_clip<32, 2><<<(unsigned)grid, 256, 0, stream>>>(table, n_rows, ld, nvec, clip_norm);
  else k_clip<32, 4><<<(unsigned)grid, 256, 0, stream>>>(table, n_rows, ld, nvec, clip_norm);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}
