// Fused minibatch update step for BPRMF / CML / GBPRMF / WRMF on sm_100a.
//
// Replaces, per minibatch, the whole TF1 train_op of the reference:
//   gathers + loss + autodiff + Adagrad sparse apply (+ CML clip)
//   bprmf.py:52-88, cml.py:55-129, gbprmf.py:58-106, basic/models/wrmf.py:52-88   (reference src/models/...)
// Math: SURVEY.md Appendix A.  One group of LPG lanes (8/16/32, by row width) owns one (user, item) pair:
// it gathers the u / i / j_1..W (/ g_1..G) rows with 128-bit L2-coherent loads, forms the gradients in
// registers and commits every row exactly once per occurrence.
//
// Minibatch-synchronous semantics (CF_UPDATE_SYNC, the reference's), three launches per minibatch:
//   k_count         per table row, how many times it occurs in the minibatch; rows occurring more than once get a slot of an
//                   L2-resident gradient staging buffer (slot id = occurrence index of the second occurrence: no counter)
//   k_step          a row that occurs once is updated straight from registers (read param + acc, write param + acc: the
//                   algorithmic minimum); a row that occurs T > 1 times gets its T gradients red.add-ed into its slot
//   k_apply_staged  applies every staged (summed) gradient once and returns slot / occurrence word to zero
// A row is only ever written after every pair that reads it has read it, so all gradients are evaluated at pre-update
// parameters, like TF.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

#include "cf_step_impl.cuh"

using namespace cfstep;

// shapes: 0: LPG 8 (ld <= 32)  1: LPG 16 (ld <= 64)  2: LPG 32 (ld <= 128)  3: LPG 32 x2 (ld <= 256)  4: LPG 32 x4 (ld <= 512)
#define CF_DECL_MODEL(M) CF_STEP_PICK_DECL(M, 0); CF_STEP_PICK_DECL(M, 1); CF_STEP_PICK_DECL(M, 2); CF_STEP_PICK_DECL(M, 3); CF_STEP_PICK_DECL(M, 4);
CF_DECL_MODEL(0) CF_DECL_MODEL(1) CF_DECL_MODEL(2) CF_DECL_MODEL(3)
CF_APPLY_PICK_DECL(0); CF_APPLY_PICK_DECL(1); CF_APPLY_PICK_DECL(2); CF_APPLY_PICK_DECL(3); CF_APPLY_PICK_DECL(4);
CF_SCATTER_PICK_DECL(0); CF_SCATTER_PICK_DECL(1); CF_SCATTER_PICK_DECL(2); CF_SCATTER_PICK_DECL(3); CF_SCATTER_PICK_DECL(4);

namespace {

// ---- occurrence counting + staging-slot assignment (SYNC mode), one thread per (pair, role)
__global__ void __launch_bounds__(256) k_count(const __grid_constant__ StepDev P) {
  const int R = (P.model == CF_MODEL_WRMF) ? 2 : 2 + P.W + P.G;
  const long long total = (long long)P.B * R;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t0 = (long long)blockIdx.x * blockDim.x; t0 < total; t0 += stride) {
    const long long t = t0 + threadIdx.x;
    bool need = false;
    int tab = 0;
    long long r = 0;
    if (t < total) {
      const long long b = t / R;
      const int k = (int)(t - b * R);
      if (k == 0) { tab = 0; r = __ldg(P.pairs + 2 * b); }
      else if (k == 1) { tab = 1; r = __ldg(P.pairs + 2 * b + 1); }
      else if (k < 2 + P.W) { tab = 1; r = __ldg(P.negs + b * P.W + (k - 2)); }
      else { tab = 0; r = __ldg(P.group + b * P.G + (k - 2 - P.W)); }
      if (tab ? P.gradV != nullptr : P.gradU != nullptr) { /* fetched / replicated rows are not applied here */ }
      else if (!in_range(r, tab ? P.n_items : P.n_users)) atomicOr(P.counters + 1, CF_FLAG_INDEX_RANGE);
      else need = atomicAdd((tab ? P.metaV : P.metaU) + r, 1u) == 1u;   // second occurrence: the row needs a staging slot
    }
    if (need) {
      // The staging slot of a duplicated row is the index t of its SECOND occurrence: unique per row, needs no shared
      // allocation counter (a same-address atomic per duplicated row serialised this kernel), at the price of a
      // staging buffer with one (mostly untouched) slot per occurrence.
      (tab ? P.slotV : P.slotU)[r] = (int)t;
      P.slot_row[t] = (uint32_t)r | (tab ? 0x80000000u : 0u);
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


step_kernel_t pick_apply(int nvec) {
  const int shape = nvec <= 8 ? 0 : nvec <= 16 ? 1 : nvec <= 32 ? 2 : nvec <= 64 ? 3 : 4;
  switch (shape) {
    case 0: return cf_apply_pick_0();
    case 1: return cf_apply_pick_1();
    case 2: return cf_apply_pick_2();
    case 3: return cf_apply_pick_3();
    default: return cf_apply_pick_4();
  }
}

scatter_kernel_t pick_scatter(int nvec) {
  const int shape = nvec <= 8 ? 0 : nvec <= 16 ? 1 : nvec <= 32 ? 2 : nvec <= 64 ? 3 : 4;
  switch (shape) {
    case 0: return cf_scatter_pick_0();
    case 1: return cf_scatter_pick_1();
    case 2: return cf_scatter_pick_2();
    case 3: return cf_scatter_pick_3();
    default: return cf_scatter_pick_4();
  }
}

// owner-side counting: one thread per received gradient row
__global__ void __launch_bounds__(256) k_count_rows(const __grid_constant__ StepDev P, const int32_t* __restrict__ rows, long long n) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long r = __ldg(rows + t);
    if (!in_range(r, P.n_users)) { atomicOr(P.counters + 1, CF_FLAG_INDEX_RANGE); continue; }
    if (atomicAdd(P.metaU + r, 1u) == 1u) {
      P.slotU[r] = (int)t;
      P.slot_row[t] = (uint32_t)r;
    }
  }
}

step_kernel_t pick_kernel(int model, int nvec, int* lpg, bool ext) {
  const int shape = nvec <= 8 ? 0 : nvec <= 16 ? 1 : nvec <= 32 ? 2 : nvec <= 64 ? 3 : 4;
  *lpg = shape == 0 ? 8 : shape == 1 ? 16 : 32;
#define CF_CASE(M)                                                                   \
  case M:                                                                            \
    switch (shape) {                                                                 \
      case 0: return ext ? cf_step_pick_ext_##M##_0() : cf_step_pick_##M##_0();     \
      case 1: return ext ? cf_step_pick_ext_##M##_1() : cf_step_pick_##M##_1();     \
      case 2: return ext ? cf_step_pick_ext_##M##_2() : cf_step_pick_##M##_2();     \
      case 3: return ext ? cf_step_pick_ext_##M##_3() : cf_step_pick_##M##_3();     \
      default: return ext ? cf_step_pick_ext_##M##_4() : cf_step_pick_##M##_4();    \
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


cfstep::step_kernel_t cf_pick_apply_kernel(int nvec) { return pick_apply(nvec); }   // used by cf_exchange.cu
cfstep::step_kernel_t cf_step_pick_fast(int model, int W, int G, int lpg, int* nbuf, int* slots, int* threads);   // cf_step_fast.cu

extern "C" int64_t cf_step_staging_rows(int32_t model, int32_t B, int32_t W, int32_t G) {
  const int64_t R = (model == CF_MODEL_WRMF) ? 2 : 2 + (int64_t)W + G;
  return (int64_t)B * R;  // slot id = occurrence index of a duplicated row's second occurrence
}

extern "C" int32_t cf_step_launches_per_batch(void) { return 3; }

static int train_steps_impl(const cf_step_args* a, cudaStream_t stream, cudaEvent_t* ev) {
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
    CF_CHECK_ARG(a->metaU && a->metaV && a->slotU && a->slotV && a->slot_row && a->staging && a->counters, "cf_train_steps: SYNC mode needs the workspace");
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
  P.metaU = a->metaU; P.metaV = a->metaV; P.slot_row = a->slot_row;
  P.slotU = a->slotU; P.slotV = a->slotV; P.staging = a->staging; P.staging_rows = a->staging_rows;
  P.lds = a->ld + 4; P.counters = a->counters;
  P.gradV = a->gradV; P.rank_items = a->rank_items > 0 ? a->rank_items : a->n_items;
  P.gradU = a->gradU; P.gradb = a->gradb;
  if (a->gradV || a->n_peers) CF_CHECK_ARG(a->update == CF_UPDATE_SYNC, "cf_train_steps: exchange mode (gradV / peers) needs SYNC mode");
  if (a->gradV && a->model == CF_MODEL_GBPR)
    CF_CHECK_ARG(a->gradU && a->gradb && a->n_peers == 0, "cf_train_steps: GBPR in exchange mode needs gradU, gradV and gradb (replicated data-parallel mode)");
  if (a->gradU) CF_CHECK_ARG(a->gradV != nullptr && a->n_peers == 0, "cf_train_steps: gradU needs gradV (dense gradient tables of the replicated mode)");
  P.n_peers = a->n_peers; P.gslot_pos = a->gslot_pos; P.gslot_neg = a->gslot_neg;
  for (int k = 0; k < CF_MAX_PEERS; ++k) P.peerV[k] = k < a->n_peers ? a->peerV[k] : nullptr;
  for (int k = 0; k < CF_MAX_PEERS; ++k) P.peerG[k] = k < a->n_peers ? a->peerG[k] : nullptr;
  if (a->n_peers != 0) {
    CF_CHECK_ARG(a->n_peers > 0 && a->n_peers <= CF_MAX_PEERS, "cf_train_steps: n_peers must be in [0, %d]", CF_MAX_PEERS);
    if (a->peerG[0] != nullptr) {   // push: gradients go to the owners' dense tables; gradV only marks the item rows as remote
      for (int k = 0; k < a->n_peers; ++k) CF_CHECK_ARG(a->peerG[k] != nullptr, "cf_train_steps: peerG[%d] is NULL", k);
      CF_CHECK_ARG(a->gslot_pos == nullptr && a->gslot_neg == nullptr, "cf_train_steps: peerG (push) and gslot_* (compact buffer) exclude each other");
      P.gradV = a->peerG[0];
      P.gslot_pos = nullptr; P.gslot_neg = nullptr;
    } else {
      CF_CHECK_ARG(a->gradV != nullptr && a->gslot_pos != nullptr && (W == 0 || a->gslot_neg != nullptr), "cf_train_steps: peer pull needs gradV and the gradient slots (or peerG)");
    }
    CF_CHECK_ARG(a->n_batches == 1, "cf_train_steps: peer pull takes one minibatch per call");
    for (int k = 0; k < a->n_peers; ++k) CF_CHECK_ARG(a->peerV[k] != nullptr, "cf_train_steps: peerV[%d] is NULL", k);
  }

  int lpg = 32;
  // the fetched-rows exchange mode (gradV only) is served by the single-GPU kernel (it always carried that branch, in
  // 64 registers); only peer pull and the dense-gradient mode need the larger variant
  step_kernel_t kern = pick_kernel(a->model, P.nvec, &lpg, a->n_peers > 0 || a->gradU != nullptr);
  static int sms = 0;
  if (!sms) sms = cf_num_sms();
  // entries (negatives + group users) staged per tile: all of them if the lanes (one per slot) and the shared
  // memory (two rows per slot per group) allow, else the largest tile that still fits two blocks per SM
  const int E = W + G;
  const long long groups_per_block = 256 / lpg;
  const long long slot_bytes = groups_per_block * 2ll * a->ld * 4ll;   // one more slot costs this much smem per block
  int T = E < lpg - 2 ? E : lpg - 2;
  const long long budget2 = 112 * 1024, budget1 = 224 * 1024;
  if ((2 + T) * slot_bytes > budget2) {
    int t2 = (int)(budget2 / slot_bytes) - 2;
    if (t2 >= 4 || t2 >= E) T = t2 < T ? t2 : T;
    else { int t1 = (int)(budget1 / slot_bytes) - 2; T = t1 < T ? t1 : T; }
  }
  if (E > 0 && T < 1) T = 1;
  CF_CHECK_ARG((2 + T) * slot_bytes <= budget1, "cf_train_steps: rows too wide for the shared-memory staging (ld=%d)", a->ld);
  P.T = T;
  size_t smem = (size_t)((2 + T) * slot_bytes);
  int threads = 256;
  // one negative per pair (BPRMF's reference setting) / GBPR with 5 negatives and a group of 3 or 1, rows of 36..128 floats,
  // single GPU, SYNC: the unrolled, software-pipelined form of the same arithmetic (cf_step_fast.cu).  CF_STEP_GENERIC=1
  // keeps the generic kernel (tests compare the two).
  const char* genv = getenv("CF_STEP_GENERIC");
  const bool generic_only = genv && atoi(genv) > 0;
  if (!generic_only && a->update == CF_UPDATE_SYNC && !a->gradV && !a->gradU && a->n_peers == 0 && P.nvec > 8 && P.nvec <= 32) {
    int nbuf = 0, slots = 0, thr = 256;
    if (step_kernel_t fast = cf_step_pick_fast(a->model, W, G, lpg, &nbuf, &slots, &thr)) {
      kern = fast;
      threads = thr;
      smem = (size_t)(threads / lpg) * nbuf * 2 * slots * a->ld * 4;
    }
  }
  CF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
  if (occ < 1) occ = 1;
  long long grid = (a->B + (threads / lpg) - 1) / (threads / lpg);
  const long long cap = (long long)sms * occ;
  if (grid > cap) grid = cap;
  step_kernel_t kapply = pick_apply(P.nvec);
  const long long R = (a->model == CF_MODEL_WRMF) ? 2 : 2 + W + G;
  P.n_occ = (long long)a->B * R;
  long long agrid = ((long long)a->B * R + 255) / 256;   // every thread scans one slot code per iteration
  // CF_APPLY_BLOCKS_PER_SM (default 16 = two waves of the 8 resident blocks): fewer leaves block slots to a sampler launch
  // that runs beside the apply on another stream (models/_base.py::_epoch)
  static int apply_bps = 0;
  if (!apply_bps) {
    const char* e = getenv("CF_APPLY_BLOCKS_PER_SM");
    apply_bps = (e && atoi(e) > 0) ? atoi(e) : 16;
  }
  if (agrid > (long long)sms * apply_bps) agrid = (long long)sms * apply_bps;
  long long cgrid = ((long long)a->B * R + 255) / 256;
  if (cgrid > (long long)sms * 8) cgrid = (long long)sms * 8;

  for (int nb = 0; nb < a->n_batches; ++nb) {
    const long long off = (long long)nb * a->B;
    P.pairs = a->pairs + 2 * off;
    P.negs = a->negs ? a->negs + off * W : nullptr;
    P.group = (G && a->group) ? a->group + off * G : nullptr;
    P.ratings = a->ratings ? a->ratings + off : nullptr;
    P.loss = a->loss ? a->loss + nb : nullptr;
    if (ev) CF_CUDA_OK(cudaEventRecord(ev[4 * nb + 0], stream));
    const bool all_ext = a->gradU && a->gradV;   // nothing is applied locally: no occurrence counts, no staged rows
    if (a->update == CF_UPDATE_SYNC && !all_ext) k_count<<<(unsigned)cgrid, 256, 0, stream>>>(P);
    if (ev) CF_CUDA_OK(cudaEventRecord(ev[4 * nb + 1], stream));
    kern<<<(unsigned)grid, threads, smem, stream>>>(P);
    if (a->event_after_step && nb == a->n_batches - 1) CF_CUDA_OK(cudaEventRecord((cudaEvent_t)a->event_after_step, stream));
    if (ev) CF_CUDA_OK(cudaEventRecord(ev[4 * nb + 2], stream));
    if (a->update == CF_UPDATE_SYNC && !all_ext) kapply<<<(unsigned)agrid, 256, 0, stream>>>(P);
    if (ev) CF_CUDA_OK(cudaEventRecord(ev[4 * nb + 3], stream));
  }
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cf_train_steps(const cf_step_args* a, void* stream_) {
  return train_steps_impl(a, (cudaStream_t)stream_, nullptr);
}

extern "C" int cf_train_steps_profiled(const cf_step_args* a, void* stream_, float* ms_count_host, float* ms_step_host,
                                       float* ms_apply_host) {
  CF_CHECK_ARG(a != nullptr && ms_count_host && ms_step_host && ms_apply_host, "cf_train_steps_profiled: NULL argument");
  CF_CHECK_ARG(a->n_batches > 0 && a->n_batches <= 4096, "cf_train_steps_profiled: n_batches must be in [1, 4096]");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int n = 4 * a->n_batches;
  cudaEvent_t* ev = new cudaEvent_t[n];
  for (int k = 0; k < n; ++k) CF_CUDA_OK(cudaEventCreate(&ev[k]));
  int rc = train_steps_impl(a, stream, ev);
  if (rc == 0) {
    CF_CUDA_OK(cudaStreamSynchronize(stream));
    double tc = 0.0, ts = 0.0, ta = 0.0;
    for (int nb = 0; nb < a->n_batches; ++nb) {
      float x = 0.f, y = 0.f, z = 0.f;
      CF_CUDA_OK(cudaEventElapsedTime(&x, ev[4 * nb], ev[4 * nb + 1]));
      CF_CUDA_OK(cudaEventElapsedTime(&y, ev[4 * nb + 1], ev[4 * nb + 2]));
      CF_CUDA_OK(cudaEventElapsedTime(&z, ev[4 * nb + 2], ev[4 * nb + 3]));
      tc += x;
      ts += y;
      ta += z;
    }
    *ms_count_host = (float)tc;
    *ms_step_host = (float)ts;
    *ms_apply_host = (float)ta;
  }
  for (int k = 0; k < n; ++k) cudaEventDestroy(ev[k]);
  delete[] ev;
  return rc;
}

extern "C" int cf_apply_rows(const cf_apply_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CF_CHECK_ARG(a != nullptr, "cf_apply_rows: args is NULL");
  CF_CHECK_ARG(a->table && a->rows && (a->grads || a->n_segs > 0) && a->meta && a->slot && a->slot_row && a->staging && a->counters, "cf_apply_rows: NULL pointer");
  CF_CHECK_ARG(a->n_segs >= 0 && a->n_segs <= CF_MAX_PEERS, "cf_apply_rows: n_segs must be in [0, %d]", CF_MAX_PEERS);
  GradSegs S = {};
  S.n = a->n_segs;
  for (int p = 0; p < a->n_segs; ++p) {
    S.base[p] = a->seg_grads[p];
    S.start[p] = a->seg_start[p];
    CF_CHECK_ARG(a->seg_start[p] <= a->seg_start[p + 1] && (a->seg_start[p] == a->seg_start[p + 1] || a->seg_grads[p] != nullptr), "cf_apply_rows: bad segment %d", p);
  }
  if (a->n_segs > 0) {
    CF_CHECK_ARG(a->first_seg >= 0 && a->first_seg < a->n_segs, "cf_apply_rows: first_seg out of range");
    S.rot = a->seg_start[a->first_seg];
    S.start[a->n_segs] = a->seg_start[a->n_segs];
    CF_CHECK_ARG(a->seg_start[0] == 0 && a->seg_start[a->n_segs] == a->n, "cf_apply_rows: the segments must cover the n rows");
  }
  CF_CHECK_ARG(a->d > 0 && a->ld >= a->d && a->ld % 4 == 0 && a->ld <= 512 && a->ldg >= a->ld && a->ldg % 4 == 0, "cf_apply_rows: bad d/ld/ldg");
  CF_CHECK_ARG(a->optimizer == CF_OPT_SGD || a->acc, "cf_apply_rows: Adagrad needs the accumulator table");
  CF_CHECK_ARG(a->n >= 0 && a->staging_rows >= a->n, "cf_apply_rows: staging_rows %lld < n %lld", (long long)a->staging_rows, (long long)a->n);
  if (a->n == 0) return 0;
  StepDev P = {};
  P.U = a->table; P.accU = a->acc; P.V = nullptr; P.accV = nullptr; P.b = nullptr; P.accb = nullptr;
  P.n_users = a->n_rows; P.n_items = 0; P.d = a->d; P.ld = a->ld; P.nvec = a->ld / 4;
  P.model = a->model; P.optimizer = a->optimizer; P.update = CF_UPDATE_SYNC;
  P.lr = a->lr; P.clip = a->clip_norm;
  P.metaU = a->meta; P.slotU = a->slot; P.slot_row = a->slot_row; P.staging = a->staging; P.staging_rows = a->staging_rows;
  P.lds = a->ld + 4; P.counters = a->counters; P.n_occ = a->n;
  static int sms = 0;
  if (!sms) sms = cf_num_sms();
  const int lpg = P.nvec <= 8 ? 8 : (P.nvec <= 16 ? 16 : 32);
  long long cgrid = (a->n + 255) / 256, sgrid = (a->n + (256 / lpg) - 1) / (256 / lpg), agrid = (a->n + 255) / 256;
  const long long cap = (long long)sms * 16;
  if (cgrid > cap) cgrid = cap;
  if (sgrid > cap) sgrid = cap;
  if (agrid > cap) agrid = cap;
  k_count_rows<<<(unsigned)cgrid, 256, 0, stream>>>(P, a->rows, a->n);
  pick_scatter(P.nvec)<<<(unsigned)sgrid, 256, 0, stream>>>(P, a->rows, a->grads, a->n, a->ldg, S);
  pick_apply(P.nvec)<<<(unsigned)agrid, 256, 0, stream>>>(P);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- replicated data-parallel mode: apply a DENSE (all-reduced) gradient table; rows with an all-zero gradient are
// skipped (for Adagrad / SGD a zero gradient is a no-op, so this equals the sparse apply), used rows are zeroed again
namespace {
template <int LPG, int NV>
__global__ void __launch_bounds__(256) k_apply_dense(const __grid_constant__ StepDev P, float* grad, int ldg) {
  const int lane = threadIdx.x & 31, gl = lane & (LPG - 1), leader = lane & ~(LPG - 1);
  const unsigned gmask = LPG == 32 ? 0xffffffffu : (((1u << LPG) - 1u) << leader);
  const long long ngroups = (long long)gridDim.x * blockDim.x / LPG;
  const bool adagrad = P.optimizer == CF_OPT_ADAGRAD;
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPG; r < P.n_users; r += ngroups) {
    const Row<NV> g = load_row<LPG, NV>(grad, r, ldg, P.nvec, gl);
    bool nz = false;
#pragma unroll
    for (int k = 0; k < NV; ++k) nz = nz || g.v[k].x != 0.f || g.v[k].y != 0.f || g.v[k].z != 0.f || g.v[k].w != 0.f;
    if (!__any_sync(gmask, nz)) continue;
    const Row<NV> cur = load_row<LPG, NV>(P.U, r, P.ld, P.nvec, gl);
    Row<NV> acc = adagrad ? load_row<LPG, NV>(P.accU, r, P.ld, P.nvec, gl, 1.f) : zero_row<NV>(), p;
    apply_math<LPG, NV>(P, cur, acc, g, p, gmask);
    if (adagrad) store_row<LPG, NV>(P.accU, r, P.ld, P.nvec, gl, acc);
    store_row<LPG, NV>(P.U, r, P.ld, P.nvec, gl, p);
    store_row<LPG, NV>(grad, r, ldg, P.nvec, gl, zero_row<NV>());
  }
}

__global__ void __launch_bounds__(256) k_apply_dense_scalar(float* b, float* accb, float* grad, long long n, int adagrad, float lr) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const float g = __ldcg(grad + r);
    if (g == 0.f) continue;
    if (adagrad) {
      const float a = fmaf(g, g, __ldcg(accb + r));
      __stcg(accb + r, a);
      __stcg(b + r, fmaf(-lr * g, rsqrtf(a), __ldcg(b + r)));
    } else {
      __stcg(b + r, fmaf(-lr, g, __ldcg(b + r)));
    }
    __stcg(grad + r, 0.f);
  }
}
}  // namespace

extern "C" int cf_apply_dense(const cf_apply_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  CF_CHECK_ARG(a != nullptr && a->table && a->grads, "cf_apply_dense: NULL pointer");
  CF_CHECK_ARG(a->n_rows > 0, "cf_apply_dense: empty table");
  CF_CHECK_ARG(a->optimizer == CF_OPT_SGD || a->acc, "cf_apply_dense: Adagrad needs the accumulator table");
  static int sms = 0;
  if (!sms) sms = cf_num_sms();
  if (a->ld == 1) {   // a bias vector (GBPR's b)
    long long grid = (a->n_rows + 255) / 256;
    if (grid > (long long)sms * 8) grid = (long long)sms * 8;
    k_apply_dense_scalar<<<(unsigned)grid, 256, 0, stream>>>(a->table, a->acc, const_cast<float*>(a->grads), a->n_rows, a->optimizer == CF_OPT_ADAGRAD, a->lr);
    CF_CUDA_OK(cudaGetLastError());
    return 0;
  }
  CF_CHECK_ARG(a->d > 0 && a->ld >= a->d && a->ld % 4 == 0 && a->ld <= 512 && a->ldg >= a->ld && a->ldg % 4 == 0, "cf_apply_dense: bad d/ld/ldg");
  if (a->model == CF_MODEL_CML) CF_CHECK_ARG(a->clip_norm > 0.f, "cf_apply_dense: CML needs clip_norm > 0");
  StepDev P = {};
  P.U = a->table; P.accU = a->acc; P.n_users = a->n_rows; P.d = a->d; P.ld = a->ld; P.nvec = a->ld / 4;
  P.model = a->model; P.optimizer = a->optimizer; P.lr = a->lr; P.clip = a->clip_norm;
  const int nvec = P.nvec;
  const int lpg = nvec <= 8 ? 8 : (nvec <= 16 ? 16 : 32);
  long long grid = (a->n_rows + (256 / lpg) - 1) / (256 / lpg);
  if (grid > (long long)sms * 8) grid = (long long)sms * 8;
  float* g = const_cast<float*>(a->grads);
  if (nvec <= 8) k_apply_dense<8, 1><<<(unsigned)grid, 256, 0, stream>>>(P, g, a->ldg);
  else if (nvec <= 16) k_apply_dense<16, 1><<<(unsigned)grid, 256, 0, stream>>>(P, g, a->ldg);
  else if (nvec <= 32) k_apply_dense<32, 1><<<(unsigned)grid, 256, 0, stream>>>(P, g, a->ldg);
  else if (nvec <= 64) k_apply_dense<32, 2><<<(unsigned)grid, 256, 0, stream>>>(P, g, a->ldg);
  else k_apply_dense<32, 4><<<(unsigned)grid, 256, 0, stream>>>(P, g, a->ldg);
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
