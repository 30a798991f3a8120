// Device-side item exchange of the multi-GPU training step (one process per GPU on one NVLink / NVSwitch node).
//
// The reference is single-device (bprmf.py:131 uses DEVICES[0] only): this replaces nothing there.  It replaces the first
// version of the exchange, which planned every minibatch with eager torch ops (flag scatter + cumsum + nonzero), read the
// request counts back to the host twice per minibatch and moved ids / rows / gradients with three NCCL all-to-alls.
// Here nothing returns to the host and no payload goes through NCCL: every rank publishes its requests in a "mailbox" that
// its peers map with CUDA IPC, and the kernels read counts, ids, item rows and gradient rows straight from peer memory.
//
// Per minibatch (two cross-GPU barriers, marked ||, are the caller's: 1-element NCCL all-reduces on the compute stream):
//   cf_exchange_route     k_route_assign  dedupe of the minibatch's item ids (first arriver wins an atomicCAS on a table
//                                         laid out [owner][local row]; positions per owner from block-aggregated counters)
//                                         -> counts[owner], req[owner][k] = local row at the owner      (the mailbox)
//                         k_route_fill    every occurrence -> its row ("slot") of the compact fetched / gradient buffers
//   || barrier 1: every rank's mailbox is complete; every rank finished the previous minibatch's owner-side apply
//   cf_exchange_prepare   k_owner_segs    reads the P x P counts from the peers' mailboxes: where this owner's rows sit
//                         k_exchange_prepare   requester half: gathers each requested row ONCE from its owner's shard over
//                                         NVLink (fetch mode), zeroes the gradient slot, returns the dedupe table to rest;
//                                         owner half: occurrence count of every requested row (ids read from the peers'
//                                         mailboxes) + staging slot for rows requested by several GPUs
//   (cf_train_steps in exchange mode: user rows updated in place, item-row gradients red.added into the compact buffer;
//    in pull mode the fused kernel reads the item rows from their owners itself)
//   || barrier 2: every rank's gradient buffer is complete
//   cf_exchange_apply     k_owner_scatter gradient rows read IN PLACE from the requesters' buffers (NVLink): applied
//                                         straight away (row requested by one GPU) or red.added into its staging slot
//                         k_apply_staged  the summed gradient of every shared row, applied once
// i.e. exactly the single-GPU minibatch-synchronous semantics on the global minibatch (sum duplicates, apply once).
#include <math.h>

#include "common.cuh"

#include "cf_step_impl.cuh"

using namespace cfstep;

cfstep::step_kernel_t cf_pick_apply_kernel(int nvec);   // cf_step.cu

namespace {

struct RouteDev {
  const int32_t *pairs, *negs;
  int B, W, P;
  long long n_items, L, cap;
  int32_t* slot_of;
  int32_t* counts;
  int32_t* req;
  int32_t *slot_pairs, *slot_negs, *slot_pos;
  int32_t* flags;
};

// ---- dedupe + routing: one thread per item occurrence of the minibatch
__global__ void __launch_bounds__(256) k_route_assign(const __grid_constant__ RouteDev R) {
  __shared__ int s_cnt[CF_MAX_PEERS], s_base[CF_MAX_PEERS];
  const int lane = threadIdx.x & 31;
  const int per = 1 + R.W;
  const long long total = (long long)R.B * per;
  for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += (long long)gridDim.x * blockDim.x) {
    if (threadIdx.x < CF_MAX_PEERS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const long long t = base + threadIdx.x;
    int owner = 0, row = 0, in_block = 0;
    long long key = 0;
    bool win = false;
    if (t < total) {
      const long long b = t / per;
      const int k = (int)(t - b * per);
      const int item = k == 0 ? __ldg(R.pairs + 2 * b + 1) : __ldg(R.negs + b * R.W + (k - 1));
      if (item < 0 || item >= R.n_items) {
        atomicOr(R.flags, CF_FLAG_INDEX_RANGE);
      } else {
        row = item / R.P;
        owner = item - row * R.P;
        key = (long long)owner * R.L + row;
        win = atomicCAS(R.slot_of + key, -1, -2) == -1;      // the first occurrence of the id claims it
      }
    }
    const unsigned winners = __ballot_sync(0xffffffffu, win);
    if (win) {   // one shared-memory atomic per (warp, owner)
      const unsigned same = __match_any_sync(winners, owner);
      const int leader = __ffs(same) - 1;
      int wbase = 0;
      if (lane == leader) wbase = atomicAdd(&s_cnt[owner], __popc(same));
      wbase = __shfl_sync(same, wbase, leader);
      in_block = wbase + __popc(same & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (threadIdx.x < R.P) {   // one global atomic per (block, owner)
      const int c = s_cnt[threadIdx.x];
      s_base[threadIdx.x] = c ? atomicAdd(R.counts + threadIdx.x, c) : 0;
    }
    __syncthreads();
    if (win) {
      const int pos = s_base[owner] + in_block;
      if (pos < R.cap) {
        R.req[(long long)owner * R.cap + pos] = row;
        R.slot_of[key] = pos;
      } else {   // cannot happen with cap >= min(occurrences, rows per owner); never write out of bounds
        atomicOr(R.flags, CF_FLAG_STAGING_FULL);
        R.slot_of[key] = 0;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_route_fill(const __grid_constant__ RouteDev R) {
  __shared__ int s_base[CF_MAX_PEERS + 1];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int o = 0; o < R.P; ++o) {
      s_base[o] = acc;
      acc += min(__ldcg(R.counts + o), (int)R.cap);
    }
    s_base[R.P] = acc;
  }
  __syncthreads();
  const int per = 1 + R.W;
  const long long total = (long long)R.B * per;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long b = t / per;
    const int k = (int)(t - b * per);
    const int item = k == 0 ? __ldg(R.pairs + 2 * b + 1) : __ldg(R.negs + b * R.W + (k - 1));
    int slot = 0;
    if (item >= 0 && item < R.n_items) {
      const int row = item / R.P, owner = item - row * R.P;
      slot = s_base[owner] + __ldcg(R.slot_of + (long long)owner * R.L + row);
    }
    if (k == 0) {
      R.slot_pairs[2 * b] = __ldg(R.pairs + 2 * b);
      R.slot_pairs[2 * b + 1] = slot;
      R.slot_pos[b] = slot;
    } else {
      R.slot_negs[b * R.W + (k - 1)] = slot;
    }
  }
}

// ---- the exchange as the device sees it
struct XDev {
  int P, me, ld, nvec;
  long long cap, L, n_rows;
  const int32_t* counts[CF_MAX_PEERS];
  const int32_t* req[CF_MAX_PEERS];
  const float* grads[CF_MAX_PEERS];
  const float* tables[CF_MAX_PEERS];
  float* my_grads;
  float* fetched;
  int32_t* slot_of;
  float* dense;          // push variant: this owner's dense gradient table [n_rows, ld] (the peers red.add into it)
  int32_t* touched;      //               list of the rows that were requested this minibatch; its length is segs[3P+2]
  // segs (local device memory, written by k_owner_segs):
  //   [0 .. P]          start of requester p's ids in this owner's received index space ([P] = total received)
  //   [P+1 .. 2P]       first row of this owner's segment inside requester p's compact buffers
  //   [2P+1 .. 3P+1]    start of owner o's segment inside MY compact buffers ([3P+1] = my total of requested rows)
  long long* segs;
};

__global__ void k_owner_segs(const __grid_constant__ XDev X) {
  // the P x P request counts live in the peers' mailboxes: every thread fetches ONE of them (a single thread walking all 64
  // is a chain of 64 NVLink round trips: 0.25 ms at 8 GPUs)
  __shared__ long long c[CF_MAX_PEERS][CF_MAX_PEERS];
  const int p = threadIdx.x / CF_MAX_PEERS, o = threadIdx.x % CF_MAX_PEERS;
  if (p < X.P && o < X.P) c[p][o] = min((long long)__ldcg(X.counts[p] + o), X.cap);
  __syncthreads();
  if (threadIdx.x != 0) return;
  long long start = 0;
  for (int q = 0; q < X.P; ++q) {
    long long gbase = 0;
    for (int w = 0; w < X.me; ++w) gbase += c[q][w];
    X.segs[q] = start;
    X.segs[X.P + 1 + q] = gbase;
    start += c[q][X.me];
  }
  X.segs[X.P] = start;
  long long mine = 0;
  for (int w = 0; w < X.P; ++w) {
    X.segs[2 * X.P + 1 + w] = mine;
    mine += c[X.me][w];
  }
  X.segs[3 * X.P + 1] = mine;
}

__device__ __forceinline__ int seg_of(const long long* s_start, int P, long long t) {
  int p = 0;
  while (p + 1 < P && t >= s_start[p + 1]) ++p;
  return p;
}

// requester half (blocks [0, ga)): one warp per requested row; owner half (blocks [ga, grid)): one thread per received id
__global__ void __launch_bounds__(256) k_exchange_prepare(const __grid_constant__ XDev X, const __grid_constant__ StepDev P, int ga) {
  __shared__ long long s_start[CF_MAX_PEERS + 1];
  const int lane = threadIdx.x & 31;
  if ((int)blockIdx.x < ga) {
    if (threadIdx.x <= X.P) s_start[threadIdx.x] = X.segs[2 * X.P + 1 + threadIdx.x];
    __syncthreads();
    const long long n = s_start[X.P];
    const long long nw = (long long)ga * (blockDim.x / 32);
    constexpr int UN = 4;   // rows in flight per warp: NVLink round trips are long, one 512-byte row per warp does not cover them
    for (long long s0 = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); s0 < n; s0 += UN * nw) {
      float4 head[UN];
      int rows[UN], own[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const long long s = s0 + u * nw;
        rows[u] = -1;
        own[u] = 0;
        if (s < n) {
          const int o = seg_of(s_start, X.P, s);
          own[u] = o;
          rows[u] = __ldcg(X.req[X.me] + (long long)o * X.cap + (s - s_start[o]));
          if (X.fetched != nullptr && lane < X.nvec) head[u] = ldcg4(X.tables[o] + (long long)rows[u] * X.ld + 4 * lane);
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const long long s = s0 + u * nw;
        if (rows[u] < 0) continue;
        if (X.fetched != nullptr) {   // the row, gathered once from its owner's shard (NVLink for owner != me)
          const float* src = X.tables[own[u]] + (long long)rows[u] * X.ld;
          float* dst = X.fetched + s * X.ld;
          if (lane < X.nvec) stcg4(dst + 4 * lane, head[u]);
          for (int v = lane + 32; v < X.nvec; v += 32) stcg4(dst + 4 * v, ldcg4(src + 4 * v));
        }
        if (X.dense == nullptr) {   // the compact gradient buffer collects this minibatch's gradients from zero
          float* g = X.my_grads + s * X.ld;
          for (int v = lane; v < X.nvec; v += 32) stcg4(g + 4 * v, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        if (lane == 0) X.slot_of[(long long)own[u] * X.L + rows[u]] = -1;
      }
    }
  } else {
    if (threadIdx.x <= X.P) s_start[threadIdx.x] = X.segs[threadIdx.x];
    __syncthreads();
    const long long n = s_start[X.P];
    const long long stride = (long long)(gridDim.x - ga) * blockDim.x;
    if (X.dense != nullptr) {
      // push variant: the first request of a row puts it on the list of rows to apply (positions from a block-aggregated
      // counter: one global atomic per block and round, not one per row)
      __shared__ int s_n;
      __shared__ long long s_base;
      for (long long t0 = (long long)(blockIdx.x - ga) * blockDim.x; t0 < n; t0 += stride) {
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const long long t = t0 + threadIdx.x;
        long long r = -1;
        int mine = -1;
        if (t < n) {
          const int p = seg_of(s_start, X.P, t);
          r = __ldcg(X.req[p] + (long long)X.me * X.cap + (t - s_start[p]));
          if (r < 0 || r >= X.n_rows) { atomicOr(P.counters + 1, CF_FLAG_INDEX_RANGE); r = -1; }
        }
        const bool first = r >= 0 && atomicAdd(P.metaU + r, 1u) == 0u;
        const unsigned m = __ballot_sync(0xffffffffu, first);
        int wbase = 0;
        if (m && lane == __ffs(m) - 1) wbase = atomicAdd(&s_n, __popc(m));
        wbase = __shfl_sync(0xffffffffu, wbase, m ? __ffs(m) - 1 : 0);
        if (first) mine = wbase + __popc(m & ((1u << lane) - 1u));
        __syncthreads();
        if (threadIdx.x == 0 && s_n) s_base = (long long)atomicAdd(reinterpret_cast<unsigned long long*>(X.segs + 3 * X.P + 2), (unsigned long long)s_n);
        __syncthreads();
        if (first) X.touched[s_base + mine] = (int)r;
        __syncthreads();
      }
      return;
    }
    for (long long t = (long long)(blockIdx.x - ga) * blockDim.x + threadIdx.x; t < n; t += stride) {
      const int p = seg_of(s_start, X.P, t);
      const long long r = __ldcg(X.req[p] + (long long)X.me * X.cap + (t - s_start[p]));
      if (r < 0 || r >= X.n_rows) { atomicOr(P.counters + 1, CF_FLAG_INDEX_RANGE); continue; }
      if (t >= P.staging_rows) { atomicOr(P.counters + 1, CF_FLAG_STAGING_FULL); continue; }
      if (atomicAdd(P.metaU + r, 1u) == 1u) {   // requested by a second GPU: the row's gradients are summed in a staging slot
        P.slotU[r] = (int)t;
        P.slot_row[t] = (uint32_t)r;
      }
    }
  }
}

template <int LPG, int NV>
__global__ void __launch_bounds__(256) k_owner_scatter(const __grid_constant__ XDev X, const __grid_constant__ StepDev P) {
  __shared__ long long s_start[CF_MAX_PEERS + 1], s_gbase[CF_MAX_PEERS];
  if (threadIdx.x <= X.P) s_start[threadIdx.x] = X.segs[threadIdx.x];
  if (threadIdx.x < X.P) s_gbase[threadIdx.x] = X.segs[X.P + 1 + threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, gl = lane & (LPG - 1), leader = lane & ~(LPG - 1);
  const unsigned gmask = LPG == 32 ? 0xffffffffu : (((1u << LPG) - 1u) << leader);
  const long long ngroups = (long long)gridDim.x * blockDim.x / LPG;
  const bool adagrad = P.optimizer == CF_OPT_ADAGRAD;
  const long long n = min(s_start[X.P], P.staging_rows);
  // owner r starts with requester r + 1: the P owners do not all read the same requester's buffer at the same time
  const long long rot = s_start[(X.me + 1) % X.P];
  for (long long k0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPG; k0 < n; k0 += ngroups) {
    long long k = k0 + rot;
    if (k >= n) k -= n;
    const int p = seg_of(s_start, X.P, k);
    const long long kk = k - s_start[p];
    const long long r = __ldcg(X.req[p] + (long long)X.me * X.cap + kk);
    if (r < 0 || r >= X.n_rows) continue;   // flagged by the counting half
    const unsigned occ = __ldcg(P.metaU + r);
    const Row<NV> g = load_row<LPG, NV>(X.grads[p], s_gbase[p] + kk, X.ld, P.nvec, gl);
    if (occ <= 1u) {
      const Row<NV> cur = load_row<LPG, NV>(P.U, r, P.ld, P.nvec, gl);
      Row<NV> acc, np;
      if (adagrad) acc = load_row<LPG, NV>(P.accU, r, P.ld, P.nvec, gl, 1.f);
      apply_math<LPG, NV>(P, cur, acc, g, np, gmask);
      if (adagrad) store_row<LPG, NV>(P.accU, r, P.ld, P.nvec, gl, acc);
      store_row<LPG, NV>(P.U, r, P.ld, P.nvec, gl, np);
      if (gl == 0) __stcg(P.metaU + r, 0u);
    } else {
      float* st = P.staging + (long long)__ldcg(P.slotU + r) * P.lds;
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        const int v = gl + q * LPG;
        if (v < P.nvec) atomicAdd(reinterpret_cast<float4*>(st + 4 * v), g.v[q]);
      }
    }
  }
}

// push variant: every requested row's gradients were red.added into this owner's dense table by the requesters' step
// kernels; apply each listed row once, return its gradient row and its request count to zero
template <int LPG, int NV>
__global__ void __launch_bounds__(256) k_owner_apply_dense(const __grid_constant__ XDev X, const __grid_constant__ StepDev P) {
  const int lane = threadIdx.x & 31, gl = lane & (LPG - 1), leader = lane & ~(LPG - 1);
  const unsigned gmask = LPG == 32 ? 0xffffffffu : (((1u << LPG) - 1u) << leader);
  const long long ngroups = (long long)gridDim.x * blockDim.x / LPG;
  const bool adagrad = P.optimizer == CF_OPT_ADAGRAD;
  const long long n = X.segs[3 * X.P + 2];
  for (long long k = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPG; k < n; k += ngroups) {
    const long long r = __ldg(X.touched + k);
    float* gp = X.dense + r * X.ld;
    const Row<NV> g = load_row<LPG, NV>(gp, 0, 0, P.nvec, gl);
    const Row<NV> cur = load_row<LPG, NV>(P.U, r, P.ld, P.nvec, gl);
    Row<NV> acc, np;
    if (adagrad) acc = load_row<LPG, NV>(P.accU, r, P.ld, P.nvec, gl, 1.f);
    apply_math<LPG, NV>(P, cur, acc, g, np, gmask);
    if (adagrad) store_row<LPG, NV>(P.accU, r, P.ld, P.nvec, gl, acc);
    store_row<LPG, NV>(P.U, r, P.ld, P.nvec, gl, np);
    store_row<LPG, NV>(gp, 0, 0, P.nvec, gl, zero_row<NV>());
    if (gl == 0) __stcg(P.metaU + r, 0u);
  }
}

int check_common(const cf_exchange_args* a, const char* who) {
  CF_CHECK_ARG(a != nullptr, "%s: args is NULL", who);
  CF_CHECK_ARG(a->n_ranks >= 1 && a->n_ranks <= CF_MAX_PEERS && a->rank >= 0 && a->rank < a->n_ranks, "%s: need 1 <= n_ranks <= %d and 0 <= rank < n_ranks", who, CF_MAX_PEERS);
  CF_CHECK_ARG(a->n_items_global > 0 && a->n_items_global < (1ll << 31), "%s: n_items_global must fit int32 ids", who);
  CF_CHECK_ARG(a->cap > 0, "%s: cap must be positive", who);
  for (int p = 0; p < a->n_ranks; ++p) CF_CHECK_ARG(a->counts[p] && a->req[p], "%s: mailbox of rank %d is NULL", who, p);
  return 0;
}

XDev make_xdev(const cf_exchange_args* a) {
  XDev X = {};
  X.P = a->n_ranks; X.me = a->rank; X.ld = a->ld; X.nvec = a->ld / 4;
  X.cap = a->cap; X.L = (a->n_items_global + a->n_ranks - 1) / a->n_ranks; X.n_rows = a->n_rows;
  for (int p = 0; p < a->n_ranks; ++p) {
    X.counts[p] = a->counts[p]; X.req[p] = a->req[p]; X.grads[p] = a->grads[p]; X.tables[p] = a->tables[p];
  }
  X.my_grads = a->grads[a->rank]; X.fetched = a->fetched; X.slot_of = a->slot_of; X.segs = reinterpret_cast<long long*>(a->segs);
  X.dense = a->dense_grads; X.touched = a->touched;
  return X;
}

StepDev make_owner_dev(const cf_exchange_args* a) {
  StepDev P = {};
  P.U = a->table; P.accU = a->acc; P.n_users = a->n_rows; P.d = a->d; P.ld = a->ld; P.nvec = a->ld / 4;
  P.model = a->model; P.optimizer = a->optimizer; P.update = CF_UPDATE_SYNC; P.lr = a->lr; P.clip = a->clip_norm;
  P.metaU = a->meta; P.slotU = a->slot; P.slot_row = a->slot_row; P.staging = a->staging; P.staging_rows = a->staging_rows;
  P.lds = a->ld + 4; P.counters = a->counters;
  const long long most = (long long)a->n_ranks * a->cap;
  P.n_occ = a->staging_rows < most ? a->staging_rows : most;   // k_apply_staged scans every possible slot (4 bytes each)
  return P;
}

int check_owner(const cf_exchange_args* a, const char* who) {
  if (a->dense_grads != nullptr) {   // push variant
    CF_CHECK_ARG(a->table && a->meta && a->touched && a->counters && a->segs, "%s: NULL owner-side pointer", who);
    CF_CHECK_ARG(a->d > 0 && a->ld >= a->d && a->ld % 4 == 0 && a->ld <= 512, "%s: bad d / ld", who);
    CF_CHECK_ARG(a->optimizer == CF_OPT_SGD || a->acc, "%s: Adagrad needs the accumulator table", who);
    CF_CHECK_ARG(a->n_rows > 0, "%s: empty shard", who);
    if (a->model == CF_MODEL_CML) CF_CHECK_ARG(a->clip_norm > 0.f, "%s: CML needs clip_norm > 0", who);
    return 0;
  }
  CF_CHECK_ARG(a->table && a->meta && a->slot && a->slot_row && a->staging && a->counters && a->segs, "%s: NULL owner-side pointer", who);
  CF_CHECK_ARG(a->d > 0 && a->ld >= a->d && a->ld % 4 == 0 && a->ld <= 512, "%s: bad d / ld", who);
  CF_CHECK_ARG(a->optimizer == CF_OPT_SGD || a->acc, "%s: Adagrad needs the accumulator table", who);
  CF_CHECK_ARG(a->n_rows > 0 && a->staging_rows > 0, "%s: empty shard or staging", who);
  for (int p = 0; p < a->n_ranks; ++p) CF_CHECK_ARG(a->grads[p] != nullptr, "%s: gradient buffer of rank %d is NULL", who, p);
  if (a->model == CF_MODEL_CML) CF_CHECK_ARG(a->clip_norm > 0.f, "%s: CML needs clip_norm > 0", who);
  return 0;
}

}  // namespace

extern "C" int cf_exchange_route(const cf_exchange_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_common(a, "cf_exchange_route")) return rc;
  CF_CHECK_ARG(a->pairs && a->negs && a->B > 0 && a->W >= 1, "cf_exchange_route: pairs / negs / B / W");
  CF_CHECK_ARG(a->slot_of && a->slot_pairs && a->slot_negs && a->slot_pos && a->counters, "cf_exchange_route: NULL output");
  RouteDev R = {};
  R.pairs = a->pairs; R.negs = a->negs; R.B = a->B; R.W = a->W; R.P = a->n_ranks;
  R.n_items = a->n_items_global; R.L = (a->n_items_global + a->n_ranks - 1) / a->n_ranks; R.cap = a->cap;
  R.slot_of = a->slot_of; R.counts = a->counts[a->rank]; R.req = a->req[a->rank];
  R.slot_pairs = a->slot_pairs; R.slot_negs = a->slot_negs; R.slot_pos = a->slot_pos; R.flags = a->counters + 1;
  const long long total = (long long)a->B * (1 + a->W);
  long long grid = (total + 255) / 256;
  const long long cap = (long long)cf_num_sms() * 8;
  if (grid > cap) grid = cap;
  CF_CUDA_OK(cudaMemsetAsync(R.counts, 0, sizeof(int32_t) * CF_MAX_PEERS, stream));
  k_route_assign<<<(unsigned)grid, 256, 0, stream>>>(R);
  k_route_fill<<<(unsigned)grid, 256, 0, stream>>>(R);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cf_exchange_prepare(const cf_exchange_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_common(a, "cf_exchange_prepare")) return rc;
  if (int rc = check_owner(a, "cf_exchange_prepare")) return rc;
  CF_CHECK_ARG(a->slot_of != nullptr, "cf_exchange_prepare: slot_of is NULL");
  if (a->fetched) for (int p = 0; p < a->n_ranks; ++p) CF_CHECK_ARG(a->tables[p] != nullptr, "cf_exchange_prepare: item shard of rank %d is NULL", p);
  const XDev X = make_xdev(a);
  const StepDev P = make_owner_dev(a);
  if (a->dense_grads) CF_CUDA_OK(cudaMemsetAsync(a->segs + 3 * a->n_ranks + 2, 0, sizeof(int64_t), stream));   // length of the touched list
  k_owner_segs<<<1, CF_MAX_PEERS * CF_MAX_PEERS, 0, stream>>>(X);
  const int sms = cf_num_sms();
  const int ga = sms * 8, gb = sms * 4;
  k_exchange_prepare<<<ga + gb, 256, 0, stream>>>(X, P, ga);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cf_exchange_apply(const cf_exchange_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_common(a, "cf_exchange_apply")) return rc;
  if (int rc = check_owner(a, "cf_exchange_apply")) return rc;
  const XDev X = make_xdev(a);
  const StepDev P = make_owner_dev(a);
  const int sms = cf_num_sms();
  const int grid = sms * 16;
  if (a->dense_grads) {
    if (P.nvec <= 8) k_owner_apply_dense<8, 1><<<grid, 256, 0, stream>>>(X, P);
    else if (P.nvec <= 16) k_owner_apply_dense<16, 1><<<grid, 256, 0, stream>>>(X, P);
    else if (P.nvec <= 32) k_owner_apply_dense<32, 1><<<grid, 256, 0, stream>>>(X, P);
    else if (P.nvec <= 64) k_owner_apply_dense<32, 2><<<grid, 256, 0, stream>>>(X, P);
    else k_owner_apply_dense<32, 4><<<grid, 256, 0, stream>>>(X, P);
    CF_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (P.nvec <= 8) k_owner_scatter<8, 1><<<grid, 256, 0, stream>>>(X, P);
  else if (P.nvec <= 16) k_owner_scatter<16, 1><<<grid, 256, 0, stream>>>(X, P);
  else if (P.nvec <= 32) k_owner_scatter<32, 1><<<grid, 256, 0, stream>>>(X, P);
  else if (P.nvec <= 64) k_owner_scatter<32, 2><<<grid, 256, 0, stream>>>(X, P);
  else k_owner_scatter<32, 4><<<grid, 256, 0, stream>>>(X, P);
  long long agrid = (P.n_occ + 255) / 256;
  if (agrid > (long long)sms * 16) agrid = (long long)sms * 16;
  cf_pick_apply_kernel(P.nvec)<<<(unsigned)agrid, 256, 0, stream>>>(P);
  CF_CUDA_OK(cudaGetLastError());
  return 0;
}
