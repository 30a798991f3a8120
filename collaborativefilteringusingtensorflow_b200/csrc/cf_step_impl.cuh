// Device side of the fused minibatch step (included by the generated instantiation units and by cf_step.cu).
//
// One GROUP of LPG lanes (8 / 16 / 32, by row width) owns one (user, item) pair at a time.  Lane s of the group is
// the book-keeper of row SLOT s of the pair: slot 0 = the user row, slot 1 = the positive item row, slots 2.. = a
// tile of up to T "entries" (the W negatives first, then the G group users of GBPR).  Per pair:
//   1. every lane loads its slot's row id and (SYNC mode) the row's occurrence word            -- parallel 4/8-byte loads
//   2. the group issues cp.async.cg (16 B per lane, L2-coherent, no registers) for every slot's parameter row and,
//      when the row will be applied by this group (unique row or HOGWILD), its Adagrad accumulator row, straight
//      into the group's shared-memory staging: all 2(2+T) rows of the pair are in flight together
//   3. scores / distances from shared memory (warp-shuffle reductions), per-slot gradient coefficients kept by the
//      slot's lane
//   4. a rolled loop over the slots forms each row gradient and either applies it from registers (unique row: read
//      param + acc, write param + acc -- the algorithmic minimum) or red.adds it into the row's staging slot
//   5. rows that occur more than once in the minibatch are NOT written here: k_apply_staged (next launch) applies each
//      summed gradient once.  The kernel boundary is the only synchronisation (no fences, no spin, no done counters:
//      an in-kernel last-arriver protocol with __threadfence() cost 2x at B = 16k).
// All loops over slots are rolled (the slot's data lives in shared memory / its lane), so the kernel is a few KB of
// SASS instead of the 80-250 KB of the first, fully unrolled version (which stalled on instruction fetch).
#pragma once
#include <math.h>

#include "common.cuh"

namespace cfstep {

struct StepDev {
  float *U, *V, *b, *accU, *accV, *accb;
  long long n_users, n_items;
  int d, ld, nvec;
  const int32_t *pairs, *negs, *group;
  const float* ratings;
  int B, W, G, T;  // T = entries (negatives + group users) staged per tile
  int model, optimizer, update, use_rank_weight;
  float lr, reg, margin, clip, rho, weight;
  unsigned int *metaU, *metaV;     // occurrences of each row in the current minibatch (0 between minibatches)
  int32_t *slotU, *slotV;          // staging slot of a row that occurs more than once
  uint32_t* slot_row;              // inverse map: slot -> row id | (item table ? 1u << 31 : 0)
  float* staging;
  long long staging_rows;
  int lds;
  int32_t* counters;
  double* loss;
  float* gradV;          // exchange mode (multi-GPU): V/n_items describe FETCHED item rows; item-row gradients are
                         // red.added into gradV[row] (stride ld) instead of being applied here
  long long rank_items;  // CML rank weight uses the GLOBAL item count
  float *gradU, *gradb;  // replicated data-parallel mode: user-row (and GBPR bias) gradients are red.added into dense
                         // tables as well (with gradV = [n_items, ld]); nothing is applied by the step kernels
  // peer-pull variant of the exchange mode: item ids are GLOBAL, row i lives at peerV[i % n_peers] + (i / n_peers) * ld
  // (the owner's shard, mapped over NVLink), its gradient goes to gradV[gslot_*]
  const float* peerV[CF_MAX_PEERS];
  const int32_t *gslot_pos, *gslot_neg;
  int n_peers;
  // push variant of the peer mode: the gradient of item i is red.added straight into its OWNER's dense gradient table
  // peerG[i % n_peers] + (i / n_peers) * ld (NVLink, outbound while the row reads are inbound); gslot_* are not used
  float* peerG[CF_MAX_PEERS];
  long long n_occ;       // slots scanned by k_apply_staged
};

template <int NV>
struct Row {
  float4 v[NV];
};

__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float sqd4(float4 a, float4 b) {
  const float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z, w = a.w - b.w;
  return x * x + y * y + z * z + w * w;
}

template <int LPG, int NV>
__device__ __forceinline__ Row<NV> load_row(const float* tab, long long r, int ld, int nvec, int gl, float fill = 0.f) {
  Row<NV> x;
  const float* p = tab + r * (long long)ld;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = gl + k * LPG;
    x.v[k] = (v < nvec) ? ldcg4(p + 4 * v) : make_float4(fill, fill, fill, fill);
  }
  return x;
}
template <int LPG, int NV>
__device__ __forceinline__ void store_row(float* tab, long long r, int ld, int nvec, int gl, const Row<NV>& x) {
  float* p = tab + r * (long long)ld;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = gl + k * LPG;
    if (v < nvec) stcg4(p + 4 * v, x.v[k]);
  }
}
// shared-memory row (stride ld floats); lanes past the row read zeros / `fill`
template <int LPG, int NV>
__device__ __forceinline__ Row<NV> smem_row(const float* s, int nvec, int gl, float fill = 0.f) {
  Row<NV> x;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = gl + k * LPG;
    x.v[k] = (v < nvec) ? *reinterpret_cast<const float4*>(s + 4 * v) : make_float4(fill, fill, fill, fill);
  }
  return x;
}
template <int NV>
__device__ __forceinline__ float dotp(const Row<NV>& a, const Row<NV>& b) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) s += dot4(a.v[k], b.v[k]);
  return s;
}
template <int NV>
__device__ __forceinline__ float sqdp(const Row<NV>& a, const Row<NV>& b) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) s += sqd4(a.v[k], b.v[k]);
  return s;
}
template <int NV>
__device__ __forceinline__ void axpy(Row<NV>& y, float s, const Row<NV>& x) {  // y += s * x
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    y.v[k].x = fmaf(s, x.v[k].x, y.v[k].x);
    y.v[k].y = fmaf(s, x.v[k].y, y.v[k].y);
    y.v[k].z = fmaf(s, x.v[k].z, y.v[k].z);
    y.v[k].w = fmaf(s, x.v[k].w, y.v[k].w);
  }
}
template <int NV>
__device__ __forceinline__ Row<NV> lin2(float a, const Row<NV>& x, float b, const Row<NV>& y) {  // a*x + b*y
  Row<NV> r;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    r.v[k].x = fmaf(a, x.v[k].x, b * y.v[k].x);
    r.v[k].y = fmaf(a, x.v[k].y, b * y.v[k].y);
    r.v[k].z = fmaf(a, x.v[k].z, b * y.v[k].z);
    r.v[k].w = fmaf(a, x.v[k].w, b * y.v[k].w);
  }
  return r;
}
template <int NV>
__device__ __forceinline__ Row<NV> zero_row() {
  Row<NV> r;
#pragma unroll
  for (int k = 0; k < NV; ++k) r.v[k] = f4zero();
  return r;
}
template <int LPG>
__device__ __forceinline__ float group_sum(float x, unsigned gmask) {
#pragma unroll
  for (int o = LPG / 2; o > 0; o >>= 1) x += __shfl_xor_sync(gmask, x, o);
  return x;
}
__device__ __forceinline__ float softplus_neg(float x) {  // -log(sigmoid(x)), bprmf.py:70
  return x > 0.f ? log1pf(__expf(-x)) : (-x + log1pf(__expf(x)));
}
__device__ __forceinline__ float sigm1(float x) { return -1.f / (1.f + expf(x)); }  // sigmoid(x) - 1
__device__ __forceinline__ bool in_range(long long r, long long n) { return r >= 0 && r < n; }

// where item row r is read from: the local table, the fetched rows (exchange mode) or its owner's shard (peer pull)
template <bool EXT>
__device__ __forceinline__ const float* item_row_ptr(const StepDev& P, int r) {
  if (EXT && P.n_peers > 0) {
    const int q = r / P.n_peers;
    return P.peerV[r - q * P.n_peers] + (long long)q * P.ld;
  }
  return P.V + (long long)r * P.ld;
}

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
template <int LPG, int NV>
__device__ __forceinline__ void stage_row(float* smem_dst, const float* tab, long long r, int ld, int nvec, int gl) {
  const float* p = tab + r * (long long)ld;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = gl + k * LPG;
    if (v < nvec) cp_async16(smem_dst + 4 * v, p + 4 * v);
  }
}

// new parameter row (and accumulator row) from the current row, its accumulator and the summed gradient;
// the CML unit-norm clip (cml.py:119-122) of the updated row is fused in
template <int LPG, int NV>
__device__ __forceinline__ void apply_math(const StepDev& P, const Row<NV>& cur, Row<NV>& acc, const Row<NV>& g, Row<NV>& p,
                                           unsigned gmask) {
  if (P.optimizer == CF_OPT_ADAGRAD) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      acc.v[k].x = fmaf(g.v[k].x, g.v[k].x, acc.v[k].x);
      acc.v[k].y = fmaf(g.v[k].y, g.v[k].y, acc.v[k].y);
      acc.v[k].z = fmaf(g.v[k].z, g.v[k].z, acc.v[k].z);
      acc.v[k].w = fmaf(g.v[k].w, g.v[k].w, acc.v[k].w);
      p.v[k].x = fmaf(-P.lr * g.v[k].x, rsqrtf(acc.v[k].x), cur.v[k].x);
      p.v[k].y = fmaf(-P.lr * g.v[k].y, rsqrtf(acc.v[k].y), cur.v[k].y);
      p.v[k].z = fmaf(-P.lr * g.v[k].z, rsqrtf(acc.v[k].z), cur.v[k].z);
      p.v[k].w = fmaf(-P.lr * g.v[k].w, rsqrtf(acc.v[k].w), cur.v[k].w);
    }
  } else {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      p.v[k].x = fmaf(-P.lr, g.v[k].x, cur.v[k].x);
      p.v[k].y = fmaf(-P.lr, g.v[k].y, cur.v[k].y);
      p.v[k].z = fmaf(-P.lr, g.v[k].z, cur.v[k].z);
      p.v[k].w = fmaf(-P.lr, g.v[k].w, cur.v[k].w);
    }
  }
  if (P.model == CF_MODEL_CML) {
    const float n2 = group_sum<LPG>(dotp<NV>(p, p), gmask);
    if (n2 > P.clip * P.clip) {
      const float sc = P.clip * rsqrtf(n2);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        p.v[k].x *= sc; p.v[k].y *= sc; p.v[k].z *= sc; p.v[k].w *= sc;
      }
    }
  }
}

__device__ __forceinline__ void apply_bias(const StepDev& P, long long r, float cur, float g) {
  if (P.optimizer == CF_OPT_ADAGRAD) {
    const float a = fmaf(g, g, __ldcg(P.accb + r));
    __stcg(P.accb + r, a);
    __stcg(P.b + r, fmaf(-P.lr * g, rsqrtf(a), cur));
  } else {
    __stcg(P.b + r, fmaf(-P.lr, g, cur));
  }
}

enum { ROLE_NONE = 0, ROLE_USER = 1, ROLE_ITEM = 2, ROLE_NEG = 3, ROLE_GROUP = 4 };

// EXT = false is the single-GPU kernel; EXT = true adds the multi-GPU variants (fetched rows / peer pull / dense gradient
// tables).  They are compiled apart because the single-GPU d=128 kernels sit exactly at 64 registers without spills (4
// blocks per SM); the extra live values of the exchange paths cost a block of occupancy (-35 % measured on configs[1]),
// and forcing 64 registers with __launch_bounds__(256, 4) makes the compiler re-load parameters everywhere (-18 %).
// Register allocation at that edge is fragile: re-check `cuobjdump -res-usage` (REG:64 STACK:0 for k_step<*,32,1,false>)
// after touching this kernel.
template <int MODEL, int LPG, int NV, bool EXT>
__global__ void __launch_bounds__(256) k_step(const __grid_constant__ StepDev P) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31;
  const int gl = lane & (LPG - 1);
  const int leader = lane & ~(LPG - 1);
  const unsigned gmask = LPG == 32 ? 0xffffffffu : (((1u << LPG) - 1u) << leader);
  const long long ngroups = (long long)gridDim.x * blockDim.x / LPG;
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPG;
  const bool sync = P.update == CF_UPDATE_SYNC;
  const bool adagrad = P.optimizer == CF_OPT_ADAGRAD;
  const bool want_loss = P.loss != nullptr;
  const bool item_ext = P.gradV != nullptr;   // (runtime also in the single-GPU kernel: ptxas allocates this form in 64 registers without spills)
  const bool user_ext = EXT && P.gradU != nullptr;
  const bool pull = EXT && P.n_peers > 0;
  const int nslot = 2 + P.T;
  float* sp = smem + (size_t)(threadIdx.x / LPG) * (2 * nslot) * P.ld;  // parameter rows of this group's slots
  float* sa = sp + (size_t)nslot * P.ld;                                // accumulator rows
  const int E = (MODEL == CF_MODEL_WRMF) ? 0 : P.W + ((MODEL == CF_MODEL_GBPR) ? P.G : 0);
  const bool single = E <= P.T;
  const float creg = (MODEL == CF_MODEL_CML) ? (P.reg > 0.f ? P.reg : 0.f) : P.reg;
  if (sync && (__ldcg(P.counters + 1) & (CF_FLAG_INDEX_RANGE | CF_FLAG_STAGING_FULL))) return;

  double loss_acc = 0.0;
  const long long iters = (P.B + ngroups - 1) / ngroups;
  for (long long it = 0; it < iters; ++it) {
    const long long b0 = gid + it * ngroups;
    if (b0 >= P.B) continue;  // group-uniform
    const long long bb = b0;

    // ---------------------------------------------------------------- slots 0 (user) and 1 (positive item)
    int my_row = -1, my_role = ROLE_NONE;
    if (gl == 0) { my_row = __ldg(P.pairs + 2 * bb); my_role = ROLE_USER; }
    if (gl == 1) { my_row = __ldg(P.pairs + 2 * bb + 1); my_role = ROLE_ITEM; }
    bool ok = my_role == ROLE_NONE || in_range(my_row, my_role == ROLE_USER ? P.n_users : P.n_items);
    // every entry id of the pair is validated up front (an invalid id skips the whole pair, nothing is written)
    for (int e = gl; e < E; e += LPG) {
      const bool isneg = e < P.W;
      const int r = isneg ? __ldg(P.negs + bb * P.W + e) : __ldg(P.group + bb * P.G + (e - P.W));
      ok = ok && in_range(r, isneg ? P.n_items : P.n_users);
    }
    if (!__all_sync(gmask, ok)) {
      if (gl == 0) atomicOr(P.counters + 1, CF_FLAG_INDEX_RANGE);
      continue;
    }
    const int i = __shfl_sync(gmask, my_row, leader + 1);

    // pair-level quantities carried across tiles
    float S = 0.f;                 // sum_w s_bw (BPR / GBPR)
    float dmin = INFINITY;         // CML: closest negative distance, its entry index, impostor count
    int wmin = -1, imp = 0;
    Row<NV> XA = zero_row<NV>();   // BPR/GBPR: -sum_w s_w V_jw ; CML: V_j*  (pieces of the user-row gradient)
    Row<NV> XB = zero_row<NV>();   // GBPR: sum_g U_g
    float lossv = 0.f, sq = 0.f, bsq = 0.f;
    Row<NV> Uu, Vi;
    float dui = 0.f, dp = 0.f, bi = 0.f, ui = 0.f, coef = 0.f, omega = 1.f, we = 0.f;

    // ---------------------------------------------------------------- tiles of entries
    // pass 0 (only when the entries do not fit one tile): CML needs min_w / impostors over ALL negatives and GBPR
    // needs sum_g U_g before any gradient can be formed; pass 1: gradients + commits.
    const int first_pass = (!single && (MODEL == CF_MODEL_CML || MODEL == CF_MODEL_GBPR)) ? 0 : 1;
    unsigned occ_lo = 0u;          // occurrences of MY slot's row in this minibatch (lanes 0 / 1 keep theirs across tiles)
    for (int pass = first_pass; pass < 2; ++pass) {
      const bool commit_pass = pass == 1;
      for (int e0 = 0; e0 < max(E, 1); e0 += max(P.T, 1)) {
        const int ne = min(P.T, E - e0);           // entries in this tile (0 for WRMF)
        const bool first_tile = e0 == 0, last_tile = e0 + P.T >= E;
        // ---- my slot in this tile
        if (gl >= 2) {
          my_role = ROLE_NONE;
          my_row = -1;
          const int e = e0 + gl - 2;
          if (gl - 2 < ne) {
            if (e < P.W) { my_row = __ldg(P.negs + bb * P.W + e); my_role = ROLE_NEG; }
            else { my_row = __ldg(P.group + bb * P.G + (e - P.W)); my_role = ROLE_GROUP; }
          }
        }
        const bool is_user_tab = my_role == ROLE_USER || my_role == ROLE_GROUP;
        const bool stage_ui = first_tile && pass == first_pass;      // u / i parameter rows: once per pair
        const bool meta_ui = first_tile && commit_pass;              // u / i occurrence words + accumulators: once
        const bool my_local = my_role != ROLE_NONE && (is_user_tab ? !user_ext : !item_ext);   // rows this GPU owns and applies
        if (sync && commit_pass && my_local && (gl >= 2 || meta_ui))
          occ_lo = __ldcg((is_user_tab ? P.metaU : P.metaV) + my_row);
        // ---- stage parameter rows (and the accumulators of rows this group will apply itself) into shared memory
        __syncwarp(gmask);   // the previous tile's shared-memory reads are done
        int my_slot = 0;     // staging slot of my row when it is a duplicated one (loaded while the rows are in flight)
        for (int s = stage_ui ? 0 : 2; s < 2 + ne; ++s) {
          const int r = __shfl_sync(gmask, my_row, leader + s);
          const int role = __shfl_sync(gmask, my_role, leader + s);
          if (EXT && pull && (role == ROLE_ITEM || role == ROLE_NEG))
            stage_row<LPG, NV>(sp + (size_t)s * P.ld, item_row_ptr<EXT>(P, r), 0, P.ld, P.nvec, gl);
          else
            stage_row<LPG, NV>(sp + (size_t)s * P.ld, (role == ROLE_USER || role == ROLE_GROUP) ? P.U : P.V, r, P.ld, P.nvec, gl);
        }
        if (adagrad && commit_pass) {
          for (int s = meta_ui ? 0 : 2; s < 2 + ne; ++s) {
            const unsigned occ = __shfl_sync(gmask, occ_lo, leader + s);
            const int r = __shfl_sync(gmask, my_row, leader + s);
            const int role = __shfl_sync(gmask, my_role, leader + s);
            const bool utab = role == ROLE_USER || role == ROLE_GROUP;
            if ((utab ? !user_ext : !item_ext) && (!sync || occ <= 1u))
              stage_row<LPG, NV>(sa + (size_t)s * P.ld, utab ? P.accU : P.accV, r, P.ld, P.nvec, gl);
          }
        }
        if (sync && commit_pass && my_local && (gl >= 2 || last_tile) && occ_lo > 1u)
          my_slot = __ldcg((is_user_tab ? P.slotU : P.slotV) + my_row);
        if (pull && commit_pass && P.gslot_pos != nullptr) {   // peer pull: the row of gradV that collects my item row's gradient
          if (my_role == ROLE_NEG) my_slot = __ldg(P.gslot_neg + bb * P.W + (e0 + gl - 2));
          if (my_role == ROLE_ITEM && last_tile) my_slot = __ldg(P.gslot_pos + bb);
        }
        cp_async_wait_all();
        __syncwarp(gmask);

        // ---- pair-level forward quantities (first tile of the first pass)
        if (first_tile && pass == first_pass) {
          Uu = smem_row<LPG, NV>(sp, P.nvec, gl);
          Vi = smem_row<LPG, NV>(sp + P.ld, P.nvec, gl);
          if (want_loss) sq = dotp<NV>(Uu, Uu) + dotp<NV>(Vi, Vi);
          if constexpr (MODEL == CF_MODEL_BPR) dui = group_sum<LPG>(dotp<NV>(Uu, Vi), gmask);
          if constexpr (MODEL == CF_MODEL_CML) dp = group_sum<LPG>(sqdp<NV>(Uu, Vi), gmask);
          if constexpr (MODEL == CF_MODEL_GBPR) {
            dui = group_sum<LPG>(dotp<NV>(Uu, Vi), gmask);
            bi = __ldcg(P.b + i);
          }
          if constexpr (MODEL == CF_MODEL_WRMF) {
            const float e = group_sum<LPG>(dotp<NV>(Uu, Vi), gmask) - __ldg(P.ratings + bb);
            we = P.weight * e;
            if (want_loss) lossv = 0.5f * P.weight * e * e;
          }
        }

        // ---- per-entry forward: scores / distances; the entry's lane keeps its coefficient
        float my_alpha = 0.f, my_b = 0.f, my_gb = 0.f;
        if constexpr (MODEL == CF_MODEL_GBPR) {
          if (my_role == ROLE_NEG) my_b = __ldcg(P.b + my_row);
          if (gl == 1) my_b = bi;
          // group rows first: sum_g U_g (pass 0, or the single tile)
          if (pass == first_pass) {
            for (int s = 2; s < 2 + ne; ++s) {
              if (__shfl_sync(gmask, my_role, leader + s) != ROLE_GROUP) continue;
              const Row<NV> Ug = smem_row<LPG, NV>(sp + (size_t)s * P.ld, P.nvec, gl);
              axpy<NV>(XB, 1.f, Ug);
              if (want_loss) sq += dotp<NV>(Ug, Ug);
            }
          }
          if (commit_pass && first_tile) {  // sum_g U_g is complete here (pass 0 covered every tile, or single tile)
            const float ui_g = group_sum<LPG>(dotp<NV>(XB, Vi), gmask) / (float)P.G;
            ui = P.rho * ui_g + (1.f - P.rho) * dui + bi;
          }
        }
        if (MODEL != CF_MODEL_WRMF && (commit_pass || MODEL == CF_MODEL_CML)) {
          for (int s = 2; s < 2 + ne; ++s) {
            if (__shfl_sync(gmask, my_role, leader + s) != ROLE_NEG) continue;
            const Row<NV> Vj = smem_row<LPG, NV>(sp + (size_t)s * P.ld, P.nvec, gl);
            if constexpr (MODEL == CF_MODEL_BPR) {
              const float x = dui - group_sum<LPG>(dotp<NV>(Uu, Vj), gmask);   // bprmf.py:68-70
              const float sw = sigm1(x);
              S += sw;
              axpy<NV>(XA, -sw, Vj);
              if (gl == s) my_alpha = -sw;
              if (want_loss) { lossv += softplus_neg(x); sq += dotp<NV>(Vj, Vj); }
            } else if constexpr (MODEL == CF_MODEL_CML) {
              if (pass == first_pass) {                                            // cml.py:63-82
                const float dn = group_sum<LPG>(sqdp<NV>(Uu, Vj), gmask);
                if (dn < dmin) { dmin = dn; wmin = e0 + s - 2; }
                imp += ((dp - dn) + P.margin) > 0.f;
                if (want_loss) sq += dotp<NV>(Vj, Vj);
              }
            } else if constexpr (MODEL == CF_MODEL_GBPR) {
              const float bj = __shfl_sync(gmask, my_b, leader + s);
              const float x = ui - (group_sum<LPG>(dotp<NV>(Uu, Vj), gmask) + bj);  // gbprmf.py:83-88
              const float sw = sigm1(x);
              S += sw;
              axpy<NV>(XA, -sw, Vj);
              if (gl == s) { my_alpha = -sw; my_gb = fmaf(P.reg, bj, -sw); }
              if (want_loss) { lossv += softplus_neg(x); bsq += bj * bj; }
            }
          }
        }
        if (!commit_pass) continue;

        if constexpr (MODEL == CF_MODEL_CML) {
          if (first_tile) {  // min / impostors are complete (pass 0 or single tile): hinge, rank weight (cml.py:73-85)
            const float h = (dp - dmin) + P.margin;
            omega = P.use_rank_weight ? __logf(((float)imp / (float)P.W) * (float)P.rank_items + 1.f) : 1.f;
            coef = h > 0.f ? 2.f * omega : 0.f;
            if (want_loss) lossv = fmaxf(h, 0.f) * omega;
          }
          if (my_role == ROLE_NEG && e0 + gl - 2 == wmin) my_alpha = coef;   // reduce_min grad -> the closest negative
          if (wmin >= e0 && wmin < e0 + ne) XA = smem_row<LPG, NV>(sp + (size_t)(2 + wmin - e0) * P.ld, P.nvec, gl);
        }

        // ---- commit loop over the slots of this tile (u and i ride with the last tile)
        for (int s = last_tile ? 0 : 2; s < 2 + ne; ++s) {
          const int role = __shfl_sync(gmask, my_role, leader + s);
          if (role == ROLE_NONE) continue;
          const int r = __shfl_sync(gmask, my_row, leader + s);
          const float alpha = __shfl_sync(gmask, my_alpha, leader + s);
          const unsigned occ = __shfl_sync(gmask, occ_lo, leader + s);
          const bool utab = role == ROLE_USER || role == ROLE_GROUP;
          const Row<NV> cur = smem_row<LPG, NV>(sp + (size_t)s * P.ld, P.nvec, gl);
          Row<NV> g;
          if constexpr (MODEL == CF_MODEL_BPR) {
            if (role == ROLE_NEG) g = lin2<NV>(alpha, Uu, creg, cur);                     // -s U + reg V_j
            else if (role == ROLE_ITEM) g = lin2<NV>(S, Uu, creg, cur);                    // S U + reg V_i
            else { g = lin2<NV>(S, Vi, creg, cur); axpy<NV>(g, 1.f, XA); }                 // S V_i - sum s V_j + reg U
          } else if constexpr (MODEL == CF_MODEL_CML) {
            if (role == ROLE_NEG) g = lin2<NV>(alpha, Uu, creg - alpha, cur);              // tie (U - V_j) + c V_j
            else if (role == ROLE_ITEM) g = lin2<NV>(-coef, Uu, creg + coef, cur);          // -coef (U - V_i) + c V_i
            else { g = lin2<NV>(coef, XA, creg, cur); axpy<NV>(g, -coef, Vi); }            // coef (V_j* - V_i) + c U
          } else if constexpr (MODEL == CF_MODEL_GBPR) {
            const float cg = P.rho * S / (float)P.G;
            if (role == ROLE_NEG) g = lin2<NV>(alpha, Uu, 0.f, cur);                       // -s U          (no L2 on V_j)
            else if (role == ROLE_GROUP) g = lin2<NV>(cg, Vi, creg, cur);                  // rho/G S V_i + reg U_g
            else if (role == ROLE_ITEM) { g = lin2<NV>((1.f - P.rho) * S, Uu, creg, cur); axpy<NV>(g, cg, XB); }
            else { g = lin2<NV>((1.f - P.rho) * S, Vi, creg, cur); axpy<NV>(g, 1.f, XA); }
          } else {
            g = role == ROLE_ITEM ? lin2<NV>(we, Uu, creg, cur) : lin2<NV>(we, Vi, creg, cur);
          }
          float* Tb = utab ? P.U : P.V;
          float* Ab = utab ? P.accU : P.accV;
          // a fetched (remote) / replicated row: its gradient goes to the exchange buffer (row = the id, or the slot
          // given by gslot_* in peer-pull mode).  The EXT = false form is kept exactly as ptxas likes it (see above).
          bool to_ext;
          if constexpr (EXT) to_ext = utab ? user_ext : item_ext;
          else to_ext = item_ext && !utab;
          if (to_ext) {
            int gs = r;
            if constexpr (EXT) {
              if (pull && !utab) gs = __shfl_sync(gmask, my_slot, leader + s);
            }
            float* gr = ((EXT && utab) ? P.gradU : P.gradV) + (long long)gs * P.ld;
            if constexpr (EXT) {
              if (pull && !utab && P.gslot_pos == nullptr) {   // push: the owner's dense gradient table, over NVLink
                const int q = r / P.n_peers;
                gr = P.peerG[r - q * P.n_peers] + (long long)q * P.ld;
              }
            }
#pragma unroll
            for (int k = 0; k < NV; ++k) {
              const int v = gl + k * LPG;
              if (v < P.nvec) atomicAdd(reinterpret_cast<float4*>(gr + 4 * v), g.v[k]);
            }
          } else if (!sync || occ <= 1u) {  // unique row (or racy mode): update straight from registers / shared memory
            Row<NV> acc = smem_row<LPG, NV>(sa + (size_t)s * P.ld, P.nvec, gl, 1.f), p;
            apply_math<LPG, NV>(P, cur, acc, g, p, gmask);
            if (adagrad) store_row<LPG, NV>(Ab, r, P.ld, P.nvec, gl, acc);
            store_row<LPG, NV>(Tb, r, P.ld, P.nvec, gl, p);
          } else {
            // the row occurs more than once in this minibatch: stage the gradient; k_apply_staged applies the sum
            const int slot = __shfl_sync(gmask, my_slot, leader + s);
            float* st = P.staging + (long long)slot * P.lds;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
              const int v = gl + k * LPG;
              if (v < P.nvec) atomicAdd(reinterpret_cast<float4*>(st + 4 * v), g.v[k]);
            }
          }
        }
        // ---- item bias (GBPR): one lane per item slot
        if constexpr (MODEL == CF_MODEL_GBPR) {
          if (gl == 1) my_gb = S;
          const bool item_slot = (my_role == ROLE_NEG) || (gl == 1 && last_tile);
          if (item_slot) {
            bool sent = false;
            if constexpr (EXT) {
              if (item_ext) { atomicAdd(P.gradb + my_row, my_gb); sent = true; }
            }
            if (sent) {}
            else if (!sync || occ_lo <= 1u) apply_bias(P, my_row, my_b, my_gb);
            else atomicAdd(P.staging + (long long)my_slot * P.lds + P.ld, my_gb);
          }
        }
        // ---- unique rows are done: return their occurrence word to zero (duplicated rows: k_apply_staged does it)
        if (sync && my_local && (gl >= 2 || last_tile) && occ_lo <= 1u)
          __stcg((is_user_tab ? P.metaU : P.metaV) + my_row, 0u);
      }  // tiles
    }    // passes
    if (want_loss) {
      const float regsq = group_sum<LPG>(sq, gmask) + bsq;
      if (gl == 0) loss_acc += (double)(lossv + 0.5f * creg * regsq);
    }
  }

  if (want_loss) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    if (lane == 0 && loss_acc != 0.0) atomicAdd(P.loss, loss_acc);
  }
}

// Applies the summed gradient of every row that occurred more than once in the minibatch (one group per staging slot):
// param + acc + staged gradient in, param + acc out, slot and occurrence word back to zero.  The parameter row is still
// the pre-update one (k_step never writes a duplicated row), so this is exactly TF's "sum duplicates, apply once".
#define CF_SLOT_EMPTY 0xffffffffu

template <int LPG, int NV>
__global__ void __launch_bounds__(256) k_apply_staged(const __grid_constant__ StepDev P) {
  const int lane = threadIdx.x & 31, gl = lane & (LPG - 1), leader = lane & ~(LPG - 1);
  const unsigned gmask = LPG == 32 ? 0xffffffffu : (((1u << LPG) - 1u) << leader);
  const long long ngroups = (long long)gridDim.x * blockDim.x / LPG;
  const long long n = P.n_occ;                 // one potential slot per row occurrence of the minibatch
  const bool adagrad = P.optimizer == CF_OPT_ADAGRAD;
  // each group scans LPG consecutive slot codes at a time (one coalesced load) and walks the occupied ones together
  for (long long t0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPG) * LPG; t0 < n; t0 += ngroups * LPG) {
    uint32_t code = CF_SLOT_EMPTY;
    if (t0 + gl < n) code = __ldcg(P.slot_row + t0 + gl);
    unsigned occupied = (__ballot_sync(gmask, code != CF_SLOT_EMPTY) & gmask) >> leader;
    if (code != CF_SLOT_EMPTY) __stcg(P.slot_row + t0 + gl, CF_SLOT_EMPTY);
    while (occupied) {
      const int k = __ffs(occupied) - 1;
      occupied &= occupied - 1;
      const uint32_t c = __shfl_sync(gmask, code, leader + k);
      const bool vtab = c >> 31;
      const long long r = c & 0x7fffffffu;
      float* st = P.staging + (t0 + k) * P.lds;
      const Row<NV> g = load_row<LPG, NV>(st, 0, 0, P.nvec, gl);
      const Row<NV> cur = load_row<LPG, NV>(vtab ? P.V : P.U, r, P.ld, P.nvec, gl);
      Row<NV> acc, p;
      if (adagrad) acc = load_row<LPG, NV>(vtab ? P.accV : P.accU, r, P.ld, P.nvec, gl, 1.f);
      apply_math<LPG, NV>(P, cur, acc, g, p, gmask);
      if (adagrad) store_row<LPG, NV>(vtab ? P.accV : P.accU, r, P.ld, P.nvec, gl, acc);
      store_row<LPG, NV>(vtab ? P.V : P.U, r, P.ld, P.nvec, gl, p);
      // The slot is zeroed only AFTER its gradient has been consumed: a store issued while a load of the same line is
      // still outstanding is ~4x slower on B200 (measured, tests/micro/apply_micro.cu: 264 us vs 61 us per 100k rows).
      store_row<LPG, NV>(st, 0, 0, P.nvec, gl, zero_row<NV>());
      if (gl == 0) {
        if (vtab && P.b != nullptr) {
          const float gb = __ldcg(st + P.ld);
          apply_bias(P, r, __ldcg(P.b + r), gb);
          __stcg(st + P.ld, 0.f);
        }
        __stcg((vtab ? P.metaV : P.metaU) + r, 0u);
      }
    }
  }
}

// Owner side of the multi-GPU exchange: n gradient rows (grads[k], stride ldg) for table rows rows[k] of THIS GPU's shard
// (described as the "U" table of P).  A row received once is applied straight away; a row requested by several GPUs is
// summed in its staging slot and applied by k_apply_staged -- the same "sum duplicates, apply once" rule.
// owner-pull variant: the gradient rows stay in the requesters' buffers (peer memory over NVLink); rows
// [start[p], start[p + 1]) are read from base[p], consecutive
struct GradSegs {
  const float* base[CF_MAX_PEERS];
  long long start[CF_MAX_PEERS + 1];
  long long rot;   // the rows are processed in rotated order (k + rot) mod n: every owner starts at another requester
  int n;
};

template <int LPG, int NV>
__global__ void __launch_bounds__(256) k_scatter_rows(const __grid_constant__ StepDev P, const int32_t* __restrict__ rows,
                                                      const float* __restrict__ grads, long long n, int ldg,
                                                      const __grid_constant__ GradSegs S) {
  const int lane = threadIdx.x & 31, gl = lane & (LPG - 1), leader = lane & ~(LPG - 1);
  const unsigned gmask = LPG == 32 ? 0xffffffffu : (((1u << LPG) - 1u) << leader);
  const long long ngroups = (long long)gridDim.x * blockDim.x / LPG;
  const bool adagrad = P.optimizer == CF_OPT_ADAGRAD;
  for (long long k0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPG; k0 < n; k0 += ngroups) {
    long long k = k0 + S.rot;
    if (k >= n) k -= n;
    const long long r = __ldg(rows + k);
    if (!in_range(r, P.n_users)) continue;   // flagged by the counting kernel
    const unsigned occ = __ldcg(P.metaU + r);
    const float* gsrc = grads;
    long long gk = k;
    if (S.n > 0) {
      int p = 0;
      while (p + 1 < S.n && k >= S.start[p + 1]) ++p;
      gsrc = S.base[p];
      gk = k - S.start[p];
    }
    const Row<NV> g = load_row<LPG, NV>(gsrc, gk, ldg, P.nvec, gl);
    if (occ <= 1u) {
      const Row<NV> cur = load_row<LPG, NV>(P.U, r, P.ld, P.nvec, gl);
      Row<NV> acc, p;
      if (adagrad) acc = load_row<LPG, NV>(P.accU, r, P.ld, P.nvec, gl, 1.f);
      apply_math<LPG, NV>(P, cur, acc, g, p, gmask);
      if (adagrad) store_row<LPG, NV>(P.accU, r, P.ld, P.nvec, gl, acc);
      store_row<LPG, NV>(P.U, r, P.ld, P.nvec, gl, p);
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

typedef void (*step_kernel_t)(const StepDev);
typedef void (*scatter_kernel_t)(const StepDev, const int32_t*, const float*, long long, int, const GradSegs);

}  // namespace cfstep

// one instantiation unit per (model, row shape): build.py generates build/gen/cf_step_inst_<m>_<s>.cu defining these
#define CF_STEP_PICK_DECL(M, S) cfstep::step_kernel_t cf_step_pick_##M##_##S(); cfstep::step_kernel_t cf_step_pick_ext_##M##_##S()
#define CF_APPLY_PICK_DECL(S) cfstep::step_kernel_t cf_apply_pick_##S()
#define CF_SCATTER_PICK_DECL(S) cfstep::scatter_kernel_t cf_scatter_pick_##S()
