// Specialised forms of the fused minibatch step (cf_step_impl.cuh: k_step) for the cases where the generic kernel is bound by
// its instruction count rather than by memory: one negative per pair (BPRMF's reference setting, testbprmf.py:30, and
// configs[4]) and GBPR with 5 negatives and a group of 3 or 1 (configs[2]; testgbprmf.py:23-32), minibatch-synchronous, single
// GPU, rows of at most 128 floats.  Same arithmetic, operation for operation, as the generic kernel (every row gradient and
// every applied row is bit-identical for rows that occur once in the minibatch; duplicated rows differ only by the order of
// their red.adds, as they do from run to run in the generic kernel), but
//   * the numbers of negatives and group users are template parameters: the loops over slots are unrolled and the per-slot
//     role / row / coefficient broadcasts of the generic kernel disappear (ncu: the generic BPR W=1 kernel issues 806 warp
//     instructions per pair, 15 % of them floating point, and sits at 66 % of the SM's issue rate);
//   * the loads of a pair are software-pipelined: while pair n is computed, the row ids of pair n+2 and the occurrence
//     words of pair n+1 are in flight, and (NBUF = 2) so are the staging slots and the parameter / accumulator rows of pair
//     n+1, cp.async-ed into the second half of the group's shared-memory staging.  The generic kernel walks ids -> occurrence
//     words -> rows as three dependent round trips per pair.
// Prefetching rows is safe in SYNC mode: k_step only ever writes rows that occur ONCE in the minibatch (by their one
// occurrence), so no pair reads a row another pair of the same launch writes.
// CML with five negatives was measured in this form too (configs[1]: 1.999 ms with one buffer at 64 registers, 2.018 ms with two
// buffers at 2 blocks per SM, generic 1.932 ms): with seven 512-byte rows per pair the generic kernel is bound by DRAM traffic,
// not by issue rate, so it stays.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

#include "cf_step_impl.cuh"

namespace cfstep {

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_pending() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// slots of a pair: 0 = user, 1 = positive item, 2 .. 2+W-1 = negatives, 2+W .. = group users (GBPR); lane s of the group keeps
// slot s's row id, occurrence word, staging slot (and item bias)
template <int MODEL, int W, int G, int LPG, int NBUF, int MINB, int THREADS = 256>
__global__ void __launch_bounds__(THREADS, MINB) k_step_fast(const __grid_constant__ StepDev P) {
  static_assert(MODEL == CF_MODEL_BPR || MODEL == CF_MODEL_CML || MODEL == CF_MODEL_GBPR, "BPR / CML / GBPR");
  static_assert(MODEL == CF_MODEL_GBPR || G == 0, "only GBPR has group users");
  constexpr int NS = 2 + W + G;
  static_assert(NS <= LPG, "one lane per slot");
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31;
  const int gl = lane & (LPG - 1);
  const int leader = lane & ~(LPG - 1);
  const unsigned gmask = LPG == 32 ? 0xffffffffu : (((1u << LPG) - 1u) << leader);
  const int ld = P.ld;
  const bool act = gl < P.nvec;
  const bool adagrad = P.optimizer == CF_OPT_ADAGRAD;
  const bool want_loss = P.loss != nullptr;
  const long long ngroups = (long long)gridDim.x * (blockDim.x / LPG);
  const float creg = (MODEL == CF_MODEL_CML) ? (P.reg > 0.f ? P.reg : 0.f) : P.reg;
  float* wbase = smem + (size_t)(threadIdx.x / LPG) * (NBUF * 2 * NS) * ld;   // [NBUF][param rows NS | accumulator rows NS][ld]
  if (__ldcg(P.counters + 1) & (CF_FLAG_INDEX_RANGE | CF_FLAG_STAGING_FULL)) return;   // (k_count validated every id)
  const bool my_utab = gl == 0 || gl >= 2 + W;                       // my slot's row lives in the user table
  const bool my_item = gl >= 1 && gl < 2 + W;

  auto load_id = [&](long long b) -> int {
    if (gl >= NS || b >= P.B) return -1;
    if (gl == 0) return __ldg(P.pairs + 2 * b);
    if (gl == 1) return __ldg(P.pairs + 2 * b + 1);
    if (gl < 2 + W) return __ldg(P.negs + b * W + (gl - 2));
    return __ldg(P.group + b * G + (gl - 2 - W));
  };
  auto load_occ = [&](int row) -> unsigned { return row >= 0 ? __ldcg((my_utab ? P.metaU : P.metaV) + row) : 0u; };
  auto load_slot = [&](int row, unsigned occ) -> int { return occ > 1u ? __ldcg((my_utab ? P.slotU : P.slotV) + row) : 0; };
  auto load_bias = [&](int row) -> float { return (MODEL == CF_MODEL_GBPR && my_item && row >= 0) ? __ldcg(P.b + row) : 0.f; };
  auto issue_rows = [&](float* buf, int row, unsigned occ) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const long long r = __shfl_sync(gmask, row, leader + s);
      const bool utab = s == 0 || s >= 2 + W;
      if (act) cp_async16(buf + s * ld + 4 * gl, (utab ? P.U : P.V) + r * ld + 4 * gl);
    }
    if (adagrad) {
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const long long r = __shfl_sync(gmask, row, leader + s);
        const unsigned o = __shfl_sync(gmask, occ, leader + s);
        const bool utab = s == 0 || s >= 2 + W;
        if (o <= 1u && act) cp_async16(buf + (NS + s) * ld + 4 * gl, (utab ? P.accU : P.accV) + r * ld + 4 * gl);
      }
    }
  };

  double loss_acc = 0.0;
  long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPG;
  long long b1 = b + ngroups;
  int row = load_id(b), row1 = load_id(b1);
  unsigned occ = load_occ(row);
  int slot = 0;
  float bias = 0.f;
  if (NBUF == 2) {
    if (b < P.B) issue_rows(wbase, row, occ);
    cp_async_commit();
    slot = load_slot(row, occ);
    bias = load_bias(row);
  }
  unsigned occ1 = load_occ(row1);
  int n = 0;
  for (; b < P.B; b = b1, b1 += ngroups) {
    float* cur = wbase + (NBUF == 2 ? (size_t)n * (2 * NS) * ld : 0);
    const int row2 = load_id(b1 + ngroups);
    int slot1 = 0;
    float bias1 = 0.f;
    if (NBUF == 2) {
      if (b1 < P.B) issue_rows(wbase + (size_t)(n ^ 1) * (2 * NS) * ld, row1, occ1);
      cp_async_commit();
      slot1 = load_slot(row1, occ1);
      bias1 = load_bias(row1);
      cp_async_wait_pending<1>();
      n ^= 1;
    } else {
      issue_rows(cur, row, occ);
      cp_async_commit();
      slot = load_slot(row, occ);
      bias = load_bias(row);
      cp_async_wait_pending<0>();
    }
    __syncwarp(gmask);

    // ---------------------------------------------------------------- forward
    const Row<1> Uu = smem_row<LPG, 1>(cur, P.nvec, gl);
    const Row<1> Vi = smem_row<LPG, 1>(cur + ld, P.nvec, gl);
    float sq = 0.f, bsq = 0.f, lossv = 0.f;
    if (want_loss) sq = dotp<1>(Uu, Uu) + dotp<1>(Vi, Vi);
    float S = 0.f, coef = 0.f, my_gb = 0.f;
    int wmin = -1;
    float alpha[W];
    Row<1> XA = zero_row<1>();   // BPR / GBPR: -sum_w s_w V_jw; CML: the closest negative
    Row<1> XB = zero_row<1>();   // GBPR: sum_g U_g
    if constexpr (MODEL == CF_MODEL_BPR) {
      const float dui = group_sum<LPG>(dotp<1>(Uu, Vi), gmask);
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const Row<1> Vj = smem_row<LPG, 1>(cur + (2 + w) * ld, P.nvec, gl);
        const float x = dui - group_sum<LPG>(dotp<1>(Uu, Vj), gmask);   // bprmf.py:68-70
        const float sw = sigm1(x);
        S += sw;
        axpy<1>(XA, -sw, Vj);
        alpha[w] = -sw;
        if (want_loss) { lossv += softplus_neg(x); sq += dotp<1>(Vj, Vj); }
      }
    } else if constexpr (MODEL == CF_MODEL_CML) {
      const float dp = group_sum<LPG>(sqdp<1>(Uu, Vi), gmask);          // cml.py:63-85
      float dmin = INFINITY;
      int imp = 0;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const Row<1> Vj = smem_row<LPG, 1>(cur + (2 + w) * ld, P.nvec, gl);
        const float dn = group_sum<LPG>(sqdp<1>(Uu, Vj), gmask);
        if (dn < dmin) { dmin = dn; wmin = w; }
        imp += ((dp - dn) + P.margin) > 0.f;
        if (want_loss) sq += dotp<1>(Vj, Vj);
      }
      const float h = (dp - dmin) + P.margin;
      const float omega = P.use_rank_weight ? __logf(((float)imp / (float)W) * (float)P.rank_items + 1.f) : 1.f;
      coef = h > 0.f ? 2.f * omega : 0.f;
      if (want_loss) lossv = fmaxf(h, 0.f) * omega;
      if (wmin >= 0) XA = smem_row<LPG, 1>(cur + (2 + wmin) * ld, P.nvec, gl);   // reduce_min's gradient goes to the closest negative
#pragma unroll
      for (int w = 0; w < W; ++w) alpha[w] = (w == wmin) ? coef : 0.f;
    } else {
      const float dui = group_sum<LPG>(dotp<1>(Uu, Vi), gmask);         // gbprmf.py:66-88
      const float bi = __shfl_sync(gmask, bias, leader + 1);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const Row<1> Ug = smem_row<LPG, 1>(cur + (2 + W + g) * ld, P.nvec, gl);
        axpy<1>(XB, 1.f, Ug);
        if (want_loss) sq += dotp<1>(Ug, Ug);
      }
      const float ui_g = group_sum<LPG>(dotp<1>(XB, Vi), gmask) / (float)G;
      const float ui = P.rho * ui_g + (1.f - P.rho) * dui + bi;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const Row<1> Vj = smem_row<LPG, 1>(cur + (2 + w) * ld, P.nvec, gl);
        const float bj = __shfl_sync(gmask, bias, leader + 2 + w);
        const float x = ui - (group_sum<LPG>(dotp<1>(Uu, Vj), gmask) + bj);
        const float sw = sigm1(x);
        S += sw;
        axpy<1>(XA, -sw, Vj);
        alpha[w] = -sw;
        if (gl == 2 + w) my_gb = fmaf(P.reg, bj, -sw);
        if (want_loss) { lossv += softplus_neg(x); bsq += bj * bj; }
      }
      if (gl == 1) my_gb = S;
    }

    // ---------------------------------------------------------------- commit, slot by slot
    const float cg = (MODEL == CF_MODEL_GBPR) ? P.rho * S / (float)(G > 0 ? G : 1) : 0.f;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const long long r = __shfl_sync(gmask, row, leader + s);
      const unsigned o = __shfl_sync(gmask, occ, leader + s);
      const bool utab = s == 0 || s >= 2 + W;
      const Row<1> cs = s == 0 ? Uu : s == 1 ? Vi : smem_row<LPG, 1>(cur + s * ld, P.nvec, gl);
      const float a = alpha[(s >= 2 && s < 2 + W) ? s - 2 : 0];
      Row<1> g;
      if constexpr (MODEL == CF_MODEL_BPR) {
        if (s == 0) { g = lin2<1>(S, Vi, creg, cs); axpy<1>(g, 1.f, XA); }        // S V_i - sum s V_j + reg U
        else if (s == 1) g = lin2<1>(S, Uu, creg, cs);                             // S U + reg V_i
        else g = lin2<1>(a, Uu, creg, cs);                                         // -s U + reg V_j
      } else if constexpr (MODEL == CF_MODEL_CML) {
        if (s == 0) { g = lin2<1>(coef, XA, creg, cs); axpy<1>(g, -coef, Vi); }   // coef (V_j* - V_i) + c U
        else if (s == 1) g = lin2<1>(-coef, Uu, creg + coef, cs);                  // -coef (U - V_i) + c V_i
        else g = lin2<1>(a, Uu, creg - a, cs);                                     // tie (U - V_j) + c V_j
      } else {
        if (s == 0) { g = lin2<1>((1.f - P.rho) * S, Vi, creg, cs); axpy<1>(g, 1.f, XA); }
        else if (s == 1) { g = lin2<1>((1.f - P.rho) * S, Uu, creg, cs); axpy<1>(g, cg, XB); }
        else if (s < 2 + W) g = lin2<1>(a, Uu, 0.f, cs);                           // -s U          (no L2 on V_j)
        else g = lin2<1>(cg, Vi, creg, cs);                                        // rho/G S V_i + reg U_g
      }
      if (o <= 1u) {   // the row occurs once in the minibatch: update it from registers / shared memory
        Row<1> acc = smem_row<LPG, 1>(cur + (NS + s) * ld, P.nvec, gl, 1.f), p;
        apply_math<LPG, 1>(P, cs, acc, g, p, gmask);
        if (adagrad) store_row<LPG, 1>(utab ? P.accU : P.accV, r, ld, P.nvec, gl, acc);
        store_row<LPG, 1>(utab ? P.U : P.V, r, ld, P.nvec, gl, p);
      } else {         // duplicated row: sum in its staging slot, k_apply_staged applies the sum once
        const int sl = __shfl_sync(gmask, slot, leader + s);
        if (act) atomicAdd(reinterpret_cast<float4*>(P.staging + (long long)sl * P.lds + 4 * gl), g.v[0]);
      }
    }
    if constexpr (MODEL == CF_MODEL_GBPR) {   // item bias: one lane per item slot
      if (my_item) {
        if (occ <= 1u) apply_bias(P, row, bias, my_gb);
        else atomicAdd(P.staging + (long long)slot * P.lds + ld, my_gb);
      }
    }
    if (gl < NS && occ <= 1u) __stcg((my_utab ? P.metaU : P.metaV) + row, 0u);   // unique rows are done
    if (want_loss) {
      const float regsq = group_sum<LPG>(sq, gmask) + bsq;
      if (gl == 0) loss_acc += (double)(lossv + 0.5f * creg * regsq);
    }
    __syncwarp(gmask);   // this pair's shared-memory reads are done before the buffer is refilled
    row = row1; occ = occ1; row1 = row2;
    if (NBUF == 2) { slot = slot1; bias = bias1; }
    occ1 = load_occ(row1);
  }

  if (want_loss) {
    // (groups of a warp may leave the loop at different iterations: reconverge before the full-warp reduction)
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    if (lane == 0 && loss_acc != 0.0) atomicAdd(P.loss, loss_acc);
  }
}

}  // namespace cfstep

// the specialised kernel for (model, W, G, lanes per pair), or NULL; *nbuf = shared-memory row buffers per group,
// *slots = rows per buffer half, *threads = block size
cfstep::step_kernel_t cf_step_pick_fast(int model, int W, int G, int lpg, int* nbuf, int* slots, int* threads) {
  using namespace cfstep;
  *slots = 2 + W + (model == CF_MODEL_GBPR ? G : 0);
  *threads = 256;
  if (model == CF_MODEL_BPR && W == 1) {
    *nbuf = 2;
    return lpg == 32 ? k_step_fast<CF_MODEL_BPR, 1, 0, 32, 2, 4> : lpg == 16 ? k_step_fast<CF_MODEL_BPR, 1, 0, 16, 2, 4> : nullptr;
  }
  if (model == CF_MODEL_CML && W == 1) {
    *nbuf = 2;
    return lpg == 32 ? k_step_fast<CF_MODEL_CML, 1, 0, 32, 2, 4> : lpg == 16 ? k_step_fast<CF_MODEL_CML, 1, 0, 16, 2, 4> : nullptr;
  }
  if (model == CF_MODEL_GBPR && W == 5 && (G == 3 || G == 1)) {
    // one row buffer per group.  Two buffers (one block of 16 groups per SM, the next pair's twenty rows in flight during the
    // compute) were measured on configs[2]: 2.22 ms against 1.18 ms -- the tables are L2-resident there and the step lives on
    // warps, not on bytes in flight (ncu of the 256-thread form: 16 warps per SM, "wait" and shared-memory dependencies lead the
    // stalls).  Hence blocks of 128 threads at <= 102 registers: five blocks = 20 warps per SM.
    *nbuf = 1;
    *threads = 128;
    if (const char* e = getenv("CF_STEP_FAST_T256")) if (atoi(e) > 0) *threads = 256;   // A/B knob
    if (*threads == 256) {
      if (G == 3) return lpg == 32 ? k_step_fast<CF_MODEL_GBPR, 5, 3, 32, 1, 2> : lpg == 16 ? k_step_fast<CF_MODEL_GBPR, 5, 3, 16, 1, 2> : nullptr;
      return lpg == 32 ? k_step_fast<CF_MODEL_GBPR, 5, 1, 32, 1, 2> : lpg == 16 ? k_step_fast<CF_MODEL_GBPR, 5, 1, 16, 1, 2> : nullptr;
    }
    if (G == 3) return lpg == 32 ? k_step_fast<CF_MODEL_GBPR, 5, 3, 32, 1, 5, 128> : lpg == 16 ? k_step_fast<CF_MODEL_GBPR, 5, 3, 16, 1, 5, 128> : nullptr;
    return lpg == 32 ? k_step_fast<CF_MODEL_GBPR, 5, 1, 32, 1, 5, 128> : lpg == 16 ? k_step_fast<CF_MODEL_GBPR, 5, 1, 16, 1, 5, 128> : nullptr;
  }
  return nullptr;
}
