// Specialised form of the fused minibatch step (cf_step_impl.cuh: k_step) for one negative per pair -- the reference's BPRMF
// setting (testbprmf.py:30) and configs[4] -- at 64 < ld <= 128, minibatch-synchronous, single GPU.  Same arithmetic, operation for operation, as the generic kernel (so every row gradient and every applied
// row is bit-identical for rows that occur once; duplicated rows differ only by the order of their red.adds, as they do
// from run to run in the generic kernel), but
//   * the number of negatives is a template parameter: the loops over slots are unrolled and the per-slot role / row /
//     coefficient broadcasts of the generic kernel disappear (ncu: the generic BPR W=1 kernel issues 806 warp
//     instructions per pair, 15 % of them floating point, and sits at 66 % of the SM's issue rate);
//   * the loads of a pair are software-pipelined: while pair n is computed, the row ids of pair n+2 and the occurrence
//     words and staging slots of pair n+1 are in flight, and (NBUF = 2) so are the parameter / accumulator rows of pair
//     n+1, cp.async-ed into the second half of the warp's shared-memory staging.  The generic kernel walks ids -> occurrence
//     words -> rows as three dependent round trips per pair.
// Prefetching rows is safe in SYNC mode: k_step only ever writes rows that occur ONCE in the minibatch (by their one
// occurrence), so no pair reads a row another pair of the same launch writes.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

#include "cf_step_impl.cuh"

namespace cfstep {

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_pending() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

template <int MODEL, int W, int NBUF, int MINB>
__global__ void __launch_bounds__(256, MINB) k_step_fast(const __grid_constant__ StepDev P) {
  static_assert(MODEL == CF_MODEL_BPR || MODEL == CF_MODEL_CML, "BPR / CML only");
  constexpr int NS = 2 + W;   // slots of a pair: user, positive item, W negatives; lane s < NS keeps slot s's row id / words
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31;
  const int ld = P.ld;
  const bool act = lane < P.nvec;
  const bool adagrad = P.optimizer == CF_OPT_ADAGRAD;
  const bool want_loss = P.loss != nullptr;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const float creg = (MODEL == CF_MODEL_CML) ? (P.reg > 0.f ? P.reg : 0.f) : P.reg;
  float* wbase = smem + (size_t)(threadIdx.x >> 5) * (NBUF * 2 * NS) * ld;   // [NBUF][param rows NS | accumulator rows NS][ld]
  if (__ldcg(P.counters + 1) & (CF_FLAG_INDEX_RANGE | CF_FLAG_STAGING_FULL)) return;   // (k_count validated every id)

  auto load_id = [&](long long b) -> int {
    if (lane >= NS || b >= P.B) return -1;
    if (lane == 0) return __ldg(P.pairs + 2 * b);
    if (lane == 1) return __ldg(P.pairs + 2 * b + 1);
    return __ldg(P.negs + b * W + (lane - 2));
  };
  auto load_occ = [&](int row) -> unsigned { return row >= 0 ? __ldcg((lane == 0 ? P.metaU : P.metaV) + row) : 0u; };
  auto load_slot = [&](int row, unsigned occ) -> int { return occ > 1u ? __ldcg((lane == 0 ? P.slotU : P.slotV) + row) : 0; };
  auto issue_rows = [&](float* buf, int row, unsigned occ) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const long long r = __shfl_sync(0xffffffffu, row, s);
      if (act) cp_async16(buf + s * ld + 4 * lane, (s == 0 ? P.U : P.V) + r * ld + 4 * lane);
    }
    if (adagrad) {
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const long long r = __shfl_sync(0xffffffffu, row, s);
        const unsigned o = __shfl_sync(0xffffffffu, occ, s);
        if (o <= 1u && act) cp_async16(buf + (NS + s) * ld + 4 * lane, (s == 0 ? P.accU : P.accV) + r * ld + 4 * lane);
      }
    }
  };

  double loss_acc = 0.0;
  long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  long long b1 = b + nwarps;
  int row = load_id(b), row1 = load_id(b1);
  unsigned occ = load_occ(row);
  int slot = 0;
  if (NBUF == 2) {
    if (b < P.B) issue_rows(wbase, row, occ);
    cp_async_commit();
    slot = load_slot(row, occ);
  }
  unsigned occ1 = load_occ(row1);
  int n = 0;
  for (; b < P.B; b = b1, b1 += nwarps) {
    float* cur = wbase + (NBUF == 2 ? (size_t)n * (2 * NS) * ld : 0);
    const int row2 = load_id(b1 + nwarps);
    int slot1 = 0;
    if (NBUF == 2) {
      if (b1 < P.B) issue_rows(wbase + (size_t)(n ^ 1) * (2 * NS) * ld, row1, occ1);
      cp_async_commit();
      slot1 = load_slot(row1, occ1);
      cp_async_wait_pending<1>();
      n ^= 1;
    } else {
      issue_rows(cur, row, occ);
      cp_async_commit();
      slot = load_slot(row, occ);
      cp_async_wait_pending<0>();
    }
    __syncwarp();

    // ---------------------------------------------------------------- forward
    const Row<1> Uu = smem_row<32, 1>(cur, P.nvec, lane);
    const Row<1> Vi = smem_row<32, 1>(cur + ld, P.nvec, lane);
    float sq = 0.f, lossv = 0.f;
    if (want_loss) sq = dotp<1>(Uu, Uu) + dotp<1>(Vi, Vi);
    float S = 0.f, coef = 0.f;
    int wmin = -1;
    float alpha[W];
    Row<1> XA = zero_row<1>();
    if constexpr (MODEL == CF_MODEL_BPR) {
      const float dui = warp_sum(dotp<1>(Uu, Vi));
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const Row<1> Vj = smem_row<32, 1>(cur + (2 + w) * ld, P.nvec, lane);
        const float x = dui - warp_sum(dotp<1>(Uu, Vj));   // bprmf.py:68-70
        const float sw = sigm1(x);
        S += sw;
        axpy<1>(XA, -sw, Vj);
        alpha[w] = -sw;
        if (want_loss) { lossv += softplus_neg(x); sq += dotp<1>(Vj, Vj); }
      }
    } else {
      const float dp = warp_sum(sqdp<1>(Uu, Vi));          // cml.py:63-85
      float dmin = INFINITY;
      int imp = 0;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const Row<1> Vj = smem_row<32, 1>(cur + (2 + w) * ld, P.nvec, lane);
        const float dn = warp_sum(sqdp<1>(Uu, Vj));
        if (dn < dmin) { dmin = dn; wmin = w; }
        imp += ((dp - dn) + P.margin) > 0.f;
        if (want_loss) sq += dotp<1>(Vj, Vj);
      }
      const float h = (dp - dmin) + P.margin;
      const float omega = P.use_rank_weight ? __logf(((float)imp / (float)W) * (float)P.rank_items + 1.f) : 1.f;
      coef = h > 0.f ? 2.f * omega : 0.f;
      if (want_loss) lossv = fmaxf(h, 0.f) * omega;
      if (wmin >= 0) XA = smem_row<32, 1>(cur + (2 + wmin) * ld, P.nvec, lane);   // the closest negative (reduce_min's gradient goes there)
#pragma unroll
      for (int w = 0; w < W; ++w) alpha[w] = (w == wmin) ? coef : 0.f;
    }

    // ---------------------------------------------------------------- commit, slot by slot
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const long long r = __shfl_sync(0xffffffffu, row, s);
      const unsigned o = __shfl_sync(0xffffffffu, occ, s);
      const Row<1> cs = s == 0 ? Uu : s == 1 ? Vi : smem_row<32, 1>(cur + s * ld, P.nvec, lane);
      Row<1> g;
      if constexpr (MODEL == CF_MODEL_BPR) {
        if (s == 0) { g = lin2<1>(S, Vi, creg, cs); axpy<1>(g, 1.f, XA); }        // S V_i - sum s V_j + reg U
        else if (s == 1) g = lin2<1>(S, Uu, creg, cs);                             // S U + reg V_i
        else g = lin2<1>(alpha[s >= 2 ? s - 2 : 0], Uu, creg, cs);                 // -s U + reg V_j
      } else {
        if (s == 0) { g = lin2<1>(coef, XA, creg, cs); axpy<1>(g, -coef, Vi); }   // coef (V_j* - V_i) + c U
        else if (s == 1) g = lin2<1>(-coef, Uu, creg + coef, cs);                  // -coef (U - V_i) + c V_i
        else { const float a = alpha[s >= 2 ? s - 2 : 0]; g = lin2<1>(a, Uu, creg - a, cs); }   // tie (U - V_j) + c V_j
      }
      if (o <= 1u) {   // the row occurs once in the minibatch: update it from registers / shared memory
        Row<1> acc = smem_row<32, 1>(cur + (NS + s) * ld, P.nvec, lane, 1.f), p;
        apply_math<32, 1>(P, cs, acc, g, p, 0xffffffffu);
        if (adagrad) store_row<32, 1>(s == 0 ? P.accU : P.accV, r, ld, P.nvec, lane, acc);
        store_row<32, 1>(s == 0 ? P.U : P.V, r, ld, P.nvec, lane, p);
      } else {         // duplicated row: sum in its staging slot, k_apply_staged applies the sum once
        const int sl = __shfl_sync(0xffffffffu, slot, s);
        if (act) atomicAdd(reinterpret_cast<float4*>(P.staging + (long long)sl * P.lds + 4 * lane), g.v[0]);
      }
    }
    if (lane < NS && occ <= 1u) __stcg((lane == 0 ? P.metaU : P.metaV) + row, 0u);   // unique rows are done
    if (want_loss) {
      const float regsq = warp_sum(sq);
      if (lane == 0) loss_acc += (double)(lossv + 0.5f * creg * regsq);
    }
    __syncwarp();   // this pair's shared-memory reads are done before the buffer is refilled
    row = row1; occ = occ1; row1 = row2;
    if (NBUF == 2) slot = slot1;
    occ1 = load_occ(row1);
  }

  if (want_loss) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    if (lane == 0 && loss_acc != 0.0) atomicAdd(P.loss, loss_acc);
  }
}

}  // namespace cfstep

// the specialised kernel for (model, W), or NULL; *nbuf = shared-memory row buffers per warp, *slots = rows per buffer half
cfstep::step_kernel_t cf_step_pick_fast(int model, int W, int* nbuf, int* slots) {
  using namespace cfstep;
  *slots = 2 + W;
  // W = 5 was measured too (CML configs[1]: 1.999 ms with one buffer at 64 registers, 2.018 ms with two buffers at 2 blocks per
  // SM, generic 1.932 ms): with seven rows per pair the generic kernel is bound by DRAM traffic, not by issue rate, so it stays.
  if (model == CF_MODEL_BPR && W == 1) { *nbuf = 2; return k_step_fast<CF_MODEL_BPR, 1, 2, 4>; }
  if (model == CF_MODEL_CML && W == 1) { *nbuf = 2; return k_step_fast<CF_MODEL_CML, 1, 2, 4>; }
  return nullptr;
}
