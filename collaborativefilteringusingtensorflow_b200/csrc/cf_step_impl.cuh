// Device templates of the fused minibatch step (included by the generated instantiation units and by cf_step.cu).
// See cf_step.cu for the design notes.
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
  int B, W, G;
  int model, optimizer, update, use_rank_weight;
  float lr, reg, margin, clip, rho, weight;
  unsigned long long *metaU, *metaV;
  int32_t *slotU, *slotV;
  float* staging;
  long long staging_rows;
  int lds;
  int32_t* counters;
  double* loss;
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
// y += s * x
template <int NV>
__device__ __forceinline__ void axpy(Row<NV>& y, float s, const Row<NV>& x) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    y.v[k].x = fmaf(s, x.v[k].x, y.v[k].x);
    y.v[k].y = fmaf(s, x.v[k].y, y.v[k].y);
    y.v[k].z = fmaf(s, x.v[k].z, y.v[k].z);
    y.v[k].w = fmaf(s, x.v[k].w, y.v[k].w);
  }
}
// a*x + b*y
template <int NV>
__device__ __forceinline__ Row<NV> lin2(float a, const Row<NV>& x, float b, const Row<NV>& y) {
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
  return x > 0.f ? log1pf(expf(-x)) : (-x + log1pf(expf(x)));
}
__device__ __forceinline__ float sigm1(float x) {  // sigmoid(x) - 1
  return -1.f / (1.f + expf(x));
}

// ---- one Adagrad / SGD apply of a (summed) row gradient; optional CML unit-norm clip (cml.py:119-122) fused in
template <int LPG, int NV>
__device__ __forceinline__ void apply_row(const StepDev& P, float* tab, float* acc, long long r, const Row<NV>& cur,
                                          const Row<NV>& g, int gl, unsigned gmask) {
  Row<NV> p;
  if (P.optimizer == CF_OPT_ADAGRAD) {
    Row<NV> a = load_row<LPG, NV>(acc, r, P.ld, P.nvec, gl, 1.f);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      a.v[k].x = fmaf(g.v[k].x, g.v[k].x, a.v[k].x);
      a.v[k].y = fmaf(g.v[k].y, g.v[k].y, a.v[k].y);
      a.v[k].z = fmaf(g.v[k].z, g.v[k].z, a.v[k].z);
      a.v[k].w = fmaf(g.v[k].w, g.v[k].w, a.v[k].w);
      p.v[k].x = cur.v[k].x - (P.lr * g.v[k].x) / sqrtf(a.v[k].x);
      p.v[k].y = cur.v[k].y - (P.lr * g.v[k].y) / sqrtf(a.v[k].y);
      p.v[k].z = cur.v[k].z - (P.lr * g.v[k].z) / sqrtf(a.v[k].z);
      p.v[k].w = cur.v[k].w - (P.lr * g.v[k].w) / sqrtf(a.v[k].w);
    }
    store_row<LPG, NV>(acc, r, P.ld, P.nvec, gl, a);
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
    const float nrm = sqrtf(group_sum<LPG>(dotp<NV>(p, p), gmask));
    const float den = fmaxf(nrm, P.clip);
    if (nrm > P.clip) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        p.v[k].x = (p.v[k].x * P.clip) / den;
        p.v[k].y = (p.v[k].y * P.clip) / den;
        p.v[k].z = (p.v[k].z * P.clip) / den;
        p.v[k].w = (p.v[k].w * P.clip) / den;
      }
    }
  }
  store_row<LPG, NV>(tab, r, P.ld, P.nvec, gl, p);
}

__device__ __forceinline__ void apply_bias(const StepDev& P, long long r, float cur, float g) {
  if (P.optimizer == CF_OPT_ADAGRAD) {
    const float a = fmaf(g, g, __ldcg(P.accb + r));
    __stcg(P.accb + r, a);
    __stcg(P.b + r, cur - (P.lr * g) / sqrtf(a));
  } else {
    __stcg(P.b + r, fmaf(-P.lr, g, cur));
  }
}

// ---- commit one occurrence of row r of table `tab` (0 = U, 1 = V[+bias]) with gradient g (bias gradient gb)
template <int LPG, int NV>
__device__ __forceinline__ void commit(const StepDev& P, int tab, long long r, unsigned long long meta_word,
                                       const Row<NV>& cur, const Row<NV>& g, float bcur, float gb, int gl,
                                       unsigned gmask, int leader) {
  float* T = tab ? P.V : P.U;
  float* A = tab ? P.accV : P.accU;
  const bool bias = tab && P.b != nullptr;
  if (P.update == CF_UPDATE_HOGWILD) {
    apply_row<LPG, NV>(P, T, A, r, cur, g, gl, gmask);
    if (bias && gl == 0) apply_bias(P, r, bcur, gb);
    return;
  }
  unsigned long long* meta = tab ? P.metaV : P.metaU;
  const unsigned occ = (unsigned)__shfl_sync(gmask, meta_word, leader);
  if (occ <= 1u) {  // the only occurrence in this minibatch: update from registers
    apply_row<LPG, NV>(P, T, A, r, cur, g, gl, gmask);
    if (gl == 0) {
      if (bias) apply_bias(P, r, bcur, gb);
      __stcg(meta + r, 0ull);
    }
    return;
  }
  int slot = 0;
  if (gl == 0) slot = __ldcg((tab ? P.slotV : P.slotU) + r);
  slot = __shfl_sync(gmask, slot, leader);
  float* st = P.staging + (long long)slot * P.lds;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = gl + k * LPG;
    if (v < P.nvec) atomicAdd(reinterpret_cast<float4*>(st + 4 * v), g.v[k]);
  }
  if (bias && gl == 0) atomicAdd(st + P.ld, gb);
  __threadfence();
  __syncwarp(gmask);
  unsigned long long old = 0;
  if (gl == 0) old = atomicAdd(meta + r, 1ull << 32);
  old = __shfl_sync(gmask, old, leader);
  if ((unsigned)(old >> 32) + 1u == occ) {  // last arriver: every gradient of this row is in the slot
    __threadfence();
    Row<NV> gt;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = gl + k * LPG;
      if (v < P.nvec) {
        gt.v[k] = ldcg4(st + 4 * v);
        stcg4(st + 4 * v, f4zero());
      } else {
        gt.v[k] = f4zero();
      }
    }
    apply_row<LPG, NV>(P, T, A, r, cur, gt, gl, gmask);
    if (gl == 0) {
      if (bias) {
        const float gbt = __ldcg(st + P.ld);
        __stcg(st + P.ld, 0.f);
        apply_bias(P, r, bcur, gbt);
      }
      __stcg(meta + r, 0ull);
    }
  }
}

__device__ __forceinline__ bool in_range(long long r, long long n) { return r >= 0 && r < n; }

constexpr int GT = 4;  // group rows kept in registers (GBPR)

template <int MODEL, int LPG, int NV, int WT>
__global__ void __launch_bounds__(256) k_step(const __grid_constant__ StepDev P) {
  const int lane = threadIdx.x & 31;
  const int gl = lane & (LPG - 1);
  const int leader = lane & ~(LPG - 1);
  const unsigned gmask = LPG == 32 ? 0xffffffffu : (((1u << LPG) - 1u) << leader);
  const long long ngroups = (long long)gridDim.x * blockDim.x / LPG;
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPG;
  const bool sync = P.update == CF_UPDATE_SYNC;
  const bool want_loss = P.loss != nullptr;
  if (sync && (__ldcg(P.counters + 1) & (CF_FLAG_INDEX_RANGE | CF_FLAG_STAGING_FULL))) return;

  double loss_acc = 0.0;
  const long long iters = (P.B + ngroups - 1) / ngroups;
  for (long long it = 0; it < iters; ++it) {
    const long long b0 = gid + it * ngroups;
    bool active = b0 < P.B;
    const long long bb = active ? b0 : (long long)P.B - 1;
    const int u = __ldg(P.pairs + 2 * bb), i = __ldg(P.pairs + 2 * bb + 1);
    if (!in_range(u, P.n_users) || !in_range(i, P.n_items)) {
      if (active && gl == 0) atomicOr(P.counters + 1, CF_FLAG_INDEX_RANGE);
      active = false;
    }
    // negatives / group ids are validated below; an invalid id deactivates the whole pair
    if (MODEL != CF_MODEL_WRMF) {
      for (int w = 0; w < P.W; ++w)
        if (!in_range(__ldg(P.negs + bb * P.W + w), P.n_items)) {
          if (active && gl == 0) atomicOr(P.counters + 1, CF_FLAG_INDEX_RANGE);
          active = false;
        }
    }
    if (MODEL == CF_MODEL_GBPR) {
      for (int g = 0; g < P.G; ++g)
        if (!in_range(__ldg(P.group + bb * P.G + g), P.n_users)) {
          if (active && gl == 0) atomicOr(P.counters + 1, CF_FLAG_INDEX_RANGE);
          active = false;
        }
    }
    if (!active) continue;  // group-uniform; every shuffle below uses the group's own mask

    unsigned long long mu = 0, mi = 0;
    if (sync && gl == 0) {
      mu = __ldcg(P.metaU + u);
      mi = __ldcg(P.metaV + i);
    }
    const Row<NV> Uu = load_row<LPG, NV>(P.U, u, P.ld, P.nvec, gl);
    const Row<NV> Vi = load_row<LPG, NV>(P.V, i, P.ld, P.nvec, gl);
    float lossv = 0.f, sq = 0.f;

    if constexpr (MODEL == CF_MODEL_BPR) {
      // bprmf.py:52-75: x_bw = <U_u,V_i> - <U_u,V_jw>; s = sigmoid(x) - 1
      const float dui = group_sum<LPG>(dotp<NV>(Uu, Vi), gmask);
      Row<NV> gU = zero_row<NV>();
      float S = 0.f;
      if (want_loss) sq = dotp<NV>(Uu, Uu) + dotp<NV>(Vi, Vi);
      for (int w0 = 0; w0 < P.W; w0 += WT) {
        Row<NV> Vj[WT];
        int j[WT];
        unsigned long long mj[WT];
#pragma unroll
        for (int t = 0; t < WT; ++t)
          if (w0 + t < P.W) {
            j[t] = __ldg(P.negs + bb * P.W + w0 + t);
            mj[t] = (sync && gl == 0) ? __ldcg(P.metaV + j[t]) : 0ull;
            Vj[t] = load_row<LPG, NV>(P.V, j[t], P.ld, P.nvec, gl);
          }
#pragma unroll
        for (int t = 0; t < WT; ++t)
          if (w0 + t < P.W) {
            const float x = dui - group_sum<LPG>(dotp<NV>(Uu, Vj[t]), gmask);
            const float s = sigm1(x);
            S += s;
            if (want_loss) {
              lossv += softplus_neg(x);
              sq += dotp<NV>(Vj[t], Vj[t]);
            }
            axpy<NV>(gU, s, Vi);
            axpy<NV>(gU, -s, Vj[t]);
            const Row<NV> g = lin2<NV>(-s, Uu, P.reg, Vj[t]);
            commit<LPG, NV>(P, 1, j[t], mj[t], Vj[t], g, 0.f, 0.f, gl, gmask, leader);
          }
      }
      axpy<NV>(gU, P.reg, Uu);
      commit<LPG, NV>(P, 0, u, mu, Uu, gU, 0.f, 0.f, gl, gmask, leader);
      const Row<NV> gV = lin2<NV>(S, Uu, P.reg, Vi);
      commit<LPG, NV>(P, 1, i, mi, Vi, gV, 0.f, 0.f, gl, gmask, leader);
      if (want_loss) lossv += 0.5f * P.reg * group_sum<LPG>(sq, gmask);
    } else if constexpr (MODEL == CF_MODEL_CML) {
      // cml.py:55-109: hinge on squared distances vs the closest of W negatives, WARP-style rank weight
      const float c = P.reg > 0.f ? P.reg : 0.f;
      const float dp = group_sum<LPG>(sqdp<NV>(Uu, Vi), gmask);
      const bool single = P.W <= WT;
      Row<NV> Vj[WT];
      int j[WT];
      unsigned long long mj[WT];
      float dmin = INFINITY;
      int wmin = -1, imp = 0;
      if (want_loss) sq = dotp<NV>(Uu, Uu) + dotp<NV>(Vi, Vi);
      for (int w0 = 0; w0 < P.W; w0 += WT) {
#pragma unroll
        for (int t = 0; t < WT; ++t)
          if (w0 + t < P.W) {
            j[t] = __ldg(P.negs + bb * P.W + w0 + t);
            mj[t] = (sync && gl == 0) ? __ldcg(P.metaV + j[t]) : 0ull;
            Vj[t] = load_row<LPG, NV>(P.V, j[t], P.ld, P.nvec, gl);
          }
#pragma unroll
        for (int t = 0; t < WT; ++t)
          if (w0 + t < P.W) {
            const float dn = group_sum<LPG>(sqdp<NV>(Uu, Vj[t]), gmask);
            if (dn < dmin) {
              dmin = dn;
              wmin = w0 + t;
            }
            imp += ((dp - dn) + P.margin) > 0.f;
            if (want_loss) sq += dotp<NV>(Vj[t], Vj[t]);
          }
      }
      const float h = (dp - dmin) + P.margin;
      const float omega = P.use_rank_weight ? logf(((float)imp / (float)P.W) * (float)P.n_items + 1.f) : 1.f;
      const float coef = h > 0.f ? 2.f * omega : 0.f;
      if (want_loss) lossv = fmaxf(h, 0.f) * omega + 0.5f * c * group_sum<LPG>(sq, gmask);
      Row<NV> dUi = lin2<NV>(1.f, Uu, -1.f, Vi);
      Row<NV> gU = lin2<NV>(coef, dUi, c, Uu);
      for (int w0 = 0; w0 < P.W; w0 += WT) {
        if (!single) {
#pragma unroll
          for (int t = 0; t < WT; ++t)
            if (w0 + t < P.W) {
              j[t] = __ldg(P.negs + bb * P.W + w0 + t);
              mj[t] = (sync && gl == 0) ? __ldcg(P.metaV + j[t]) : 0ull;
              Vj[t] = load_row<LPG, NV>(P.V, j[t], P.ld, P.nvec, gl);
            }
        }
#pragma unroll
        for (int t = 0; t < WT; ++t)
          if (w0 + t < P.W) {
            const float tie = (w0 + t == wmin) ? coef : 0.f;  // reduce_min grad -> the (first) closest negative
            const Row<NV> dUj = lin2<NV>(1.f, Uu, -1.f, Vj[t]);
            axpy<NV>(gU, -tie, dUj);
            const Row<NV> g = lin2<NV>(tie, dUj, c, Vj[t]);
            commit<LPG, NV>(P, 1, j[t], mj[t], Vj[t], g, 0.f, 0.f, gl, gmask, leader);
          }
      }
      commit<LPG, NV>(P, 0, u, mu, Uu, gU, 0.f, 0.f, gl, gmask, leader);
      const Row<NV> gV = lin2<NV>(-coef, dUi, c, Vi);
      commit<LPG, NV>(P, 1, i, mi, Vi, gV, 0.f, 0.f, gl, gmask, leader);
    } else if constexpr (MODEL == CF_MODEL_GBPR) {
      // gbprmf.py:58-93: r_ui = rho * mean_g <U_g,V_i> + (1-rho) <U_u,V_i> + b_i ; r_uj = <U_u,V_j> + b_j
      const float invG = 1.f / (float)P.G;
      float bi = 0.f;
      if (gl == 0) bi = __ldcg(P.b + i);
      bi = __shfl_sync(gmask, bi, leader);
      Row<NV> Ug[GT];
      int gi[GT];
      unsigned long long mg[GT];
      Row<NV> Ugs = zero_row<NV>();
      const bool gsingle = P.G <= GT;
      if (want_loss) sq = dotp<NV>(Uu, Uu) + dotp<NV>(Vi, Vi);
      for (int g0 = 0; g0 < P.G; g0 += GT) {
#pragma unroll
        for (int t = 0; t < GT; ++t)
          if (g0 + t < P.G) {
            gi[t] = __ldg(P.group + bb * P.G + g0 + t);
            mg[t] = (sync && gl == 0) ? __ldcg(P.metaU + gi[t]) : 0ull;
            Ug[t] = load_row<LPG, NV>(P.U, gi[t], P.ld, P.nvec, gl);
          }
#pragma unroll
        for (int t = 0; t < GT; ++t)
          if (g0 + t < P.G) {
            axpy<NV>(Ugs, 1.f, Ug[t]);
            if (want_loss) sq += dotp<NV>(Ug[t], Ug[t]);
          }
      }
      const float ui_u = group_sum<LPG>(dotp<NV>(Uu, Vi), gmask);
      const float ui_g = group_sum<LPG>(dotp<NV>(Ugs, Vi), gmask) * invG;
      const float ui = P.rho * ui_g + (1.f - P.rho) * ui_u + bi;
      Row<NV> gU = zero_row<NV>();
      float S = 0.f, bsq = 0.f;
      for (int w0 = 0; w0 < P.W; w0 += WT) {
        Row<NV> Vj[WT];
        int j[WT];
        unsigned long long mj[WT];
        float bj[WT];
#pragma unroll
        for (int t = 0; t < WT; ++t)
          if (w0 + t < P.W) {
            j[t] = __ldg(P.negs + bb * P.W + w0 + t);
            mj[t] = (sync && gl == 0) ? __ldcg(P.metaV + j[t]) : 0ull;
            bj[t] = (gl == 0) ? __ldcg(P.b + j[t]) : 0.f;
            Vj[t] = load_row<LPG, NV>(P.V, j[t], P.ld, P.nvec, gl);
          }
#pragma unroll
        for (int t = 0; t < WT; ++t)
          if (w0 + t < P.W) {
            const float bjt = __shfl_sync(gmask, bj[t], leader);
            const float x = ui - (group_sum<LPG>(dotp<NV>(Uu, Vj[t]), gmask) + bjt);
            const float s = sigm1(x);
            S += s;
            if (want_loss) {
              lossv += softplus_neg(x);
              bsq += bjt * bjt;
            }
            axpy<NV>(gU, -s, Vj[t]);
            const Row<NV> g = lin2<NV>(-s, Uu, 0.f, Vj[t]);  // no L2 on V_j (gbprmf.py:60-64)
            commit<LPG, NV>(P, 1, j[t], mj[t], Vj[t], g, bjt, fmaf(P.reg, bjt, -s), gl, gmask, leader);
          }
      }
      axpy<NV>(gU, (1.f - P.rho) * S, Vi);
      axpy<NV>(gU, P.reg, Uu);
      commit<LPG, NV>(P, 0, u, mu, Uu, gU, 0.f, 0.f, gl, gmask, leader);
      const float cg = P.rho * invG * S;
      for (int g0 = 0; g0 < P.G; g0 += GT) {
        if (!gsingle) {
#pragma unroll
          for (int t = 0; t < GT; ++t)
            if (g0 + t < P.G) {
              gi[t] = __ldg(P.group + bb * P.G + g0 + t);
              mg[t] = (sync && gl == 0) ? __ldcg(P.metaU + gi[t]) : 0ull;
              Ug[t] = load_row<LPG, NV>(P.U, gi[t], P.ld, P.nvec, gl);
            }
        }
#pragma unroll
        for (int t = 0; t < GT; ++t)
          if (g0 + t < P.G) {
            const Row<NV> g = lin2<NV>(cg, Vi, P.reg, Ug[t]);
            commit<LPG, NV>(P, 0, gi[t], mg[t], Ug[t], g, 0.f, 0.f, gl, gmask, leader);
          }
      }
      Row<NV> gV = lin2<NV>(P.rho * invG * S, Ugs, (1.f - P.rho) * S, Uu);
      axpy<NV>(gV, P.reg, Vi);
      commit<LPG, NV>(P, 1, i, mi, Vi, gV, bi, S, gl, gmask, leader);
      if (want_loss) lossv += 0.5f * P.reg * (group_sum<LPG>(sq, gmask) + bsq);
    } else {
      // wrmf.py:52-75: e = <U_u,V_i> - r ; L = weight/2 e^2 + reg/2 (|U_u|^2 + |V_i|^2)
      const float r = __ldg(P.ratings + bb);
      const float e = group_sum<LPG>(dotp<NV>(Uu, Vi), gmask) - r;
      const float we = P.weight * e;
      if (want_loss)
        lossv = 0.5f * P.weight * e * e + 0.5f * P.reg * group_sum<LPG>(dotp<NV>(Uu, Uu) + dotp<NV>(Vi, Vi), gmask);
      const Row<NV> gU = lin2<NV>(we, Vi, P.reg, Uu);
      const Row<NV> gV = lin2<NV>(we, Uu, P.reg, Vi);
      commit<LPG, NV>(P, 0, u, mu, Uu, gU, 0.f, 0.f, gl, gmask, leader);
      commit<LPG, NV>(P, 1, i, mi, Vi, gV, 0.f, 0.f, gl, gmask, leader);
    }
    if (want_loss && gl == 0) loss_acc += (double)lossv;
  }

  if (want_loss) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    if (lane == 0 && loss_acc != 0.0) atomicAdd(P.loss, loss_acc);
  }
  if (sync) {  // last block out resets the slot counter for the next minibatch
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const int t = atomicAdd(P.counters + 2, 1);
      if (t == (int)gridDim.x - 1) {
        P.counters[0] = 0;
        P.counters[2] = 0;
        __threadfence();
      }
    }
  }
}


typedef void (*step_kernel_t)(const StepDev);

template <int MODEL, int LPG, int NV>
inline step_kernel_t pick_wt(int W) {
  if (MODEL == CF_MODEL_WRMF) return k_step<MODEL, LPG, NV, 1>;
  if (W <= 1) return k_step<MODEL, LPG, NV, 1>;
  if constexpr (NV >= 4) {  // very wide rows (ld > 256): fewer register-resident negatives, longer tile loops
    return k_step<MODEL, LPG, NV, 4>;
  } else {
    if (W <= 2) return k_step<MODEL, LPG, NV, 2>;
    if (W <= 4) return k_step<MODEL, LPG, NV, 4>;
    return k_step<MODEL, LPG, NV, 8>;
  }
}

}  // namespace cfstep

// one instantiation unit per (model, row shape): build.py generates build/gen/cf_step_inst_<m>_<s>.cu defining these
#define CF_STEP_PICK_DECL(M, S) cfstep::step_kernel_t cf_step_pick_##M##_##S(int W)
