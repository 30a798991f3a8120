// Tensor-core full-catalog top-K for sm_100a: tcgen05.mma (fp16 -> fp32 in TMEM) fed by TMA, with an epilogue that keeps a
// per-row candidate superset instead of writing scores, followed by an exact fp64 re-rank of the candidates.
//
// Replaces `sess.run(tf.nn.top_k(matmul(U[test_users], V^T) (+b | cml distance), K'))` + the Python filter loop
// (reference src/models/pl/models/bprmf.py:77-103, cml.py:111-144, gbprmf.py:95-121, basic/models/wrmf.py:77-111).
//
// Stage 0 (k_prep): fp32 tables -> fp16 operand matrices (fp16, not bf16: 8x tighter error band for the same MMA rate)
//   with K padded to a multiple of 64.  The three scoring kinds all become plain dot products a'.b':   DOT       a' = u            b' = v
//                                                               DOT_BIAS  a' = [u, 1, 1]    b' = [v, hi(b_i), lo(b_i)]
//                                                               NEG_SQDIST a' = [2u, 1, 1]  b' = [v, hi(-|v|^2), lo(-|v|^2)]
//   (per-user constants such as -|u|^2 do not change a user's ranking).  eps_row ~ 2^-10 |a'| max_i|b'_i| (exact form at
//   its computation below) bounds the fp16 rounding + fp32 accumulation error of every score of the row.
// Stage 1 (k_topk_tc): one CTA per (256 query rows, item split): A = 2 x (128 x Kp) resident in smem, B tiles of 128 items
//   streamed by TMA through a ring, 2 x tcgen05.mma (M=128, N=128) per k-step into a double-buffered TMEM accumulator,
//   8 epilogue warps (one thread per row) read the accumulators with tcgen05.ld and append (score, item) to the row's
//   candidate buffer iff score >= theta_row - 2 eps_row, where theta_row is the running K-th best approximate score of
//   unmasked items.  Any item of the exact top-K satisfies that test (proof in DESIGN.md), so the buffer is a superset.
//   Training items are masked by a binary search of the user's CSR row -- only for the rare candidates.  A full buffer is
//   compacted warp-cooperatively (radix select of the K-th key, drop everything below theta - 2 eps).
// Stage 2 (k_rerank): exact scores (fp32 inputs, fp64 sequential-k accumulation, identical to cf_topk_exact) of the
//   candidates, bitonic sort by (score desc, id asc), top K.  Rows whose buffer overflowed (degenerate score
//   distributions) are recomputed by the exact streaming kernel.
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TC_M = 128;        // rows per MMA
constexpr int TC_MT = 2;         // M tiles per CTA (256 query rows)
constexpr int TC_N = 128;        // items per B tile of the narrow kernel (n_factors > 126: KC >= 3)
constexpr int TC_NW = 256;       // items per B tile of the wide kernel: ONE tcgen05.mma covers N = 256 items.  Measured on B200
                                 // (tests/micro/mma_micro.cu, 148 CTAs, no epilogue): a 128 x 128 x 16 MMA costs ~122 cycles
                                 // whatever feeds it (A in shared or tensor memory, static B or a TMA ring, M-tiles interleaved or
                                 // not), a 128 x 256 x 16 one ~175: ~70 cycles per instruction + 0.41 per column, so N = 256
                                 // does 1.4x the flops per cycle (1.15 -> 1.56 PFLOP/s static, 1.35 through a 2-stage TMA ring)
constexpr int TC_KCH = 64;       // bf16 elements per 128-byte swizzle chunk
constexpr int TC_CHUNK_BYTES = TC_M * TC_KCH * 2;   // 16 KB: one TMA box {64, 128}
constexpr int TC_CAP = 512;      // candidate buffer entries per (row, split)
constexpr int TC_KMAX = 200;     // largest K of ONE sweep (a compacted row keeps K + the 2-eps band <= CAP - 128)
constexpr int TC_KMAX_ROUNDS = 1024;   // largest K of a call: K > TC_KMAX is served in rounds of TC_KMAX (see cf_topk_tc)
constexpr int TC_THREADS = 384;  // warp 0: TMA, 1: MMA, 2: TMEM alloc, 3: idle, 4..11: epilogue
constexpr int TC_MAX_STAGES = 6;

struct TcParams {
  int T, N, KC, n_tiles, S, stages, K;   // n_tiles: B tiles of NB items
  const int32_t* users;            // [T] user id of every query row (for the training-row mask), or NULL = row index
  const long long* tr_indptr;      // training CSR (NULL = no mask)
  const int32_t* tr_indices;
  const float* eps2;               // [T_pad] 2 * eps_row
  uint2* cand;                     // [T_pad, S, CAP] entries {score bits, item}: one 8-byte store per hit
  int32_t* cand_cnt;               // [T_pad, S]
  int32_t* overflow;               // [T_pad]
  float* dbg_scores;               // optional [T_pad, dbg_ld] dump of the raw accumulators
  long long dbg_ld;
  int mask_by_row;                 // rounds (K > TC_KMAX): the mask CSR is indexed by the query row, not by the user id
  int warm;                        // warm-up tiles per split (k_topk_tc; see RowSweep::sweep_warm)
  int coop, coop_low;              // cooperative compactions (k_topk_tc): on / rows above this many entries join
};

using namespace tc;

__device__ __forceinline__ unsigned ord_key(float f) {  // monotone float -> uint map
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_unkey(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// Warp-cooperative compaction of ONE row's candidate buffer (n entries): entries [ver, n) were appended WITHOUT looking at
// the user's training row (a binary search per hit inside the per-thread sweep is a chain of dependent global loads that
// nothing hides: it was most of the sweep on small catalogues) -- they are checked here; then
// theta <- (a lower bound within 2^-15 relative of) the K-th largest value among the unmasked entries, keep >= theta - eps2.
// A compaction stalls its warp and, two tiles later, the CTA's MMA pipeline, and on a 500 k-item catalogue every row needs
// ~6 of them (the thresholds start at -inf in every CTA): its latency is what the small-catalogue throughput hangs on.
// Hence (1) the mask check is a MERGE, not 16 global binary searches per lane: a split sweeps its items in ascending order,
// so the unchecked entries are ascending and all lie above the last checked one; `tcur` remembers how far into the user's
// sorted training row the earlier compactions got, the next 32 training items arrive with ONE coalesced load (one per lane)
// and every entry looks itself up in that register segment with five shuffles -- one memory round trip per compaction
// instead of seven dependent ones; (2) the counting rounds of the radix select reduce with redux.sync (one instruction)
// instead of a five-step shuffle tree, and stop after the 24 leading bits (any lower bound of the K-th value keeps the
// superset property).
__device__ __forceinline__ void compact_row(uint2* ce, int n, int ver, int K, float eps2, const int32_t* tr_indices,
                                            long long& tcur, long long thi, int lane, int n_items, float& theta_out, int& cnt_out) {
  constexpr int EPL = TC_CAP / 32;
  float v[EPL];
  int32_t id[EPL];
  const bool masking = thi > tcur && n > ver;
  const int last_x = masking ? (int)__ldcg(&ce[n - 1].y) : -1;       // the largest unchecked item id (entries ascend)
  int seg = (masking && tcur + lane < thi) ? __ldg(tr_indices + tcur + lane) : 0x7fffffff;
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const int e = lane + 32 * i;
    const uint2 x = e < n ? __ldcg(ce + e) : make_uint2(0xffffffffu, 0xffffffffu);
    v[i] = __uint_as_float(x.x);
    id[i] = (int32_t)x.y;
    if ((unsigned)id[i] >= (unsigned)n_items) {   // an empty slot, or a padding column of the last tile: key 0 (below every real score)
      v[i] = __uint_as_float(0xffffffffu);
      id[i] = -1;
    }
  }
  if (masking) {
    long long pos = tcur;
    for (;;) {
      const int seg_next = (pos + 32 + lane < thi) ? __ldg(tr_indices + pos + 32 + lane) : 0x7fffffff;   // in flight during the lookups
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        const int x = (lane + 32 * i >= ver) ? id[i] : -1;           // (checked entries and empty slots look up -1: never found)
        int lb = 0;                                                   // lower bound of x among the 32 sorted lanes
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const int probe = __shfl_sync(0xffffffffu, seg, lb + step - 1);
          if (probe < x) lb += step;
        }
        const int found = __shfl_sync(0xffffffffu, seg, lb & 31);
        if (found == x && lb < 32 && x >= 0) {                        // a training item: drop it
          v[i] = __uint_as_float(0xffffffffu);
          id[i] = -1;
        }
      }
      const int c = __popc(__ballot_sync(0xffffffffu, seg <= last_x));   // (sorted: a prefix of the lanes)
      pos += c;
      if (c < 32) break;                                              // the segment reaches beyond the last entry (or the row ended)
      seg = seg_next;
    }
    tcur = pos;
  }
  unsigned key[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) key[i] = ord_key(v[i]);
  unsigned prefix = 0u;
#pragma unroll 1
  for (int bit = 31; bit >= 8; --bit) {
    const unsigned cand = prefix | (1u << bit);
    int c = 0;
#pragma unroll
    for (int i = 0; i < EPL; ++i) c += key[i] >= cand;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= K) prefix = cand;
  }
  // fewer than K unmasked entries so far: no threshold yet, keep them all
  const float theta = prefix ? ord_unkey(prefix) : -INFINITY;
  const float cut = theta - eps2;
  int pos = 0;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const bool keep = id[i] >= 0 && v[i] >= cut;   // (padding and masked entries have id -1)
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) ce[pos + __popc(m & ((1u << lane) - 1u))] = make_uint2(__float_as_uint(v[i]), (unsigned)id[i]);
    pos += __popc(m);
  }
  __syncwarp();
  theta_out = theta;
  cnt_out = pos;
}

// The per-thread state of one query row's candidate buffer (one per (row, item split) -- and, in the paired kernel, per column
// half), shared by both sweep kernels.
struct RowSweep {
  float theta, eps2;
  int cnt, ver;              // entries [0, ver) of the buffer are known not to be training items
  long long tcur, thi;       // cursor into / end of the user's sorted training row
  uint2* ce;
  bool valid, overflowed;

  __device__ __forceinline__ void init(const TcParams& P, int row, long long buffer) {
    valid = row < P.T;
    theta = valid ? -INFINITY : INFINITY;
    cnt = 0;
    ver = 0;
    eps2 = valid ? P.eps2[row] : 0.f;
    tcur = 0;
    thi = 0;
    if (valid && P.tr_indptr) {
      const long long u = (P.users && !P.mask_by_row) ? P.users[row] : row;
      tcur = P.tr_indptr[u];
      thi = P.tr_indptr[u + 1];
    }
    ce = P.cand + buffer * TC_CAP;
    overflowed = false;
  }

  // compaction of the rows of the lanes in `need`, one row at a time by the whole warp
  __device__ __forceinline__ void compact(unsigned need, int lane, const TcParams& P) {
    while (need) {
      const int l = __ffs(need) - 1;
      need &= need - 1;
      uint2* rce = reinterpret_cast<uint2*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(ce), l));
      const int rn = __shfl_sync(0xffffffffu, cnt, l);
      const int rver = __shfl_sync(0xffffffffu, ver, l);
      const float re = __shfl_sync(0xffffffffu, eps2, l);
      long long rcur = __shfl_sync(0xffffffffu, tcur, l);
      const long long rhi = __shfl_sync(0xffffffffu, thi, l);
      float nth;
      int ncnt;
      compact_row(rce, rn, rver, P.K, re, P.tr_indices, rcur, rhi, lane, P.N, nth, ncnt);
      if (lane == l) {
        tcur = rcur;
        theta = nth;
        cnt = ncnt;
        ver = ncnt;
        if (ncnt > TC_CAP - 128) {  // too many items within 2 eps of the K-th best: give this row to the exact kernel
          overflowed = true;
          theta = INFINITY;
          cnt = 0;
          ver = 0;
        }
      }
    }
  }

  // One chunk of 32 scores (items item0 ..).  Group maxima of 8 (FMNMX3 trees), then their maximum: a chunk without a hit
  // costs 18 instructions; a hit makes the warp scan only the groups that hold one (on a 500 k-item catalogue 6 % of a
  // row's chunks hold a hit, so 86 % of a WARP's chunks do: the scan is the common path there, and it appends with one
  // 8-byte store per hit; padding columns of the last tile are dropped by the compaction and the re-rank).
  __device__ __forceinline__ void sweep(const uint32_t (&r)[32], int item0, float thr) {
    float mg[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float x = fmaxf(fmaxf(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1])), __uint_as_float(r[8 * g + 2]));
      x = fmaxf(fmaxf(x, __uint_as_float(r[8 * g + 3])), __uint_as_float(r[8 * g + 4]));
      x = fmaxf(fmaxf(x, __uint_as_float(r[8 * g + 5])), __uint_as_float(r[8 * g + 6]));
      mg[g] = fmaxf(x, __uint_as_float(r[8 * g + 7]));
    }
    const float m = fmaxf(fmaxf(fmaxf(mg[0], mg[1]), mg[2]), mg[3]);
    if (m >= thr) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (mg[g] >= thr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (__uint_as_float(r[8 * g + j]) >= thr) {
              ce[cnt] = make_uint2(r[8 * g + j], (unsigned)(item0 + 8 * g + j));
              ++cnt;
            }
          }
        }
      }
    }
  }

  // Threshold warm-up.  A sweep that starts at theta = -inf appends EVERY score until the first compaction and every fifth one
  // until the second: the first ~16 k items of a split cost ten times the steady state, which is nothing on a 10 M-item
  // catalogue and a third of the sweep on a 500 k one.  So the first `warm` tiles of a split are swept twice.  The first time
  // only ONE entry per chunk of 32 columns is appended -- the chunk's maximum, with its item id: at most TC_CAP real
  // (score, item) entries, ascending by item like any other -- and a compaction at the end of the warm-up turns them into
  // theta_0 = the K-th best of the unmasked ones.  Any K unmasked items bound the K-th best score of the whole catalogue from
  // below, so the superset argument holds for theta_0 as it does for the running threshold.  Then the buffer is
  // emptied, the mask cursor rewound and the sweep proper starts at tile 0 with theta_0 (~ the 110th best of 16 k items for
  // K = 100: 0.7 % of the scores pass from the first tile on).  Price: `warm` extra tiles of MMA (3 % at 500 k items).
  __device__ __forceinline__ void sweep_warm(const uint32_t (&r)[32], int item0) {
    float mg[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float x = fmaxf(fmaxf(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1])), __uint_as_float(r[8 * g + 2]));
      x = fmaxf(fmaxf(x, __uint_as_float(r[8 * g + 3])), __uint_as_float(r[8 * g + 4]));
      x = fmaxf(fmaxf(x, __uint_as_float(r[8 * g + 5])), __uint_as_float(r[8 * g + 6]));
      mg[g] = fmaxf(x, __uint_as_float(r[8 * g + 7]));
    }
    const float m = fmaxf(fmaxf(fmaxf(mg[0], mg[1]), mg[2]), mg[3]);
    if (valid) {
      int idx = 0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (mg[g] == m) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (__uint_as_float(r[8 * g + j]) == m) idx = 8 * g + j;
        }
      }
      ce[cnt] = make_uint2(__float_as_uint(m), (unsigned)(item0 + idx));
      ++cnt;
    }
  }
  __device__ __forceinline__ void end_warmup(const TcParams& P, int row, int lane) {
    compact(__ballot_sync(0xffffffffu, valid && cnt > 0), lane, P);
    cnt = 0;
    ver = 0;
    if (valid && P.tr_indptr) tcur = P.tr_indptr[(P.users && !P.mask_by_row) ? P.users[row] : row];
  }

  // a last compaction leaves K + the 2-eps band per buffer instead of whatever arrived since the previous one
  // (~280 -> ~110 at K = 100): the exact re-rank scores and sorts that many fewer candidates
  __device__ __forceinline__ void finish(const TcParams& P, int row, long long buffer, int lane) {
    compact(__ballot_sync(0xffffffffu, valid && !overflowed && cnt > P.K + 16), lane, P);
    if (valid) {
      P.cand_cnt[buffer] = cnt | (ver << 16);   // (entries below ver are known not to be training items)
      if (overflowed) P.overflow[row] = 1;
    }
  }
};

template <int NB>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_topk_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmV, const __grid_constant__ TcParams P) {
  constexpr bool WIDE = NB == TC_NW;   // WIDE: one 256-column accumulator per M tile (tfull / tempty indexed by the M tile);
                                       // narrow: two stages of 2 x 128 columns (indexed by the tile's parity)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KC = P.KC;
  const int a_bytes = TC_MT * KC * TC_CHUNK_BYTES, b_chunk_bytes = (NB / 128) * TC_CHUNK_BYTES, b_stage_bytes = KC * b_chunk_bytes;
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)P.stages * b_stage_bytes);
  uint64_t* full = bars;                       // [stages]
  uint64_t* empty = bars + TC_MAX_STAGES;      // [stages]
  uint64_t* a_full = bars + 2 * TC_MAX_STAGES;
  uint64_t* tfull = a_full + 1;                // [2]
  uint64_t* tempty = tfull + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  volatile int* s_coop = reinterpret_cast<volatile int*>(tmem_slot + 1);   // tile (+1) of the latest forced compaction in this CTA

  const int row0 = blockIdx.x * (TC_MT * TC_M);
  const int split = blockIdx.y;
  const int tiles_per_split = (P.n_tiles + P.S - 1) / P.S;
  const int tile_lo = split * tiles_per_split;
  const int tile_hi = min(P.n_tiles, tile_lo + tiles_per_split);
  const int nt = max(0, tile_hi - tile_lo);
  const int wt = min(P.warm, nt);   // warm-up tiles: iterations [0, wt) sweep tiles 0 .. wt-1 for thresholds only, [wt, wt + nt) are the sweep
  const int ntt = nt + wt;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(a_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull + s, 1);
      mbar_init(tempty + s, (WIDE ? 4 : 8) * 32);
    }
    *s_coop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Register budget by role (warpgroups of 4 warps): the TMA / MMA / allocator warps need a handful, the epilogue warps hold
  // two chunks of 32 scores plus the compaction's 32 buffer entries: 128 x 80 + 256 x 208 <= 64 K registers (ptxas allocates per region: no spills at 80 / 208)
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 80;" ::: "memory");
  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(a_full, (uint32_t)a_bytes);
      for (int mt = 0; mt < TC_MT; ++mt)
        for (int kc = 0; kc < KC; ++kc)
          tma_load_2d(&tmQ, a_full, sA + (size_t)(mt * KC + kc) * TC_CHUNK_BYTES, kc * TC_KCH, row0 + mt * TC_M);
      int st = 0;
      uint32_t ph = 0u;
      for (int t = 0; t < ntt; ++t) {
        mbar_wait(empty + st, ph ^ 1u);
        mbar_arrive_expect_tx(full + st, (uint32_t)b_stage_bytes);
        for (int kc = 0; kc < KC; ++kc)
          tma_load_2d(&tmV, full + st, sB + (size_t)st * b_stage_bytes + (size_t)kc * b_chunk_bytes, kc * TC_KCH,
                      (tile_lo + (t < wt ? t : t - wt)) * NB);
        if (++st == P.stages) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_fp16(TC_M, NB);
      // Descriptors are precomputed: the single issuing thread must spend only a few instructions per MMA (building two
      // 64-bit descriptors from scratch took ~30 dependent instructions = ~3x the 64-cycle MMA itself and starved the
      // tensor pipe).  Within the 256 KB shared window the 14-bit address field never carries, so an offset is one add.
      uint64_t a0[TC_MT];
#pragma unroll
      for (int mt = 0; mt < TC_MT; ++mt) a0[mt] = umma_desc_sw128(smem_u32(sA + (size_t)mt * KC * TC_CHUNK_BYTES));
      const uint64_t b00 = umma_desc_sw128(smem_u32(sB));
      const uint64_t b_stage_step = (uint64_t)(b_stage_bytes >> 4);
      mbar_wait(a_full, 0u);
      int st = 0;
      uint32_t ph = 0u;
      uint64_t b0 = b00;
      for (int t = 0; t < ntt; ++t) {
        const int acc = t & 1;
        if (!WIDE) mbar_wait(tempty + acc, ((uint32_t)(t >> 1) & 1u) ^ 1u);
        mbar_wait(full + st, ph);
        tc_fence_after();
#pragma unroll
        for (int mt = 0; mt < TC_MT; ++mt) {
          if (WIDE) mbar_wait(tempty + mt, ((uint32_t)t & 1u) ^ 1u);   // the epilogue has drained this M tile's previous scores
          const uint32_t d_tmem = tmem_base + (uint32_t)(WIDE ? mt * NB : acc * (TC_MT * TC_N) + mt * TC_N);
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) {
            if (kc < KC) {
#pragma unroll
              for (int k = 0; k < TC_KCH / 16; ++k) {   // UMMA_K = 16 halfs = 32 bytes inside the 128-byte swizzle atom
                const uint64_t offa = (uint64_t)((kc * TC_CHUNK_BYTES + k * 32) >> 4);
                const uint64_t offb = (uint64_t)((kc * b_chunk_bytes + k * 32) >> 4);
                tc_mma_bf16(d_tmem, a0[mt] + offa, b0 + offb, idesc, (kc | k) ? 1u : 0u);
              }
            }
          }
          if (WIDE) tc_commit(tfull + mt);   // this M tile's 256 scores per row are ready while the other M tile computes
        }
        tc_commit(empty + st);     // smem stage free once these MMAs have read it
        if (!WIDE) tc_commit(tfull + acc);    // accumulator ready for the epilogue
        b0 += b_stage_step;
        if (++st == P.stages) { st = 0; ph ^= 1u; b0 = b00; }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;" ::: "memory");
    // ================================================================== epilogue: one thread per query row
    const int ew = warp - 4;
    const int mt = ew >> 2, q = warp & 3;
    const int row = row0 + mt * TC_M + q * 32 + lane;
    const long long buffer = (long long)row * P.S + split;
    RowSweep rs;
    rs.init(P, row, buffer);
    // The sweep costs instructions, not bandwidth: ncu showed ~110 warp instructions per 32 scores in the first version
    // (room check, four group maxima, four reconvergence points per chunk) and the epilogue warps busy 80 % of the time,
    // with instruction-fetch stalls on top.  Now: ONE room check per 128 columns (a block appends at most 128 entries per
    // row), per chunk 18 instructions unless it holds a hit (RowSweep::sweep).
    const bool dbg = P.dbg_scores != nullptr;
    // Cooperative compactions.  A warp that compacts does not drain its quadrant of the accumulator, and one tile later the MMA
    // pipeline -- the whole CTA -- waits for it: compactions scattered over the eight epilogue warps stall the CTA for the SUM of
    // their durations (on a 500 k-item catalogue at K = 100: ~3 per row, ~2 us each, 8 warps x 32 rows = a fifth of the sweep).
    // So a warp that MUST compact (a row above CAP - 128) announces it in shared memory, and every warp that sees the
    // announcement compacts, within a tile, all of its rows above a lower mark: the warps stall together instead of in turn.
    int coop_seen = 0;
    const int coop_low = P.coop_low;
    for (int t = 0; t < ntt; ++t) {
      const int acc = WIDE ? mt : (t & 1);
      const bool warm = t < wt;
      if (t == wt && wt > 0) rs.end_warmup(P, row, lane);   // thresholds from the warm-up entries; the sweep proper starts at tile 0
      mbar_wait(tfull + acc, WIDE ? ((uint32_t)t & 1u) : ((uint32_t)(t >> 1) & 1u));
      tc_fence_after();
      const int n0 = (tile_lo + (warm ? t : t - wt)) * NB;
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(WIDE ? mt * NB : acc * (TC_MT * TC_N) + mt * TC_N);
#pragma unroll 1
      for (int hb = 0; hb < NB / 128; ++hb) {
        if (!warm) {   // make room for the next 128 columns
          unsigned need = __ballot_sync(0xffffffffu, rs.cnt > TC_CAP - 128);
          int f = *s_coop;
          if (need) {
            if (lane == 0) *s_coop = t + 1;
            f = t + 1;
          }
          f = __shfl_sync(0xffffffffu, f, 0);
          if (f > coop_seen && P.coop) {
            coop_seen = f;
            need = __ballot_sync(0xffffffffu, rs.valid && !rs.overflowed && rs.cnt > coop_low);
          }
          rs.compact(need, lane, P);
        }
        const float thr = rs.theta - rs.eps2;   // (theta only moves at a compaction)
        // software-pipelined TMEM reads: the load of chunk c + 1 is in flight while chunk c is swept
        uint32_t ra[32], rb[32];
        tc_ld32(tbase + (uint32_t)(hb * 128), ra);
#pragma unroll
        for (int cc = 0; cc < 4; cc += 2) {
          const int c = hb * 4 + cc;
          tc_wait_ld(ra);
          tc_ld32(tbase + (uint32_t)((c + 1) * 32), rb);
          if (dbg && rs.valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) P.dbg_scores[(long long)row * P.dbg_ld + n0 + c * 32 + j] = __uint_as_float(ra[j]);
          }
          if (warm) rs.sweep_warm(ra, n0 + c * 32); else rs.sweep(ra, n0 + c * 32, thr);
          tc_wait_ld(rb);
          if (cc + 2 < 4) {
            tc_ld32(tbase + (uint32_t)((c + 2) * 32), ra);
          } else if (c + 1 == NB / 32 - 1) {   // every column of this accumulator is in registers: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(tempty + acc);
          }
          if (dbg && rs.valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) P.dbg_scores[(long long)row * P.dbg_ld + n0 + (c + 1) * 32 + j] = __uint_as_float(rb[j]);
          }
          if (warm) rs.sweep_warm(rb, n0 + (c + 1) * 32); else rs.sweep(rb, n0 + (c + 1) * 32, thr);
        }
      }
    }
    rs.finish(P, row, buffer, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- the paired kernel (CF_TC_PAIR): clusters of two CTAs on the two SMs of a TPC issue ONE tcgen05.mma.cta_group::2 of
// M = 256 (128 query rows per CTA), N = 256 per k-step.  Each CTA TMA-loads HALF of every 256-item B tile (the 2-SM form of
// the load counts its bytes on the leader's mbarrier) and the hardware feeds both halves to both tensor cores: the N = 256
// MMA rate (~1.5 PFLOP/s ceiling instead of ~1.13 for N = 128, tests/micro/mma_micro.cu) at the L2 -> SM traffic of the
// single-CTA kernel, and a two-deep ring of 256-column accumulators.  Only the leader's thread issues MMAs; its commits are
// multicast to both CTAs' barriers; both CTAs' epilogue warps hand an accumulator back on the LEADER's barrier.  Per CTA the
// epilogue is the same sweep, but 128 rows x 256 columns per tile: two warps per TMEM lane quadrant, one per column half,
// each with its own candidate buffer and threshold (so a row has 2 S buffers; the re-rank merges them as it merges splits).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
k_topk_tc_pair(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmV, const __grid_constant__ TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0u;
  const int KC = P.KC;
  const int a_bytes = KC * TC_CHUNK_BYTES, b_stage_bytes = KC * TC_CHUNK_BYTES;   // 128 query rows; 128 of a tile's 256 items
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)P.stages * b_stage_bytes);
  uint64_t* full = bars;                       // [stages]  (the leader's are used)
  uint64_t* empty = bars + TC_MAX_STAGES;      // [stages]  (every CTA's own)
  uint64_t* a_full = bars + 2 * TC_MAX_STAGES; //           (leader)
  uint64_t* tfull = a_full + 1;                // [2]       (every CTA's own)
  uint64_t* tempty = tfull + 2;                // [2]       (leader: 8 epilogue warps of each CTA arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int row0 = (int)(blockIdx.x >> 1) * 256 + (int)rank * TC_M;
  const int split = blockIdx.y;
  const int tiles_per_split = (P.n_tiles + P.S - 1) / P.S;
  const int tile_lo = split * tiles_per_split;
  const int tile_hi = min(P.n_tiles, tile_lo + tiles_per_split);
  const int nt = max(0, tile_hi - tile_lo);

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(a_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull + s, 1);
      mbar_init(tempty + s, 16);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {   // (the same warp in both CTAs)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();      // both CTAs' barriers exist before any remote arrive, multicast commit or 2-SM load targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;" ::: "memory");
    if (warp == 0) {
      // ================================================================== TMA producer (both CTAs: each its half)
      if (lane == 0) {
        if (leader) mbar_arrive_expect_tx(a_full, 2u * (uint32_t)a_bytes);
        for (int kc = 0; kc < KC; ++kc) tma_load_2d_pair(&tmQ, a_full, sA + (size_t)kc * TC_CHUNK_BYTES, kc * TC_KCH, row0);
        int st = 0;
        uint32_t ph = 0u;
        for (int t = 0; t < nt; ++t) {
          mbar_wait(empty + st, ph ^ 1u);
          if (leader) mbar_arrive_expect_tx(full + st, 2u * (uint32_t)b_stage_bytes);
          for (int kc = 0; kc < KC; ++kc)
            tma_load_2d_pair(&tmV, full + st, sB + (size_t)st * b_stage_bytes + (size_t)kc * TC_CHUNK_BYTES, kc * TC_KCH,
                             (tile_lo + t) * TC_NW + (int)rank * TC_M);
          if (++st == P.stages) { st = 0; ph ^= 1u; }
        }
      }
    } else if (warp == 1 && leader) {
      // ================================================================== MMA issuer (one thread of the leader CTA)
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_fp16(2 * TC_M, TC_NW);
        const uint64_t a0 = umma_desc_sw128(smem_u32(sA));
        const uint64_t b00 = umma_desc_sw128(smem_u32(sB));
        const uint64_t b_stage_step = (uint64_t)(b_stage_bytes >> 4);
        mbar_wait(a_full, 0u);
        int st = 0;
        uint32_t ph = 0u;
        uint64_t b0 = b00;
        for (int t = 0; t < nt; ++t) {
          const int acc = t & 1;
          mbar_wait(tempty + acc, ((uint32_t)(t >> 1) & 1u) ^ 1u);   // both CTAs have drained this accumulator
          mbar_wait(full + st, ph);                                  // both halves of the B tile have landed
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_NW);
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) {
            if (kc < KC) {
#pragma unroll
              for (int k = 0; k < TC_KCH / 16; ++k) {
                const uint64_t off = (uint64_t)((kc * TC_CHUNK_BYTES + k * 32) >> 4);
                tc_mma_f16_pair(d_tmem, a0 + off, b0 + off, idesc, (kc | k) ? 1u : 0u);
              }
            }
          }
          tc_commit_pair(empty + st);      // the smem stage is free in both CTAs once these MMAs have read it
          tc_commit_pair(tfull + acc);     // the accumulator is ready for both CTAs' epilogues
          b0 += b_stage_step;
          if (++st == P.stages) { st = 0; ph ^= 1u; b0 = b00; }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;" ::: "memory");
    // ================================================================== epilogue: one thread per (query row, column half)
    const int q = warp & 3, h = (warp - 4) >> 2;
    const int row = row0 + q * 32 + lane;
    const long long buffer = ((long long)row * P.S + split) * 2 + h;
    RowSweep rs;
    rs.init(P, row, buffer);
    const bool dbg = P.dbg_scores != nullptr;
    for (int t = 0; t < nt; ++t) {
      const int acc = t & 1;
      mbar_wait(tfull + acc, (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
      const int n0 = (tile_lo + t) * TC_NW + h * 128;
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TC_NW + h * 128);
      rs.compact(__ballot_sync(0xffffffffu, rs.cnt > TC_CAP - 128), lane, P);   // make room for the next 128 columns
      const float thr = rs.theta - rs.eps2;
      uint32_t ra[32], rb[32];
      tc_ld32(tbase, ra);
#pragma unroll
      for (int c = 0; c < 4; c += 2) {
        tc_wait_ld(ra);
        tc_ld32(tbase + (uint32_t)((c + 1) * 32), rb);
        if (dbg && rs.valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) P.dbg_scores[(long long)row * P.dbg_ld + n0 + c * 32 + j] = __uint_as_float(ra[j]);
        }
        rs.sweep(ra, n0 + c * 32, thr);
        tc_wait_ld(rb);
        if (c + 2 < 4) {
          tc_ld32(tbase + (uint32_t)((c + 2) * 32), ra);
        } else {        // this warp's 128 columns are in registers: one arrival per warp on the leader's barrier
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty + acc, 0u);
        }
        if (dbg && rs.valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) P.dbg_scores[(long long)row * P.dbg_ld + n0 + (c + 1) * 32 + j] = __uint_as_float(rb[j]);
        }
        rs.sweep(rb, n0 + (c + 1) * 32, thr);
      }
    }
    rs.finish(P, row, buffer, lane);
  }

  tc_fence_before();
  cluster_sync_all();      // no CTA of the pair leaves (or frees tensor memory) while the other may still signal it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ stage 0: operands
struct PrepParams {
  const float* src;      // fp32 table [rows, ld]
  const float* bias;     // item bias (DOT_BIAS) or NULL
  const int32_t* ids;    // optional row gather (query users)
  long long n_valid, n_pad;
  int d, ld, Kp, kind, is_query;
  __half* dst;    // [n_pad, Kp]
  float* bmax;           // items: atomicMax of |b'|; queries: read
  float* eps2;           // queries: [n_pad]
  int32_t* stats;        // optional [4]: {overflow rows, candidates, max eps2 bits, bmax bits}
};

__global__ void __launch_bounds__(256) k_prep(const __grid_constant__ PrepParams P) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * blockDim.x / 32;
  float w_nrm = 0.f, w_cabs = 0.f, w_big = 0.f;   // per-warp maxima: one atomic per warp, not one per row (a same-address
                                                  // atomic per row serialised this kernel)
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32; r < P.n_pad; r += nw) {
    __half* out = P.dst + r * P.Kp;
    float nrm2 = 0.f;
    bool big = false;
    const bool valid = r < P.n_valid;
    const long long src_row = valid ? (P.ids ? (long long)P.ids[r] : r) : 0;
    const float scale = (P.is_query && P.kind == CF_SCORE_NEG_SQDIST) ? 2.f : 1.f;
    float vsq = 0.f;
    for (int k = lane; k < P.Kp; k += 32) {
      float x = 0.f;
      if (valid && k < P.d) {
        x = P.src[src_row * P.ld + k];
        vsq += x * x;
        x *= scale;
        big = big || !(fabsf(x) < 65000.f);   // outside fp16 range (or NaN): this row / table cannot use the fp16 pass
      }
      if (k < P.d) {
        out[k] = __float2half_rn(x);
        nrm2 += x * x;
      } else if (k >= P.d + 2 || P.kind == CF_SCORE_DOT) {
        out[k] = __float2half_rn(0.f);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      nrm2 += __shfl_xor_sync(0xffffffffu, nrm2, o);
      vsq += __shfl_xor_sync(0xffffffffu, vsq, o);
    }
    big = __any_sync(0xffffffffu, big);
    float cabs = 0.f;
    if (P.kind != CF_SCORE_DOT && lane == 0) {   // the two augmentation columns
      float c_hi = 0.f, c_lo = 0.f;
      if (valid) {
        if (P.is_query) {
          c_hi = 1.f;
          c_lo = 1.f;
        } else {
          const float c = P.kind == CF_SCORE_DOT_BIAS ? P.bias[src_row] : -vsq;
          c_hi = __half2float(__float2half_rn(c));
          c_lo = c - c_hi;
          cabs = fabsf(c);
        }
      }
      out[P.d] = __float2half_rn(c_hi);
      out[P.d + 1] = __float2half_rn(c_lo);
    }
    if (lane == 0) {
      const float nrm = sqrtf(nrm2);   // norm of the d "main" columns (query: after the x2 of the CML form)
      if (!P.is_query) {
        if (valid) {
          w_nrm = fmaxf(w_nrm, nrm);
          w_cabs = fmaxf(w_cabs, cabs);
          if (big || !(cabs < 65000.f)) w_big = 1.f;
        }
      } else {
        // Operands are rounded to fp16 (11-bit significand): |dx| <= 2^-11 |x| in the normal range, <= 2^-25 below it.
        // main columns: |err| <= (2^-10 + 2^-22) |a||b| + 2^-25 sqrt(Kp) (|a| + |b|); augmentation: c = hi + lo with lo
        // rounded to fp16: |err| <= 2^-22 |c| (charged 2^-20); x1.02 + 1e-6 covers the fp32 accumulation in the tensor core.
        const float e = 0.00097680f * nrm * P.bmax[0] * 1.02f + 2.98e-8f * sqrtf((float)P.Kp) * (nrm + P.bmax[0]) +
                        9.6e-7f * P.bmax[1] + 1e-6f;
        const float e2 = valid ? ((big || P.bmax[2] > 0.f) ? INFINITY : 2.f * e) : 0.f;   // inf -> row goes to the exact kernel
        P.eps2[r] = e2;
        if (P.stats && valid) {
          atomicMax(P.stats + 2, __float_as_int(e2));
          P.stats[3] = __float_as_int(P.bmax[0]);
        }
      }
    }
  }
  if (!P.is_query && lane == 0) {   // non-negative floats order like ints
    atomicMax(reinterpret_cast<int*>(P.bmax), __float_as_int(w_nrm));
    atomicMax(reinterpret_cast<int*>(P.bmax) + 1, __float_as_int(w_cabs));
    if (w_big > 0.f) atomicMax(reinterpret_cast<int*>(P.bmax) + 2, __float_as_int(1.f));
  }
}

// ------------------------------------------------------------------------------------------------ stage 2: exact re-rank
constexpr int RR_CAP = 2048;

struct RerankParams {
  const float *U, *V, *b;
  int ld, nvec, kind, T, S, K;
  const int32_t* users;
  const uint2* cand;
  int n_items;
  const int32_t* cand_cnt;
  const int32_t* overflow;
  const long long* tr_indptr;   // training CSR (NULL = no mask): candidates appended after a row's last compaction are unchecked
  const int32_t* tr_indices;
  int32_t* out_idx;
  double* out_val;
  int32_t* stats;
  int out_ld, out_off;          // the row's K results go to out[t * out_ld + out_off ..]  (rounds: a slice of the caller's [T, K])
  int mask_by_row;              // rounds: the mask CSR is indexed by the query row
  int cap;                      // shared-memory entries (power of two >= S * TC_CAP)
  int count_overflow;           // add the rows left to the exact kernel to stats[0] (rounds count them once, from the sticky flags)
};

__device__ __forceinline__ bool rr_before(double va, int ia, double vb, int ib) { return va > vb || (va == vb && ia < ib); }

// Dynamic shared memory: [cap] fp64 scores | [512] fp64 query row | [cap] item ids, cap = the power of two >= S_cand * TC_CAP
// (<= RR_CAP).  With one split per row (every call with >= 148 * 256 query rows) that is 10 KB instead of 28: sixteen blocks of
// 128 threads per SM instead of eight of 256 -- a row's re-rank is a chain of short latency-bound phases (gather, score, ~30
// barriers of sort), so rows in flight are what its throughput hangs on.
__global__ void __launch_bounds__(256) k_rerank(const __grid_constant__ RerankParams P) {
  extern __shared__ __align__(16) uint8_t rr_smem[];
  double* s_u = reinterpret_cast<double*>(rr_smem);   // the query row, widened once (fp32 -> fp64 conversions run at a quarter of the FMA rate)
  double* s_val = s_u + 512;
  int* s_idx = reinterpret_cast<int*>(s_val + P.cap);
  for (int t = blockIdx.x; t < P.T; t += gridDim.x) {
    if (P.overflow[t]) {            // recomputed by the exact streaming kernel
      if (P.stats && P.count_overflow && threadIdx.x == 0) atomicAdd(P.stats, 1);
      continue;
    }
    const long long u = P.users ? P.users[t] : t;
    const long long mrow = P.mask_by_row ? t : u;
    for (int k = threadIdx.x; k < P.ld; k += blockDim.x) s_u[k] = (double)P.U[u * P.ld + k];
    const long long tlo = P.tr_indptr ? P.tr_indptr[mrow] : 0, thi = P.tr_indptr ? P.tr_indptr[mrow + 1] : 0;
    int total = 0;
    for (int s = 0; s < P.S; ++s) total += P.cand_cnt[(long long)t * P.S + s] & 0xffff;
    int n2 = 32;
    while (n2 < total) n2 <<= 1;
    if (P.stats && threadIdx.x == 0) atomicAdd(P.stats + 1, total);
    __syncthreads();
    // gather + exact score.  The kernel is bound by shared-memory instructions (ncu: mio_throttle + short_scoreboard = 29 of 45
    // stall cycles per issue): every candidate re-reads the whole query row.  So a thread scores TWO candidates (e and e + half)
    // per pass over the row, with 16-byte shared loads; each candidate's own sum keeps its sequential-k order (bit-identical
    // to k_topk_exact).  Entries below `ver` were checked against the training row by the sweep's last compaction.
    int base = 0;
    for (int s = 0; s < P.S; ++s) {
      const int packed = P.cand_cnt[(long long)t * P.S + s];
      const int c = packed & 0xffff, ver = packed >> 16;
      const uint2* ce = P.cand + ((long long)t * P.S + s) * TC_CAP;
      const int half = (c + 1) >> 1;
      for (int e = threadIdx.x; e < half; e += blockDim.x) {
        const int e1 = e + half;
        long long item[2] = {(long long)ce[e].y, e1 < c ? (long long)ce[e1].y : (long long)P.n_items};
        bool live[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int eh = h ? e1 : e;
          live[h] = item[h] < P.n_items && !(eh >= ver && csr_contains(P.tr_indices, tlo, thi, (int)item[h]));
          if (!live[h]) item[h] = 0;   // (a padding column / a training item: scored on row 0 and discarded)
        }
        const float4* vp0 = reinterpret_cast<const float4*>(P.V + item[0] * P.ld);
        const float4* vp1 = reinterpret_cast<const float4*>(P.V + item[1] * P.ld);
        double sc0 = 0.0, sc1 = 0.0;
        if (P.kind == CF_SCORE_NEG_SQDIST) {
#pragma unroll 4
          for (int k4 = 0; k4 < P.nvec; ++k4) {
            const float4 v0 = __ldg(vp0 + k4), v1 = __ldg(vp1 + k4);
            const double2 qa = *reinterpret_cast<const double2*>(s_u + 4 * k4), qb = *reinterpret_cast<const double2*>(s_u + 4 * k4 + 2);
            double df = qa.x - (double)v0.x; sc0 = __dadd_rn(sc0, __dmul_rn(df, df));
            df = qa.y - (double)v0.y; sc0 = __dadd_rn(sc0, __dmul_rn(df, df));
            df = qb.x - (double)v0.z; sc0 = __dadd_rn(sc0, __dmul_rn(df, df));
            df = qb.y - (double)v0.w; sc0 = __dadd_rn(sc0, __dmul_rn(df, df));
            df = qa.x - (double)v1.x; sc1 = __dadd_rn(sc1, __dmul_rn(df, df));
            df = qa.y - (double)v1.y; sc1 = __dadd_rn(sc1, __dmul_rn(df, df));
            df = qb.x - (double)v1.z; sc1 = __dadd_rn(sc1, __dmul_rn(df, df));
            df = qb.y - (double)v1.w; sc1 = __dadd_rn(sc1, __dmul_rn(df, df));
          }
          sc0 = -sc0;
          sc1 = -sc1;
        } else {
#pragma unroll 4
          for (int k4 = 0; k4 < P.nvec; ++k4) {
            const float4 v0 = __ldg(vp0 + k4), v1 = __ldg(vp1 + k4);
            const double2 qa = *reinterpret_cast<const double2*>(s_u + 4 * k4), qb = *reinterpret_cast<const double2*>(s_u + 4 * k4 + 2);
            sc0 = fma(qa.x, (double)v0.x, sc0);
            sc0 = fma(qa.y, (double)v0.y, sc0);
            sc0 = fma(qb.x, (double)v0.z, sc0);
            sc0 = fma(qb.y, (double)v0.w, sc0);
            sc1 = fma(qa.x, (double)v1.x, sc1);
            sc1 = fma(qa.y, (double)v1.y, sc1);
            sc1 = fma(qb.x, (double)v1.z, sc1);
            sc1 = fma(qb.y, (double)v1.w, sc1);
          }
          if (P.kind == CF_SCORE_DOT_BIAS) {
            sc0 = __dadd_rn(sc0, (double)__ldg(P.b + item[0]));
            sc1 = __dadd_rn(sc1, (double)__ldg(P.b + item[1]));
          }
        }
        s_val[base + e] = live[0] ? sc0 : -INFINITY;
        s_idx[base + e] = live[0] ? (int)item[0] : 0x7fffffff;
        if (e1 < c) {
          s_val[base + e1] = live[1] ? sc1 : -INFINITY;
          s_idx[base + e1] = live[1] ? (int)item[1] : 0x7fffffff;
        }
      }
      base += c;
    }
    for (int e = total + threadIdx.x; e < n2; e += blockDim.x) {
      s_val[e] = -INFINITY;
      s_idx[e] = 0x7fffffff;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int x = threadIdx.x; x < n2; x += blockDim.x) {
          const int p = x ^ j;
          if (p > x) {
            const bool up = (x & k) == 0;
            const double va = s_val[x], vb = s_val[p];
            const int ia = s_idx[x], ib = s_idx[p];
            const bool sw = up ? rr_before(vb, ib, va, ia) : rr_before(va, ia, vb, ib);
            if (sw) { s_val[x] = vb; s_idx[x] = ib; s_val[p] = va; s_idx[p] = ia; }
          }
        }
        __syncthreads();
      }
    }
    for (int k = threadIdx.x; k < P.K; k += blockDim.x) {
      const bool ok = k < total && s_idx[k] != 0x7fffffff;
      P.out_idx[(long long)t * P.out_ld + P.out_off + k] = ok ? s_idx[k] : -1;
      if (P.out_val) P.out_val[(long long)t * P.out_ld + P.out_off + k] = ok ? s_val[k] : -INFINITY;
    }
    __syncthreads();
  }
}


// ------------------------------------------------------------------------------------------------ K > TC_KMAX: rounds
// A call with K in (TC_KMAX, TC_KMAX_ROUNDS] (the tail of cml.py:203-211 re-recommends once at topN = 1000) is served in
// ceil(K / TC_KMAX) rounds of the same sweep + re-rank: round r returns the exact top-TC_KMAX among the items that are
// neither training items nor results of the rounds before it, so the concatenation of the rounds IS the exact top-K in
// (score desc, id asc) order.  "Results of the rounds before" are masked like training items: before round r >= 1,
// k_mask_merge writes a mask CSR indexed by QUERY ROW whose row t is the sorted union of the user's training row and the
// row's 200 r earlier results (missing results -- fewer unmasked items than K -- become 0x7fffffff sentinels at the end
// of the row, which keeps the row sorted and matches no item).  Rows of the merged CSR sit at
// rowoff[t] + t * 200 r (rowoff = exclusive scan of the query rows' training-row lengths, computed once), so no
// allocation and no host round trip depends on the data.  The operand matrices are prepared once.  A row whose candidate
// buffer overflows in any round, or whose merged row does not fit the workspace (only possible when `users` repeats a
// user), is flagged "sticky" and recomputed in full by the exact kernel at the end of the call.
constexpr int MM_CAP = TC_KMAX_ROUNDS;
constexpr int MM_SENTINEL = 0x7fffffff;

struct MaskParams {
  const int32_t* users;           // [T] or NULL
  const long long* tr_indptr;     // the caller's training CSR (by user id) or NULL
  const int32_t* tr_indices;
  int T, K, prev, prev_max;       // prev: earlier results per row in this round; prev_max: in the last round
  long long cap;                  // entries of ind2
  long long* rowoff;              // [T + 1]
  long long* indptr2;             // [T + 1] the round's mask CSR, by query row
  int32_t* ind2;
  int32_t* sticky;                // [T]
  const int32_t* out_idx;         // the caller's [T, K]
};

__device__ __forceinline__ long long mm_row_len(const MaskParams& P, int t) {
  if (!P.tr_indptr) return 0;
  const long long u = P.users ? P.users[t] : t;
  return P.tr_indptr[u + 1] - P.tr_indptr[u];
}

__global__ void __launch_bounds__(1024) k_mask_scan(const __grid_constant__ MaskParams P) {
  __shared__ long long s_part[1024];
  const int tid = threadIdx.x;
  const long long per = ((long long)P.T + 1023) / 1024;
  const long long lo = min((long long)P.T, tid * per), hi = min((long long)P.T, lo + per);
  long long sum = 0;
  for (long long t = lo; t < hi; ++t) sum += mm_row_len(P, (int)t);
  s_part[tid] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const long long v = tid >= off ? s_part[tid - off] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  long long run = s_part[tid] - sum;
  for (long long t = lo; t < hi; ++t) {
    P.rowoff[t] = run;
    run += mm_row_len(P, (int)t);
    P.sticky[t] = (run + (t + 1) * P.prev_max > P.cap) ? 1 : 0;   // the row's widest merged form must end inside ind2
  }
  if (tid == 1023) P.rowoff[P.T] = s_part[1023];
}

__device__ __forceinline__ int mm_lower_bound(const int32_t* a, int n, int x) {   // #elements < x of the sorted a[0..n)
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) k_mask_merge(const __grid_constant__ MaskParams P) {
  __shared__ int32_t s_prev[MM_CAP];
  for (int t = blockIdx.x; t < P.T; t += gridDim.x) {
    const long long u = P.users ? P.users[t] : t;
    const long long tlo = P.tr_indptr ? P.tr_indptr[u] : 0;
    const int len = P.tr_indptr ? (int)(P.tr_indptr[u + 1] - tlo) : 0;
    const long long f_lo = P.rowoff[t] + (long long)t * P.prev, f_hi = P.rowoff[t + 1] + (long long)(t + 1) * P.prev;
    const long long lo = min(f_lo, P.cap), hi = min(f_hi, P.cap);
    if (threadIdx.x == 0) {
      P.indptr2[t] = lo;
      if (t == P.T - 1) P.indptr2[P.T] = hi;
    }
    if (hi < f_hi) {   // does not fit (flagged sticky by the scan): a harmless row of sentinels
      for (long long e = lo + threadIdx.x; e < hi; e += blockDim.x) P.ind2[e] = MM_SENTINEL;
      continue;
    }
    const bool dead = P.sticky[t] != 0;   // its earlier results were never written
    int n2 = 32;
    while (n2 < P.prev) n2 <<= 1;
    for (int j = threadIdx.x; j < n2; j += blockDim.x) {
      int id = MM_SENTINEL;
      if (j < P.prev && !dead) {
        id = P.out_idx[(long long)t * P.K + j];
        if (id < 0) id = MM_SENTINEL;
      }
      s_prev[j] = id;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int x = threadIdx.x; x < n2; x += blockDim.x) {
          const int p = x ^ j;
          if (p > x) {
            const int va = s_prev[x], vb = s_prev[p];
            if (((x & k) == 0) ? (va > vb) : (va < vb)) { s_prev[x] = vb; s_prev[p] = va; }
          }
        }
        __syncthreads();
      }
    }
    const int32_t* tr = P.tr_indices + tlo;
    for (int j = threadIdx.x; j < len; j += blockDim.x) {
      const int x = tr[j];
      P.ind2[lo + j + mm_lower_bound(s_prev, P.prev, x)] = x;
    }
    for (int j = threadIdx.x; j < P.prev; j += blockDim.x) {
      const int x = s_prev[j];
      P.ind2[lo + j + (x == MM_SENTINEL ? len : mm_lower_bound(tr, len, x))] = x;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_sticky_or(int32_t* sticky, const int32_t* ovf, int T, int32_t* count) {
  int c = 0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const int s = sticky[t] | ovf[t];
    sticky[t] = s;
    c += s != 0;
  }
  if (count) {   // (last round only) rows handed to the exact kernel
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*encode_tiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_t get_encode() {
  static encode_tiled_t fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_t>(p);
  }
  return fn;
}

int make_map(CUtensorMap* tm, void* base, long long rows, int Kp, int box_rows = TC_M) {
  encode_tiled_t enc = get_encode();
  CF_CHECK_ARG(enc != nullptr, "cf_topk_tc: cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
  const cuuint32_t box[2] = {TC_KCH, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CF_CHECK_ARG(r == CUDA_SUCCESS, "cf_topk_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

struct TcPlan {
  int Kp, KC, S, stages, NB, pair, S_cand;   // S_cand: candidate buffers per row (S item splits, x 2 column halves in the paired kernel)
  long long T_pad, N_pad;
  size_t off_vb, off_qb, off_eps, off_cand, off_ccnt, off_ovf, off_bmax, total;
  size_t smem;
  int rounds;                                // ceil(K / TC_KMAX); the rest is used when rounds > 1
  long long mask_cap;                        // entries of the per-round mask CSR
  size_t off_sticky, off_rowoff, off_indptr2, off_ind2;
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int plan_tc(const cf_topk_args* a, TcPlan* p) {
  const int aug = a->kind == CF_SCORE_DOT ? 0 : 2;
  p->Kp = (int)align_up((size_t)a->d + aug, TC_KCH);
  p->KC = p->Kp / TC_KCH;
  CF_CHECK_ARG(p->KC >= 1 && p->KC <= 4, "cf_topk_tc: n_factors up to 254 are served by the tensor path (d=%d)", a->d);
  // The wide kernel (needs KC <= 2: A 64 KB + two 64 KB stages of B) issues 1.4x the flops per MMA cycle, but with one accumulator per
  // M tile the epilogue of an M tile must finish inside the OTHER M tile's MMA time, and under MMA load a tcgen05.ld of 32 columns
  // takes ~350 cycles (TMEM accumulate traffic crowds the reads out; 118 cycles idle): 8 chunks = 2800 > 1400.  Measured end to
  // end it ties with the narrow kernel (329 k vs 335 k users/s on 10 M items), so it is opt-in: CF_TC_WIDE=1.
  p->NB = TC_N;
  p->pair = 0;
  if (const char* e = getenv("CF_TC_WIDE")) if (atoi(e) > 0 && p->KC <= 2) p->NB = TC_NW;
  // The cta_group::2 kernel is built, tested bit-identical and opt-in (CF_TC_PAIR=1): measured on 1 M users x 10 M items it
  // TIES with the single-CTA kernel to 0.04 % (2956.7 vs 2957.9 ms).  Both are power-bound: nvidia-smi during the sweep shows
  // 990-1015 W with sw_power_cap active and the SM clock at 1.44-1.48 GHz (max 1.965), and reading only HALF of every
  // accumulator (an experiment, wrong results) buys 4.5 %.  At that clock the N = 128 MMA form tops out at ~840 TFLOP/s:
  // the sweep runs at the MMA rate of the throttled clock, and a faster MMA form only lowers the clock further.
  if (const char* e = getenv("CF_TC_PAIR")) if (atoi(e) > 0) { p->pair = 1; p->NB = TC_NW; }
  p->T_pad = (long long)align_up((size_t)a->T, TC_MT * TC_M);
  p->N_pad = (long long)align_up((size_t)a->n_items, p->NB);
  const long long row_tiles = p->T_pad / (TC_MT * TC_M), n_tiles = p->N_pad / p->NB;   // (row_tiles: CTAs, or CTA pairs)
  const long long slots = p->pair ? cf_num_sms() / 2 : cf_num_sms();
  int S = (int)((slots + row_tiles - 1) / row_tiles);   // just enough item splits to give every SM a CTA: each
                                                        // split restarts its rows' thresholds from -inf
  if (S < 1) S = 1;
  if (const char* e = getenv("CF_TC_SPLITS")) S = atoi(e) > 0 ? atoi(e) : S;   // tuning knob
  const int s_max = RR_CAP / TC_CAP / (p->pair ? 2 : 1);
  if (S > s_max) S = s_max;
  if (S > n_tiles) S = (int)n_tiles;
  p->S = S;
  p->S_cand = p->pair ? 2 * S : S;
  const size_t a_bytes = (size_t)(p->pair ? 1 : TC_MT) * p->KC * TC_CHUNK_BYTES;
  const size_t b_bytes = (size_t)p->KC * TC_CHUNK_BYTES * (p->pair ? 1 : p->NB / 128);
  int stages = (int)((200 * 1024 - a_bytes) / b_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  CF_CHECK_ARG(stages >= 2, "cf_topk_tc: not enough shared memory for a 2-stage ring");
  p->stages = stages;
  p->smem = a_bytes + stages * b_bytes + 256 + 1024;
  size_t off = 0;
  p->off_vb = off; off = align_up(off + (size_t)p->N_pad * p->Kp * 2, 1024);
  p->off_qb = off; off = align_up(off + (size_t)p->T_pad * p->Kp * 2, 1024);
  p->off_eps = off; off = align_up(off + (size_t)p->T_pad * 4, 256);
  p->off_cand = off; off = align_up(off + (size_t)p->T_pad * p->S_cand * TC_CAP * 8, 256);
  p->off_ccnt = off; off = align_up(off + (size_t)p->T_pad * p->S_cand * 4, 256);
  p->off_ovf = off; off = align_up(off + (size_t)p->T_pad * 4, 256);
  p->off_bmax = off; off = align_up(off + 256, 256);
  p->rounds = (a->K + TC_KMAX - 1) / TC_KMAX;
  p->mask_cap = 0;
  if (p->rounds > 1) {
    CF_CHECK_ARG(a->train.indptr == nullptr || a->train.nnz >= 0, "cf_topk_tc: train.nnz is required for K > %d", TC_KMAX);
    p->mask_cap = (a->train.indptr ? (long long)a->train.nnz : 0ll) + (long long)a->T * TC_KMAX * (p->rounds - 1);
    p->off_sticky = off; off = align_up(off + (size_t)p->T_pad * 4, 256);
    p->off_rowoff = off; off = align_up(off + ((size_t)a->T + 1) * 8, 256);
    p->off_indptr2 = off; off = align_up(off + ((size_t)a->T + 1) * 8, 256);
    p->off_ind2 = off; off = align_up(off + (size_t)p->mask_cap * 4 + 4, 256);
  }
  p->total = off;
  return 0;
}

int validate_tc(const cf_topk_args* a, const char* who) {
  CF_CHECK_ARG(a != nullptr, "%s: args is NULL", who);
  CF_CHECK_ARG(a->d > 0 && a->ld >= a->d && a->ld % 4 == 0 && a->ld <= 512, "%s: need 0 < d <= ld <= 512, ld %% 4 == 0", who);
  CF_CHECK_ARG(a->T > 0 && a->n_items > 0 && a->n_items < (1ll << 31), "%s: T and n_items must be positive", who);
  CF_CHECK_ARG(a->K > 0 && a->K <= TC_KMAX_ROUNDS, "%s: K must be in [1, %d] for the tensor path (got %d)", who, TC_KMAX_ROUNDS, a->K);
  CF_CHECK_ARG(a->kind >= CF_SCORE_DOT && a->kind <= CF_SCORE_NEG_SQDIST, "%s: unknown scoring kind %d", who, a->kind);
  CF_CHECK_ARG(a->item_lo == 0 && (a->item_hi == 0 || a->item_hi == a->n_items), "%s: item ranges are not supported here (shard V instead)", who);
  return 0;
}

}  // namespace

int cf_topk_exact_flagged(const cf_topk_args* a, const int32_t* only_if_flag, cudaStream_t stream);

extern "C" int64_t cf_topk_tc_workspace_bytes(const cf_topk_args* a) {
  if (validate_tc(a, "cf_topk_tc_workspace_bytes")) return -1;
  TcPlan p;
  if (plan_tc(a, &p)) return -1;
  return (int64_t)p.total;
}

extern "C" int cf_topk_tc(const cf_topk_args* a, void* workspace, int64_t workspace_bytes, float* dbg_scores, int32_t* stats,
                          void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = validate_tc(a, "cf_topk_tc")) return rc;
  CF_CHECK_ARG(a->U && a->V && a->out_idx && workspace, "cf_topk_tc: U, V, out_idx and workspace are required");
  CF_CHECK_ARG(a->kind != CF_SCORE_DOT_BIAS || a->b, "cf_topk_tc: DOT_BIAS needs the bias vector");
  CF_CHECK_ARG(((uintptr_t)workspace % 1024) == 0, "cf_topk_tc: workspace must be 1024-byte aligned");
  TcPlan p;
  if (int rc = plan_tc(a, &p)) return rc;
  CF_CHECK_ARG(workspace_bytes >= (int64_t)p.total, "cf_topk_tc: workspace %lld < required %lld bytes", (long long)workspace_bytes, (long long)p.total);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  __half* Vb = reinterpret_cast<__half*>(ws + p.off_vb);
  __half* Qb = reinterpret_cast<__half*>(ws + p.off_qb);
  float* eps2 = reinterpret_cast<float*>(ws + p.off_eps);
  uint2* cand = reinterpret_cast<uint2*>(ws + p.off_cand);
  int32_t* ccnt = reinterpret_cast<int32_t*>(ws + p.off_ccnt);
  int32_t* ovf = reinterpret_cast<int32_t*>(ws + p.off_ovf);
  float* bmax = reinterpret_cast<float*>(ws + p.off_bmax);
  CF_CUDA_OK(cudaMemsetAsync(bmax, 0, 256, stream));
  CF_CUDA_OK(cudaMemsetAsync(ovf, 0, (size_t)p.T_pad * 4, stream));
  CF_CUDA_OK(cudaMemsetAsync(ccnt, 0, (size_t)p.T_pad * p.S_cand * 4, stream));
  const int sms = cf_num_sms();

  PrepParams pi = {};
  pi.src = a->V; pi.bias = a->b; pi.ids = nullptr; pi.n_valid = a->n_items; pi.n_pad = p.N_pad;
  pi.d = a->d; pi.ld = a->ld; pi.Kp = p.Kp; pi.kind = a->kind; pi.is_query = 0; pi.dst = Vb; pi.bmax = bmax; pi.eps2 = nullptr; pi.stats = stats;
  if (stats) CF_CUDA_OK(cudaMemsetAsync(stats, 0, 16, stream));
  long long g = (p.N_pad + 7) / 8;
  if (g > (long long)sms * 16) g = (long long)sms * 16;
  k_prep<<<(unsigned)g, 256, 0, stream>>>(pi);
  PrepParams pq = pi;
  pq.src = a->U; pq.bias = nullptr; pq.ids = a->users; pq.n_valid = a->T; pq.n_pad = p.T_pad; pq.is_query = 1; pq.dst = Qb; pq.eps2 = eps2;
  g = (p.T_pad + 7) / 8;
  if (g > (long long)sms * 16) g = (long long)sms * 16;
  k_prep<<<(unsigned)g, 256, 0, stream>>>(pq);

  CUtensorMap tmQ, tmV;
  if (int rc = make_map(&tmQ, Qb, p.T_pad, p.Kp)) return rc;
  if (int rc = make_map(&tmV, Vb, p.N_pad, p.Kp, p.pair ? TC_M : p.NB)) return rc;   // (the paired kernel: each CTA loads half a tile)
  TcParams P = {};
  P.T = a->T; P.N = (int)a->n_items; P.KC = p.KC; P.n_tiles = (int)(p.N_pad / p.NB); P.S = p.S; P.stages = p.stages;
  P.users = a->users; P.tr_indptr = (const long long*)a->train.indptr; P.tr_indices = a->train.indices;
  P.eps2 = eps2; P.cand = cand; P.cand_cnt = ccnt; P.overflow = ovf;
  P.dbg_scores = dbg_scores; P.dbg_ld = (long long)align_up((size_t)a->n_items, TC_NW);   // the same stride for both kernels
  {   // threshold warm-up (RowSweep::sweep_warm): at most TC_CAP chunk maxima per buffer, at most an eighth of the split
    const int tiles_per_split = (P.n_tiles + p.S - 1) / p.S;
    int warm = TC_CAP / (p.NB / 32);
    if (warm > tiles_per_split / 8) warm = tiles_per_split / 8;
    if (const char* e = getenv("CF_TC_WARM")) warm = atoi(e) < warm ? atoi(e) : warm;   // tuning knob (0 = off)
    P.warm = (p.pair || warm < 0) ? 0 : warm;
    P.coop = 1;
    if (const char* e = getenv("CF_TC_COOP")) P.coop = atoi(e) > 0;   // tuning knobs
    P.coop_low = 0;                                                   // (0: chosen per round from its K, below)
    if (const char* e = getenv("CF_TC_COOP_LOW")) P.coop_low = atoi(e);
  }
  RerankParams R = {};
  R.U = a->U; R.V = a->V; R.b = a->b; R.ld = a->ld; R.nvec = a->ld / 4; R.kind = a->kind; R.T = a->T; R.S = p.S_cand;
  R.users = a->users; R.cand = cand; R.n_items = (int)a->n_items; R.cand_cnt = ccnt; R.overflow = ovf;
  R.tr_indptr = P.tr_indptr; R.tr_indices = P.tr_indices;
  R.out_idx = a->out_idx; R.out_val = a->out_val; R.stats = stats; R.out_ld = a->K; R.count_overflow = p.rounds == 1;
  int rr_cap = 512;
  while (rr_cap < p.S_cand * TC_CAP) rr_cap <<= 1;
  R.cap = rr_cap;
  const int rr_threads = rr_cap <= 512 ? 128 : 256;
  const size_t rr_smem = (size_t)rr_cap * 12 + 512 * 8;
  int rg = a->T;
  if (rg > sms * 16) rg = sms * 16;

  MaskParams M = {};
  int32_t* sticky = nullptr;
  if (p.rounds > 1) {     // K > TC_KMAX: rounds of TC_KMAX over a growing mask (see k_mask_merge)
    sticky = reinterpret_cast<int32_t*>(ws + p.off_sticky);
    M.users = a->users; M.tr_indptr = P.tr_indptr; M.tr_indices = P.tr_indices;
    M.T = a->T; M.K = a->K; M.prev_max = TC_KMAX * (p.rounds - 1); M.cap = p.mask_cap;
    M.rowoff = reinterpret_cast<long long*>(ws + p.off_rowoff);
    M.indptr2 = reinterpret_cast<long long*>(ws + p.off_indptr2);
    M.ind2 = reinterpret_cast<int32_t*>(ws + p.off_ind2);
    M.sticky = sticky; M.out_idx = a->out_idx;
    k_mask_scan<<<1, 1024, 0, stream>>>(M);
    CF_CUDA_OK(cudaGetLastError());
  }
  const int coop_low_env = P.coop_low;
  for (int r = 0; r < p.rounds; ++r) {
    const int Kr = a->K - r * TC_KMAX < TC_KMAX ? a->K - r * TC_KMAX : TC_KMAX;
    if (r > 0) {
      CF_CUDA_OK(cudaMemsetAsync(ovf, 0, (size_t)p.T_pad * 4, stream));
      CF_CUDA_OK(cudaMemsetAsync(ccnt, 0, (size_t)p.T_pad * p.S_cand * 4, stream));
      M.prev = r * TC_KMAX;
      int mg = a->T;
      if (mg > sms * 8) mg = sms * 8;
      k_mask_merge<<<mg, 256, 0, stream>>>(M);
      CF_CUDA_OK(cudaGetLastError());
      P.tr_indptr = M.indptr2; P.tr_indices = M.ind2; P.mask_by_row = 1; P.dbg_scores = nullptr;
      R.tr_indptr = M.indptr2; R.tr_indices = M.ind2; R.mask_by_row = 1;
    }
    P.K = Kr;
    if (coop_low_env <= 0) P.coop_low = (Kr + 16 + TC_CAP - 128) / 2;   // half way between a compacted row and a full one
    dim3 grid((unsigned)(p.T_pad / (TC_MT * TC_M)), (unsigned)p.S);
    if (p.pair) {
      grid.x *= 2;     // clusters of two CTAs, 128 query rows each
      CF_CUDA_OK(cudaFuncSetAttribute(k_topk_tc_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
      k_topk_tc_pair<<<grid, TC_THREADS, p.smem, stream>>>(tmQ, tmV, P);
    } else if (p.NB == TC_NW) {
      CF_CUDA_OK(cudaFuncSetAttribute(k_topk_tc<TC_NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
      k_topk_tc<TC_NW><<<grid, TC_THREADS, p.smem, stream>>>(tmQ, tmV, P);
    } else {
      CF_CUDA_OK(cudaFuncSetAttribute(k_topk_tc<TC_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
      k_topk_tc<TC_N><<<grid, TC_THREADS, p.smem, stream>>>(tmQ, tmV, P);
    }
    CF_CUDA_OK(cudaGetLastError());
    R.K = Kr; R.out_off = r * TC_KMAX;
    k_rerank<<<rg, rr_threads, rr_smem, stream>>>(R);
    CF_CUDA_OK(cudaGetLastError());
    if (p.rounds > 1) {
      int sg = (a->T + 255) / 256;
      if (sg > sms * 8) sg = sms * 8;
      k_sticky_or<<<sg, 256, 0, stream>>>(sticky, ovf, a->T, (r == p.rounds - 1) ? stats : nullptr);
      CF_CUDA_OK(cudaGetLastError());
    }
  }
  return cf_topk_exact_flagged(a, p.rounds > 1 ? sticky : ovf, stream);
}
