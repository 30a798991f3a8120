/*
 * cf_b200.h -- C ABI of libcf_b200.so: the B200 (sm_100a) implementation of the pairwise-ranking
 * training + full-catalog top-K evaluation hot path of BinFuPKU/CollaborativeFilteringUsingTensorflow.
 *
 * The reference has no FFI/plugin boundary of its own (it is pure Python over TensorFlow 1.x); the
 * boundary it does have is "numpy batch in -> sess.run(...) -> numpy out".  Every entry point below
 * replaces one such sess.run / numpy-loop call site; the citation beside each names it
 * (paths relative to the reference checkout, src/...).
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless the name ends in _host.  Nothing is owned or freed here:
 *    tables, batches and workspaces are borrowed from the caller (torch tensors on the Python side).
 *  - Every call is asynchronous on the cudaStream_t passed as `stream` (a void* holding the handle).
 *  - Return value: 0 = launched; <0 = rejected on the host (cf_last_error() has the reason).  Device-side
 *    conditions (index out of range, staging overflow, sampler gave up) are reported through the int32
 *    `flags` word of the workspace (CF_FLAG_*), which the caller reads back when it synchronises.
 *  - Embedding tables are fp32, row-major with a leading dimension `ld` (floats) that is a multiple of 4
 *    and >= d; columns d..ld-1 must be zero and stay zero.  Base pointers are 16-byte aligned.
 */
#ifndef CF_B200_H
#define CF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CF_ABI_VERSION 21

/* models */
enum { CF_MODEL_BPR = 0, CF_MODEL_CML = 1, CF_MODEL_GBPR = 2, CF_MODEL_WRMF = 3 };
/* optimizers: ADAGRAD = TF1 AdagradOptimizer (acc0 = 0.1, no eps; bprmf.py:86, cml.py:127, gbprmf.py:104, wrmf.py:86) */
enum { CF_OPT_ADAGRAD = 0, CF_OPT_SGD = 1 };
/* update discipline: SYNC = the reference's minibatch-synchronous semantics (all gradients at pre-update
 * parameters, duplicate rows summed, one apply per unique row); HOGWILD = per-occurrence racy apply. */
enum { CF_UPDATE_SYNC = 1, CF_UPDATE_HOGWILD = 0 };
/* scoring kinds: bprmf.py:80 / wrmf.py:80 ; gbprmf.py:98 ; cml.py:116 */
enum { CF_SCORE_DOT = 0, CF_SCORE_DOT_BIAS = 1, CF_SCORE_NEG_SQDIST = 2 };
/* device-side condition flags (bitwise OR into workspace flags word) */
enum {
  CF_FLAG_INDEX_RANGE = 1,      /* a batch index was outside its table */
  CF_FLAG_STAGING_FULL = 2,     /* more duplicated rows than staging_rows */
  CF_FLAG_SAMPLER_GAVEUP = 4,   /* rejection sampling found no negative in CF_SAMPLER_MAX_TRIES draws */
  CF_FLAG_TOPK_OVERFLOW = 8     /* candidate buffer overflow in the tensor-core top-K path */
};
#define CF_SAMPLER_MAX_TRIES 256

const char* cf_last_error(void);
int cf_abi_version(void);
/* compiled SASS target, e.g. "sm_100a" */
const char* cf_build_arch(void);

/* ------------------------------------------------------------------------------------------------
 * Training step.  Replaces `sess.run(train_op, feed_dict)` of
 *   bprmf.py:145  (BPRMF),  cml.py:184 (CML, incl. the clip of cml.py:119-129),
 *   gbprmf.py:163 (GBPRMF), wrmf.py:145 (WRMF)
 * for `n_batches` consecutive minibatches of `B` rows each (the reference's inner loop bprmf.py:143-148).
 * Per minibatch it launches a row-occurrence counting kernel, ONE fused gather/gradient/update kernel and a
 * kernel that applies the summed gradient of the rows that occurred more than once.
 * ------------------------------------------------------------------------------------------------ */
#define CF_MAX_PEERS 8
typedef struct cf_step_args {
  /* parameters */
  float* U;          /* [n_users, ld] */
  float* V;          /* [n_items, ld] */
  float* b;          /* [n_items] item bias (GBPR) or NULL */
  float* accU;       /* Adagrad accumulators, same shapes; may be NULL for CF_OPT_SGD */
  float* accV;
  float* accb;
  int64_t n_users;
  int64_t n_items;
  int32_t d;
  int32_t ld;
  /* batches: int32 ids, n_batches*B rows */
  const int32_t* pairs;   /* [n_batches*B, 2] (user, item)                     (useritem_placeholder) */
  const int32_t* negs;    /* [n_batches*B, W] negative items; NULL for WRMF    (negItems_placeholder) */
  const int32_t* group;   /* [n_batches*B, G] group users (GBPR) or NULL       (group_placeholder)    */
  const float* ratings;   /* [n_batches*B] (WRMF) or NULL                      (rating_placeholder)   */
  int32_t B;
  int32_t W;
  int32_t G;
  int32_t n_batches;
  /* hyper-parameters (constructor arguments of the reference classes) */
  int32_t model;           /* CF_MODEL_* */
  int32_t optimizer;       /* CF_OPT_* */
  int32_t update;          /* CF_UPDATE_* */
  int32_t use_rank_weight; /* CML */
  float lr;
  float reg;               /* reg (BPR/GBPR/WRMF) or reg_cov (CML; <=0 disables, cml.py:109) */
  float margin;            /* CML */
  float clip_norm;         /* CML */
  float rho;               /* GBPR */
  float weight;            /* WRMF */
  /* workspace (see cf_step_workspace_sizes); metaU/metaV/slots/staging must be zero before the first
   * call and are returned to zero by every successful call */
  uint32_t* metaU;         /* [n_users]  occurrences of the row in the current minibatch */
  uint32_t* metaV;         /* [n_items] */
  int32_t* slotU;          /* [n_users]  staging slot of a row that occurs more than once */
  int32_t* slotV;          /* [n_items] */
  uint32_t* slot_row;      /* [staging_rows] inverse map: slot -> row id | (item table ? 1u<<31 : 0); 0xffffffff = empty,
                            * must be all-empty before the first call and is returned all-empty */
  float* staging;          /* [staging_rows, ld + 4] gradient staging for rows that occur more than once */
  int64_t staging_rows;
  int32_t* counters;       /* [4]: {reserved, flags, reserved, reserved}, zero-initialised */
  double* loss;            /* [n_batches] summed minibatch loss, or NULL to skip the loss */
  /* multi-GPU exchange mode (NULL / 0 on a single GPU): V, n_items then describe the FETCHED item rows of this
   * minibatch (ids in pairs[:,1] / negs index into them), item rows are not applied locally; their summed gradients
   * are red.added into gradV[n_items, ld] (zeroed by the caller) and travel back to the rows' owners (cf_apply_rows) */
  float* gradV;
  int64_t rank_items;      /* CML rank weight: global number of items (0 = n_items) */
  /* peer-pull variant of the exchange mode (NVLink peer memory, n_peers > 0; gradV is required): the item ids in
   * pairs[:,1] / negs are GLOBAL, n_items is the global item count, item i is read straight from its owner's shard
   * peerV[i % n_peers] at row i / n_peers (stride ld; the pointers come from cf_ipc_open / the local table), V is not
   * used, and the gradient of the occurrence is red.added into gradV[gslot_pos[b]] / gradV[gslot_neg[b, w]] */
  const float* peerV[CF_MAX_PEERS];
  const int32_t* gslot_pos; /* [n_batches*B]    row of gradV of the positive item of pair b */
  const int32_t* gslot_neg; /* [n_batches*B, W] row of gradV of negative w of pair b */
  int32_t n_peers;
  int32_t reserved0;
  /* replicated data-parallel mode (small tables, e.g. GBPR's configs[2]: every GPU holds all tables): with gradU (and
   * gradV = [n_items, ld], gradb = [n_items] for GBPR) set, EVERY row gradient of the minibatch is red.added into the
   * dense gradient tables and nothing is applied; the caller all-reduces them and applies them with cf_apply_dense */
  float* gradU;             /* [n_users, ld] or NULL */
  float* gradb;             /* [n_items] (GBPR) or NULL */
  /* push variant of the peer-pull mode (n_peers > 0, peerG[0] != NULL; gradV / gslot_* are then not used): the gradient of
   * every occurrence of item i is red.added straight into its OWNER's dense gradient table peerG[i % n_peers] at row
   * i / n_peers (stride ld; zero at rest; mapped with cf_ipc_open) -- gradient rows leave over NVLink while item rows
   * arrive, and the owner applies its table locally (cf_exchange_apply with dense_grads) */
  float* peerG[CF_MAX_PEERS];
  /* optional cudaEvent_t recorded on `stream` right after the fused step kernel of the LAST minibatch of the call (before its
   * staged apply): lets the caller start independent work -- the sampler launch of a later minibatch on another stream --
   * under the short kernels that follow, not under the occupancy-bound step kernel */
  void* event_after_step;
} cf_step_args;

int cf_train_steps(const cf_step_args* args, void* stream);
/* same work, but brackets every kernel with CUDA events on `stream`, synchronises, and returns the summed device
 * time (ms) of the counting, fused-step and staged-apply kernels (bench.py's per-kernel roofline numbers) */
int cf_train_steps_profiled(const cf_step_args* args, void* stream, float* ms_count_host, float* ms_step_host,
                            float* ms_apply_host);
/* rows of staging (and entries of slot_row) needed for minibatches of B rows: one per row occurrence */
int64_t cf_step_staging_rows(int32_t model, int32_t B, int32_t W, int32_t G);
/* number of kernels cf_train_steps launches per minibatch (for gpu_launches accounting) */
int32_t cf_step_launches_per_batch(void);

/* Owner side of the multi-GPU exchange: apply n gradient rows (grads[k], stride ldg floats) to rows[k] of this GPU's
 * table shard with the same rule as the fused step (rows received from several GPUs are summed, then applied once;
 * model == CF_MODEL_CML also clips the updated row).  Replaces nothing in the reference (it is single-device). */
typedef struct cf_apply_args {
  float* table;            /* [n_rows, ld] */
  float* acc;              /* Adagrad accumulators or NULL (SGD) */
  int64_t n_rows;
  int32_t d;
  int32_t ld;
  const int32_t* rows;     /* [n] local row ids */
  const float* grads;      /* [n, ldg] */
  int64_t n;
  int32_t ldg;
  int32_t model;
  int32_t optimizer;
  float lr;
  float clip_norm;
  uint32_t* meta;          /* [n_rows] zero */
  int32_t* slot;           /* [n_rows] */
  uint32_t* slot_row;      /* [staging_rows] all 0xffffffff */
  float* staging;          /* [staging_rows, ld + 4] zero */
  int64_t staging_rows;    /* >= n */
  int32_t* counters;       /* [4] */
  /* owner-pull variant (NVLink peer memory, n_segs > 0; grads is not used): the n gradient rows are read in place from
   * the requesters' gradient buffers instead of a receive buffer -- rows [seg_start[p], seg_start[p+1]) of `rows` come
   * from seg_grads[p] (a pointer into requester p's buffer, mapped with cf_ipc_open), consecutive, stride ldg */
  const float* seg_grads[CF_MAX_PEERS];
  int64_t seg_start[CF_MAX_PEERS + 1];
  int32_t n_segs;
  int32_t first_seg;       /* segment to start with (the rows are processed in rotated order): owner r starts at requester
                            * r + 1 so that the P owners do not all read the same requester's buffer at the same time */
} cf_apply_args;
int cf_apply_rows(const cf_apply_args* args, void* stream);

/* CUDA IPC plumbing of the peer-pull mode (one process per GPU on one NVLink / NVSwitch node).  cf_ipc_export fills the
 * 64-byte handle of the allocation that contains devptr and the byte offset of devptr inside it; cf_ipc_open maps a
 * handle exported by another process into this one (peer access enabled lazily) and returns the allocation's base;
 * cf_ipc_close unmaps it.  The reference is single-device: replaces nothing. */
int cf_ipc_export(const void* devptr, void* handle64, int64_t* offset_bytes);
int cf_ipc_open(const void* handle64, void** base);
int cf_ipc_close(void* base);

/* Device-side item exchange of the sharded training step (csrc/cf_exchange.cu): nothing returns to the host and no payload
 * goes through NCCL.  Every rank owns a MAILBOX (counts int32[CF_MAX_PEERS] + req int32[n_ranks, cap]: the unique local rows
 * it requests from every owner), a compact gradient buffer grads[slots, ld] and its item shard; it maps its peers' with
 * cf_ipc_open and passes all of them here (its own at index `rank`).  Per minibatch, with the caller's two cross-GPU barriers
 * (any collective on `stream` that completes only after every rank has enqueued it, e.g. a 1-element NCCL all-reduce):
 *   cf_exchange_route   (dedupe + routing of this rank's item ids: fills its mailbox and slot_pairs / slot_negs / slot_pos =
 *                        the minibatch re-indexed by rows ("slots") of the compact buffers, owner-major)
 *   -- barrier 1 --
 *   cf_exchange_prepare (requester: gather each requested row once from its owner into fetched[slot] when fetched != NULL,
 *                        zero grads[slot], return slot_of to rest; owner: count the requests per row of its shard)
 *   cf_train_steps      (exchange mode: V = fetched, pairs = slot_pairs, negs = slot_negs, gradV = grads;
 *                        or peer-pull mode: global ids, peerV = tables, gslot_pos = slot_pos, gslot_neg = slot_negs)
 *   -- barrier 2 --
 *   cf_exchange_apply   (owner: gradient rows read in place from the requesters' buffers, summed per row, applied once)
 * Replaces nothing in the reference (it is single-device, bprmf.py:131). */
typedef struct cf_exchange_args {
  int32_t n_ranks;
  int32_t rank;
  int64_t n_items_global;
  int64_t cap;             /* entries per (requester, owner) request list: >= min(B * (1 + W), ceil(n_items_global / n_ranks)) */
  /* requester side: this rank's minibatch (GLOBAL item ids in pairs[:, 1] / negs) */
  const int32_t* pairs;    /* [B, 2] */
  const int32_t* negs;     /* [B, W] */
  int32_t B;
  int32_t W;
  int32_t* slot_of;        /* [n_ranks * ceil(n_items_global / n_ranks)] dedupe table, all -1 at rest (returned to rest by prepare) */
  int32_t* slot_pairs;     /* out [B, 2]: (user, slot of the positive item) */
  int32_t* slot_negs;      /* out [B, W] */
  int32_t* slot_pos;       /* out [B] */
  /* every rank's mailbox, gradient buffer and item shard */
  int32_t* counts[CF_MAX_PEERS];      /* int32[CF_MAX_PEERS] */
  int32_t* req[CF_MAX_PEERS];         /* int32[n_ranks, cap] */
  float* grads[CF_MAX_PEERS];         /* float[slots, ld], slots >= min(B * (1 + W), n_items_global) */
  const float* tables[CF_MAX_PEERS];  /* float[rows of the shard, ld] */
  float* fetched;          /* local float[slots, ld], or NULL (peer-pull mode: the step kernel reads the shards itself) */
  int32_t d;
  int32_t ld;
  /* owner side: this rank's shard and the apply rule (as cf_apply_args) */
  float* table;
  float* acc;
  int64_t n_rows;
  int32_t model;
  int32_t optimizer;
  float lr;
  float clip_norm;
  uint32_t* meta;          /* [n_rows] zero at rest */
  int32_t* slot;           /* [n_rows] */
  uint32_t* slot_row;      /* [staging_rows] all 0xffffffff at rest */
  float* staging;          /* [staging_rows, ld + 4] zero at rest */
  int64_t staging_rows;    /* >= n_ranks * cap */
  int64_t* segs;           /* device scratch, int64[4 * CF_MAX_PEERS + 4] */
  int32_t* counters;       /* [4] (flags in [1]) */
  /* push variant (with cf_step_args.peerG): dense_grads = this rank's dense gradient table [n_rows, ld] (zero at rest; the
   * peers red.add into it), touched = int32[n_rows] scratch for the list of rows that received a gradient.  grads / fetched /
   * slot / slot_row / staging are then not used: prepare only counts and lists the requested rows, apply walks the list. */
  float* dense_grads;
  int32_t* touched;
} cf_exchange_args;
int cf_exchange_route(const cf_exchange_args* args, void* stream);
int cf_exchange_prepare(const cf_exchange_args* args, void* stream);
int cf_exchange_apply(const cf_exchange_args* args, void* stream);

/* Replicated data-parallel mode: apply a dense, already all-reduced gradient table grads[n_rows, ldg] to table[n_rows, ld]
 * (rows / n / meta / slot / staging of cf_apply_args are not used; ld == 1 applies a bias vector).  Rows whose gradient
 * is all zero are skipped -- a zero gradient is a no-op for Adagrad and SGD, so this equals the sparse apply of
 * bprmf.py:83-88 -- and the rows that were applied are zeroed in grads. */
int cf_apply_dense(const cf_apply_args* args, void* stream);

/* row <- row * c / max(||row||_2, c) over a whole table: cml.py:119-122 (used once, after the first step) */
int cf_clip_rows(float* table, int64_t n_rows, int32_t d, int32_t ld, float clip_norm, void* stream);

/* ------------------------------------------------------------------------------------------------
 * On-device samplers.  Replace the producer threads of
 *   samplers/sampler_ranking.py:22-37, sampler_uij_ranking.py:22-38, sampler_gbpr.py:25-43,
 *   sampler_rating.py:22-39.
 * Stream position is the counter (seed, epoch, batch): identical arguments give identical batches.
 * Positives: position p of epoch e is training pair perm_{seed,e}(p) (a keyed Feistel bijection on
 * [0,nnz) with cycle walking = the reference's per-epoch shuffle without materialising it); the tail
 * nnz mod B is dropped like sampler_ranking.py:25.  Negatives: Philox4x32-10 uniform draws re-drawn while
 * they hit the user's CSR row (sampler_ranking.py:35-36).
 * ------------------------------------------------------------------------------------------------ */
typedef struct cf_csr {
  const int64_t* indptr;   /* [n_rows + 1] */
  const int32_t* indices;  /* [nnz] sorted within a row */
  const int32_t* rows;     /* [nnz] row id of every entry (COO expansion), or NULL */
  const float* values;     /* [nnz] ratings (sampler_rating) or NULL = all 1.0 */
  int64_t n_rows;
  int64_t n_cols;
  int64_t nnz;
} cf_csr;

typedef struct cf_sample_args {
  cf_csr train;            /* user -> items; rows[] required */
  cf_csr train_t;          /* item -> users (GBPR group sampling); indptr NULL when G == 0 */
  uint64_t seed;
  int64_t epoch;
  int64_t batch0;          /* first minibatch of the epoch to generate */
  int32_t n_batches;
  int32_t B;               /* positives per minibatch */
  int32_t W;               /* negatives per positive (ranking/gbpr/uij) */
  int32_t G;               /* group size (gbpr) */
  int32_t n_neg_rows;      /* rating sampler: int(B*negRatio) extra (user, negative, 0) rows per minibatch */
  int32_t shuffle;         /* ranking: 1 = per-epoch permutation (reference), 0 = file order */
  int32_t* out_pairs;      /* [n_batches*B, 2]            (rating: [n_batches*(B+n_neg_rows), 2]) */
  int32_t* out_negs;       /* [n_batches*B, W] or NULL */
  int32_t* out_group;      /* [n_batches*B, G] or NULL */
  float* out_ratings;      /* rating sampler: [n_batches*(B+n_neg_rows)] or NULL */
  int32_t* flags;          /* device int32, CF_FLAG_* OR-ed in */
  const uint64_t* pair_set;     /* optional: open-addressing set of the training pairs built by cf_pair_set_build; the    */
  int32_t pair_set_bits;        /* negatives' membership test is then ONE probe (a 32-byte sector) instead of a bisection */
  int32_t reserved;             /* of the user's CSR row (its last 3-4 probes are cold: the sampler is bound by them)    */
} cf_sample_args;

/* The set of all (user, item) training pairs as a linear-probing hash table of 2^bits 64-bit slots (key = user * n_cols +
 * item + 1, 0 = empty; bits = cf_pair_set_bits(nnz) keeps the load at or below one half).  Same answers as the bisection. */
int32_t cf_pair_set_bits(int64_t nnz);
int cf_pair_set_build(const cf_csr* train, uint64_t* table, int32_t bits, void* stream);

int cf_sample_ranking(const cf_sample_args* args, void* stream);   /* ranking / uij / gbpr */
int cf_sample_rating(const cf_sample_args* args, void* stream);    /* sampler_rating */

/* ------------------------------------------------------------------------------------------------
 * Full-catalog scoring + masked top-K.  Replaces
 *   sess.run(tf.nn.top_k(self.__predict__, maxsz + topN)) + the Python filter loop of
 *   bprmf.py:90-103 (same in cml.py:131-144, gbprmf.py:108-121, wrmf.py:98-111).
 * out_idx[t, 0..K) = the K best items of user users[t] that are not in its training row, ordered by
 * (score desc, item id asc); scores are fp32 inputs accumulated in fp64 sequentially over k (the
 * oracle's definition, oracle/scoring.py).  Short rows are padded with -1 / -inf.
 * cf_topk_exact runs entirely on CUDA cores in fp64 (reference-quality path, any K <= 1024).
 * ------------------------------------------------------------------------------------------------ */
typedef struct cf_topk_args {
  const float* U;
  const float* V;
  const float* b;          /* item bias for CF_SCORE_DOT_BIAS or NULL */
  int64_t n_users;
  int64_t n_items;
  int32_t d;
  int32_t ld;
  const int32_t* users;    /* [T] query users (test_users), or NULL = 0..T-1 */
  int32_t T;
  int32_t K;
  int32_t kind;            /* CF_SCORE_* */
  cf_csr train;            /* rows to mask; indptr NULL = no masking */
  int32_t* out_idx;        /* [T, K] */
  double* out_val;         /* [T, K] or NULL */
  int32_t* flags;
  /* restrict scoring to items [item_lo, item_hi) (item-sharded multi-GPU evaluation); 0,0 = all */
  int64_t item_lo;
  int64_t item_hi;
} cf_topk_args;

int cf_topk_exact(const cf_topk_args* args, void* stream);

/* Same result as cf_topk_exact (bit-identical indices and fp64 scores), computed by the tensor-core path: fp16
 * tcgen05.mma/TMA scoring with a candidate-superset epilogue (the score matrix never leaves the SM), then an exact fp64
 * re-rank of the candidates; rows whose candidate buffer overflows fall back to the exact kernel inside the same call.
 * K <= 1024, d <= 254.  K > 200 (cml.py:203-211 re-recommends at topN = 1000) runs ceil(K / 200) rounds of the same sweep:
 * round r masks the training items AND the results of the earlier rounds (a per-query-row mask CSR merged on the device),
 * so the rounds concatenate to the exact top-K; args->train.nnz must be set then (it sizes that mask).
 * `workspace` (device, 1024-byte aligned) must hold cf_topk_tc_workspace_bytes(args) bytes.
* dbg_scores: NULL, or [T, round_up(n_items, 256)] to receive the raw fp16-GEMM scores (tests; row stride = n_items rounded
 * up to 256).
 * stats: NULL, or device int32[4] = {rows that fell back to the exact kernel, total candidates re-ranked,
 * float bits of the largest 2*eps, float bits of max_i |b'_i|}. */
int64_t cf_topk_tc_workspace_bytes(const cf_topk_args* args);
int cf_topk_tc(const cf_topk_args* args, void* workspace, int64_t workspace_bytes, float* dbg_scores, int32_t* stats,
               void* stream);

/* the dense score matrix of `__predict__` (bprmf.py:77-81, cml.py:111-117, gbprmf.py:95-99, wrmf.py:77-81) as
 * out_scores[T, n_items] fp64; for small inputs only -- the top-K path never materialises it */
int cf_scores(const cf_topk_args* args, double* out_scores, void* stream);

/* merge P per-shard top-K lists ([P, T, K] idx/val, e.g. after an all-gather) into the global top-K */
int cf_topk_merge(const int32_t* idx, const double* val, int32_t P, int32_t T, int32_t K,
                  int32_t* out_idx, double* out_val, void* stream);

/* ------------------------------------------------------------------------------------------------
 * WRMF by weighted alternating least squares (new solver; the reference's wrmf.py is minibatch Adagrad, which is
 * cf_train_steps with CF_MODEL_WRMF).  One half-sweep solves every row x_u of X:
 *   (Y^T Y + (weight-1) sum_{i in P_u} y_i y_i^T + reg I) x_u = weight * sum_{i in P_u} y_i
 * with a tensor-core Gram (tcgen05, bf16 hi/lo split = fp32-grade) and one shared-memory Cholesky per row.  d <= 128.
 * The item half-sweep is the same call with X <-> Y and the transposed CSR.
 * ------------------------------------------------------------------------------------------------ */
typedef struct cf_als_args {
  float* X;                /* [n_x, ldx] rows to solve (overwritten) */
  const float* Y;          /* [n_y, ldy] fixed factors */
  int64_t n_x;
  int64_t n_y;
  int32_t d;
  int32_t ldx;
  int32_t ldy;
  int32_t reserved;
  const int64_t* indptr;   /* CSR of X's rows: [n_x + 1] */
  const int32_t* indices;  /* observed columns = row ids of Y */
  float weight;            /* confidence of observed pairs (1 elsewhere) */
  float reg;
  void* workspace;         /* device, 1024-byte aligned, cf_als_workspace_bytes(n_y) bytes */
  int64_t workspace_bytes;
} cf_als_args;
int64_t cf_als_workspace_bytes(int64_t n_y);
int cf_als_half_sweep(const cf_als_args* args, void* stream);
/* The two stages of the half-sweep on their own, for the multi-GPU sweep (rows of X sharded, Y replicated):
 * cf_als_gram ACCUMULATES Y^T Y of the given rows into G[128, 128] (row stride 128; the caller zeroes G, every rank
 * passes its slice of Y and the partial Grams are all-reduced); cf_als_solve_rows solves args->X's rows from the
 * complete Gram.  args->workspace is optional there: with cf_als_workspace_bytes(n_y) bytes (and weight > 1) the rows with
 * at most 64 observed columns are solved by the low-rank update  x = p - Z_u^T (I / (weight - 1) + Y_u Z_u^T)^-1 Y_u p,
 * Z = Y (G + reg I)^-1, p = weight * sum z  (an n_u x n_u Cholesky instead of a 128 x 128 one), as cf_als_half_sweep does. */
int cf_als_gram(const float* Y, int64_t n_y, int32_t d, int32_t ldy, float* G, void* workspace, int64_t workspace_bytes,
                void* stream);
int cf_als_solve_rows(const cf_als_args* args, const float* G, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Ranking metrics.  Replaces metrics/ranking.py:11-67 (pre/recall/ndcg/map/mrr) and :75-91 (hr/arhr):
 * per-user values are written to out[T, 8] = {pre, recall, ndcg, map, mrr, hit, rr_of_truth0, n_pred};
 * the mean (CV) or sum (LOOV) over users is taken by the caller, as ranking.py does.
 * pred[T, ldp] holds each user's list (-1 terminated / padded), truth is the test CSR restricted to the
 * same T users (row t = user t).
 * ------------------------------------------------------------------------------------------------ */
int cf_rank_metrics(const int32_t* pred, int32_t T, int32_t ldp, int32_t k, const int64_t* truth_indptr,
                    const int32_t* truth_indices, double* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Rating path.  cf_predict_pairs replaces `__predict` of models/basic/models/mf.py:66-72 (the
 * scores of the fed (user, item) rows; CF_SCORE_* as in cf_topk_args; out-of-range ids set CF_FLAG_INDEX_RANGE in
 * counters[1] and give NaN).  cf_rating_metrics replaces metrics/rating.py:4-17: it ADDS sum |t - p| to sums[0] and
 * sum (t - p)^2 to sums[1] (fp64; the caller zeroes sums, divides by n and takes the root where rating.py does), with
 * p clipped to [lo, hi] first (tf.clip_by_value of mf.py:81; pass -inf / +inf for the plain metrics).
 * pred is float32 (pred_is_f64 = 0) or float64.
 * ------------------------------------------------------------------------------------------------ */
int cf_predict_pairs(const float* U, const float* V, const float* b, int64_t n_users, int64_t n_items, int32_t d,
                     int32_t ld, int32_t score, const int32_t* pairs, int64_t n, float* out, int32_t* counters,
                     void* stream);
int cf_rating_metrics(const void* pred, int32_t pred_is_f64, const double* truth, int64_t n, double lo, double hi,
                      double* sums, void* stream);

/* SVD rating model (models/basic/models/svd.py): pred = sum((U_u K) * V_i) with a d x d kernel matrix K.
 * cf_svd_grads is the gradient-only step of svd.py:52-80: for the B rows (pairs, ratings) it ADDS every pair's row
 * gradients into the dense tables gradU[n_users, ld] / gradV[n_items, ld] and the summed e * U_u (x) V_i into
 * gradK[d, ldk] (all zeroed by the caller) and the minibatch loss into *loss (or NULL); cf_apply_dense then applies the
 * three tables (Adagrad on user_embed, kernel, item_embed: svd.py:74-80).  cf_svd_predict_pairs replaces `__predict`
 * (svd.py:66-72) for the B rows of `pairs` (ratings / grad* / loss unused).  d <= 128. */
typedef struct cf_svd_args {
  const float* U;          /* [n_users, ld] */
  const float* V;          /* [n_items, ld] */
  const float* K;          /* [d, ldk] */
  int64_t n_users;
  int64_t n_items;
  int32_t d;
  int32_t ld;
  int32_t ldk;
  int32_t reserved0;
  const int32_t* pairs;    /* [B, 2] */
  const float* ratings;    /* [B] */
  int64_t B;
  float reg;
  int32_t reserved1;
  float* gradU;
  float* gradV;
  float* gradK;
  double* loss;            /* [1] or NULL */
  int32_t* counters;       /* [4] (flags in [1]) */
} cf_svd_args;
int cf_svd_grads(const cf_svd_args* args, void* stream);
int cf_svd_predict_pairs(const cf_svd_args* args, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * The remaining pairwise family: PRIGP (models/pl/models/prigp.py) and CPLR (cplr_u.py).
 * Both score x_um = <U_u, V_m> + b_m over up to four item slots of a tuple.  cf_tuple_grads is the gradient-only step of
 * prigp.py:92-137 (tuples [B, 5] = (u, i, j, t, k); bias NOT trained, :134) / cplr_u.py:99-144 (tuples [B, 4] = (u, i, t, j),
 * coefs [B, 2] = (coef[u, i], coef[u, t])): it ADDS every tuple's row gradients into the dense tables gradU / gradV (and
 * gradb for CPLR; all zeroed by the caller) and the minibatch loss into *loss (or NULL); cf_apply_dense then applies them
 * (TF1 Adagrad on the touched rows).  Scoring is CF_SCORE_DOT_BIAS (prigp.py:124-128).
 * cf_sample_tuples replaces the producer threads of samplers/sampler_prigp.py:22-52 (PRIGP: `train` with its COO rows, `coef`
 * = the coefficient matrix of prigp.py:83-90 as a CSR with values) and samplers/sampler_uitj_ranking.py:22-38 (CPLR: `coef`
 * of cplr_u.py:89-97, `collab` = its rows minus the positives, `eligible` = the users with positives, collaborative items
 * and room for a negative).  Stream position = (seed, epoch, batch0): identical arguments give identical batches.
 * ------------------------------------------------------------------------------------------------ */
enum { CF_TUPLE_PRIGP = 0, CF_TUPLE_CPLR = 1 };
typedef struct cf_tuple_args {
  const float* U;          /* [n_users, ld] */
  const float* V;          /* [n_items, ld] */
  const float* b;          /* [n_items] */
  int64_t n_users;
  int64_t n_items;
  int32_t d;
  int32_t ld;
  int32_t model;           /* CF_TUPLE_* */
  int32_t reserved;
  const int32_t* tuples;   /* [B, 5] (PRIGP) or [B, 4] (CPLR) */
  const float* coefs;      /* [B, 2] (CPLR) or NULL */
  int64_t B;
  float alpha;
  float beta;              /* CPLR */
  float gamma;             /* CPLR */
  float reg;
  float* gradU;
  float* gradV;
  float* gradb;            /* CPLR; ignored for PRIGP */
  double* loss;            /* [1] or NULL */
  int32_t* counters;       /* [4] (flags in [1]) */
} cf_tuple_args;
int cf_tuple_grads(const cf_tuple_args* args, void* stream);

typedef struct cf_tuple_sample_args {
  cf_csr train;            /* user -> positive items (rows[] required for PRIGP) */
  cf_csr coef;             /* user -> items with a non-zero coefficient, values = the coefficients */
  cf_csr collab;           /* CPLR: user -> collaborative items (coefficient row minus positives); indptr NULL for PRIGP */
  const int32_t* eligible; /* CPLR: users that may be drawn */
  int64_t n_eligible;
  uint64_t seed;
  int64_t epoch;
  int64_t batch0;
  int32_t n_batches;
  int32_t B;
  int32_t model;           /* CF_TUPLE_* */
  int32_t reserved;
  int32_t* out_tuples;     /* [n_batches * B, 5 | 4] */
  float* out_coefs;        /* CPLR: [n_batches * B, 2] */
  int32_t* flags;
} cf_tuple_sample_args;
int cf_sample_tuples(const cf_tuple_sample_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Neighbourhood models and the user-similarity preprocessing of PRIGP / CPLR.
 * cf_neighbors replaces `__calsim__` + `__topk__` of models/basic/models/itemcf.py:19-40 (entities = items, features =
 * users), usercf.py:19-28,36-37 and pl/models/prigp.py:60-81, cplr_u.py:64-87 (entities = users, features = items): for
 * every entity the K most similar other entities by cosine similarity over their feature rows,
 *   sim(a, b) = (<a, b> / |lo|) / |hi|  in IEEE float32, lo < hi the two indices (the order of the reference's in-place
 * divisions, itemcf.py:21-26), diagonal zero, ordered by (sim desc, index -- higher first when tie_high_index_first, what a
 * stable ascending argsort followed by [-K:] keeps; the reference's unstable argsort leaves ties undefined).  Entities with
 * fewer than K positive similarities are padded with -1 / 0.
 * cf_neighbor_scores replaces `__predict__` (itemcf.py:42-50 with mode 0, nbr_* = the items' neighbours; usercf.py:31-44
 * with mode 1, nbr_* = the users' neighbours): out_scores[t, :] += the reference's float64 accumulation of float32
 * products (exact, hence order-independent); the caller zeroes out_scores.
 * cf_topk_dense replaces `np.argsort(predicts)[-maxsz-topN:][::-1]` + the filter loop (itemcf.py:52-66): the N best
 * columns of every dense score row outside the user's training row (scores is used as scratch: taken and masked
 * entries are overwritten with -inf).
 * ------------------------------------------------------------------------------------------------ */
typedef struct cf_neighbor_args {
  cf_csr rows;             /* entity -> features (values NULL = binary) */
  cf_csr cols;             /* its transpose: feature -> entities */
  int32_t K;
  int32_t tie_high_index_first;
  int32_t* out_idx;        /* [n_entities, K] */
  float* out_sim;          /* [n_entities, K] */
  float* norms;            /* scratch [n_entities] */
  float* scratch;          /* [2 * grid_rows, n_entities] zero at rest */
  int32_t* cand;           /* scratch [grid_rows, n_entities] */
  int64_t grid_rows;       /* entities processed concurrently (cf_neighbors_concurrent_rows() is a good value) */
} cf_neighbor_args;
int64_t cf_neighbors_concurrent_rows(void);
int cf_neighbors(const cf_neighbor_args* args, void* stream);

typedef struct cf_neighbor_score_args {
  cf_csr train;            /* user -> items training CSR (values NULL = binary) */
  const int32_t* users;    /* [T] query users or NULL = 0..T-1 */
  int32_t T;
  int32_t K;
  const int32_t* nbr_idx;  /* [n_items, K] (mode 0) or [n_users, K] (mode 1) */
  const float* nbr_sim;
  int32_t mode;
  int32_t reserved;
  double* out_scores;      /* [T, n_items] */
} cf_neighbor_score_args;
int cf_neighbor_scores(const cf_neighbor_score_args* args, void* stream);
int cf_topk_dense(double* scores, int64_t n_items, int32_t T, int32_t N, int32_t tie_high_index_first, const int32_t* users,
                  const cf_csr* mask, int32_t* out_idx, double* out_val, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host-side loader.  Replaces the per-line Python loop of utils/IOUtil.py:7-16 (`loadSparseR`): parses
 * "u<sep>i[<sep>rating]" lines (separators ',' ';' or whitespace, Util.py:5-11) into host arrays allocated with malloc
 * (release each with cf_free_host).  Lines with another number of fields are skipped, like the reference.
 * ------------------------------------------------------------------------------------------------ */
int cf_parse_triplets(const char* path_host, int64_t* n_out_host, int64_t** users_out_host, int64_t** items_out_host,
                      double** ratings_out_host);
void cf_free_host(void* p_host);

#ifdef __cplusplus
}
#endif
#endif /* CF_B200_H */
