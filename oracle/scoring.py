"""CPU oracle: full-catalog scoring + masked top-N.  TEST INFRASTRUCTURE.

Restates /root/reference/src/models/pl/models/bprmf.py:77-103 (``__predict__`` + ``__recommend``),
cml.py:111-117,131-144, gbprmf.py:95-121 and basic/models/wrmf.py:77-111:

    scores[T, n_items] -> tf.nn.top_k(scores, max_u |train(u)| + topN)  (sorted, ties -> lower index)
    -> per user drop the training items, keep the first topN.

which equals "mask training items, take top-N by (score desc, index asc)" (SURVEY.md 3.4); both
forms are implemented here and tested equal.

Scoring definition (the contract the CUDA path is bit-exact against): fp32 embeddings, products and
sums in float64, accumulated sequentially over the factor index k = 0..d-1.  A product of two fp32
values is exact in fp64, so the result does not depend on FMA contraction, and ordering ties can only
come from (near-)identical rows.  PARITY UNPINNED against TensorFlow's own sgemm/top_k (not installable).
"""
import numpy as np

DOT, DOT_BIAS, NEG_SQDIST = 0, 1, 2


def scores_f64(Uq, V, kind=DOT, bias=None):
    """[T, n_items] float64 scores; sequential-k accumulation."""
    Uq = np.asarray(Uq, dtype=np.float32).astype(np.float64)
    Vd = np.asarray(V, dtype=np.float32).astype(np.float64)
    s = np.zeros((Uq.shape[0], Vd.shape[0]), dtype=np.float64)
    for k in range(Uq.shape[1]):
        if kind == NEG_SQDIST:          # cml.py:116  -sum_k (u_k - v_k)^2
            df = Uq[:, k, None] - Vd[None, :, k]
            s += df * df
        else:                           # bprmf.py:80 / gbprmf.py:98 / wrmf.py:80
            s += Uq[:, k, None] * Vd[None, :, k]
    if kind == NEG_SQDIST:
        s = -s
    if kind == DOT_BIAS:
        s = s + np.asarray(bias, dtype=np.float32).astype(np.float64)[None, :]
    return s


def topn_masked(scores, train_sets, topn):
    """Mask each user's training items, return top-N indices by (score desc, index asc)."""
    out = np.full((scores.shape[0], topn), -1, dtype=np.int64)
    idx = np.arange(scores.shape[1])
    for t in range(scores.shape[0]):
        s = scores[t].copy()
        tr = np.fromiter(train_sets[t], dtype=np.int64, count=len(train_sets[t]))
        s[tr] = -np.inf
        order = np.lexsort((idx, -s))
        order = order[np.isfinite(s[order])][:topn]
        out[t, :len(order)] = order
    return out


def recommend_reference_form(scores, train_sets, topn):
    """The reference's own two-stage form (bprmf.py:90-103): top-K' then Python filter."""
    maxsz = max(len(s) for s in train_sets)
    kp = min(maxsz + topn, scores.shape[1])
    idx = np.arange(scores.shape[1])
    out = []
    for t in range(scores.shape[0]):
        order = np.lexsort((idx, -scores[t]))[:kp]     # tf.nn.top_k: sorted, ties -> lower index
        keep = []
        for y in order:
            if int(y) not in train_sets[t]:
                keep.append(int(y))
            if len(keep) >= topn:
                break
        out.append(keep)
    return out
