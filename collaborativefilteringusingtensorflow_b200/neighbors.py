"""Host side of the neighbourhood kernels (csrc/cf_neighbors.cu): cosine top-K neighbours, neighbourhood scores, masked
top-N of dense scores.  Used by ItemCF / UserCF (reference src/models/basic/models/itemcf.py, usercf.py) and by the
user-similarity preprocessing of PRIGP / CPLR (pl/models/prigp.py:60-90, cplr_u.py:64-97).  torch holds the buffers; all
arithmetic is in libcf_b200.so."""
from . import _lib

TIE_HIGH_INDEX_FIRST = 1   # what a stable ascending argsort followed by [-K:] keeps (the reference's argsort is unstable:
                           # ties at the cut are undefined there)


def cosine_topk(csr, K, tie_high_index_first=TIE_HIGH_INDEX_FIRST):
    """The K most similar other ROWS of ``csr`` for every row (cosine over the rows' columns): returns
    (idx int32 [n, K] padded with -1, sim float32 [n, K]) ordered by similarity, on the device of ``csr``."""
    torch = _lib.require_cuda()
    lib = _lib.lib()
    n, K = csr.shape[0], int(K)
    if K <= 0:
        raise ValueError('K must be positive')
    t = csr.transpose()
    dev = csr.indices.device
    g = int(min(n, lib.cf_neighbors_concurrent_rows()))
    idx = torch.empty(n, K, dtype=torch.int32, device=dev)
    sim = torch.empty(n, K, dtype=torch.float32, device=dev)
    a = _lib.NeighborArgs()
    with_values = csr.values is not None
    a.rows, a.cols = csr.as_c(with_values), t.as_c(with_values)
    a.K, a.tie_high_index_first = K, int(tie_high_index_first)
    norms = torch.empty(n, dtype=torch.float32, device=dev)
    scratch = torch.zeros(2 * g, n, dtype=torch.float32, device=dev)
    cand = torch.empty(g, n, dtype=torch.int32, device=dev)
    a.out_idx, a.out_sim, a.norms, a.scratch, a.cand, a.grid_rows = (_lib.ptr(idx), _lib.ptr(sim), _lib.ptr(norms), _lib.ptr(scratch),
                                                                     _lib.ptr(cand), g)
    _lib.check(lib.cf_neighbors(a, torch.cuda.current_stream(dev).cuda_stream), 'cf_neighbors')
    return idx, sim


def neighbor_scores(train_csr, users, nbr_idx, nbr_sim, mode):
    """Dense float64 scores [T, n_items]: mode 'item' = itemcf.py:42-50 (nbr_* are the ITEMS' neighbours), mode 'user' =
    usercf.py:31-44 (nbr_* are the USERS' neighbours)."""
    torch = _lib.require_cuda()
    lib = _lib.lib()
    dev = train_csr.indices.device
    users = users.to(device=dev, dtype=torch.int32).contiguous()
    T = int(users.numel())
    out = torch.zeros(T, train_csr.shape[1], dtype=torch.float64, device=dev)
    a = _lib.NeighborScoreArgs()
    a.train = train_csr.as_c(train_csr.values is not None)
    a.users, a.T, a.K = _lib.ptr(users), T, int(nbr_idx.shape[1])
    a.nbr_idx, a.nbr_sim = _lib.ptr(nbr_idx.contiguous()), _lib.ptr(nbr_sim.contiguous())
    a.mode, a.out_scores = {'item': 0, 'user': 1}[mode], _lib.ptr(out)
    _lib.check(lib.cf_neighbor_scores(a, torch.cuda.current_stream(dev).cuda_stream), 'cf_neighbor_scores')
    return out


def topk_dense(scores, N, users=None, mask_csr=None, tie_high_index_first=TIE_HIGH_INDEX_FIRST, return_values=False):
    """The N best columns of every row of ``scores`` (float64 [T, n], consumed as scratch) outside the row's mask."""
    torch = _lib.require_cuda()
    lib = _lib.lib()
    import ctypes as C
    T, n = int(scores.shape[0]), int(scores.shape[1])
    out = torch.empty(T, int(N), dtype=torch.int32, device=scores.device)
    val = torch.empty(T, int(N), dtype=torch.float64, device=scores.device) if return_values else None
    m = mask_csr.as_c(False) if mask_csr is not None else None
    u = users.to(device=scores.device, dtype=torch.int32).contiguous() if users is not None else None
    _lib.check(lib.cf_topk_dense(_lib.ptr(scores), n, T, int(N), int(tie_high_index_first), _lib.ptr(u),
                                 C.byref(m) if m is not None else None, _lib.ptr(out), _lib.ptr(val),
                                 torch.cuda.current_stream(scores.device).cuda_stream), 'cf_topk_dense')
    return (out, val) if return_values else out
