"""ctypes binding of libcf_b200.so (C ABI declared in include/cf_b200.h).

There is NO CPU fallback: if the CUDA library is missing or the ABI version differs this raises, and every
product entry point goes through :func:`lib`.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libcf_b200.so')
ABI_VERSION = 21

# enums of cf_b200.h
MODEL_BPR, MODEL_CML, MODEL_GBPR, MODEL_WRMF = 0, 1, 2, 3
OPT_ADAGRAD, OPT_SGD = 0, 1
UPDATE_SYNC, UPDATE_HOGWILD = 1, 0
SCORE_DOT, SCORE_DOT_BIAS, SCORE_NEG_SQDIST = 0, 1, 2
FLAG_INDEX_RANGE, FLAG_STAGING_FULL, FLAG_SAMPLER_GAVEUP, FLAG_TOPK_OVERFLOW = 1, 2, 4, 8

_p = C.c_void_p


MAX_PEERS = 8


class StepArgs(C.Structure):
    _fields_ = [
        ('U', _p), ('V', _p), ('b', _p), ('accU', _p), ('accV', _p), ('accb', _p),
        ('n_users', C.c_int64), ('n_items', C.c_int64), ('d', C.c_int32), ('ld', C.c_int32),
        ('pairs', _p), ('negs', _p), ('group', _p), ('ratings', _p),
        ('B', C.c_int32), ('W', C.c_int32), ('G', C.c_int32), ('n_batches', C.c_int32),
        ('model', C.c_int32), ('optimizer', C.c_int32), ('update', C.c_int32), ('use_rank_weight', C.c_int32),
        ('lr', C.c_float), ('reg', C.c_float), ('margin', C.c_float), ('clip_norm', C.c_float),
        ('rho', C.c_float), ('weight', C.c_float),
        ('metaU', _p), ('metaV', _p), ('slotU', _p), ('slotV', _p), ('slot_row', _p), ('staging', _p),
        ('staging_rows', C.c_int64), ('counters', _p), ('loss', _p), ('gradV', _p), ('rank_items', C.c_int64),
        ('peerV', _p * MAX_PEERS), ('gslot_pos', _p), ('gslot_neg', _p), ('n_peers', C.c_int32), ('reserved0', C.c_int32),
        ('gradU', _p), ('gradb', _p), ('peerG', _p * MAX_PEERS), ('event_after_step', _p),
    ]


class ApplyArgs(C.Structure):
    _fields_ = [
        ('table', _p), ('acc', _p), ('n_rows', C.c_int64), ('d', C.c_int32), ('ld', C.c_int32),
        ('rows', _p), ('grads', _p), ('n', C.c_int64), ('ldg', C.c_int32), ('model', C.c_int32),
        ('optimizer', C.c_int32), ('lr', C.c_float), ('clip_norm', C.c_float),
        ('meta', _p), ('slot', _p), ('slot_row', _p), ('staging', _p), ('staging_rows', C.c_int64), ('counters', _p),
        ('seg_grads', _p * MAX_PEERS), ('seg_start', C.c_int64 * (MAX_PEERS + 1)), ('n_segs', C.c_int32), ('first_seg', C.c_int32),
    ]


class ExchangeArgs(C.Structure):
    _fields_ = [
        ('n_ranks', C.c_int32), ('rank', C.c_int32), ('n_items_global', C.c_int64), ('cap', C.c_int64),
        ('pairs', _p), ('negs', _p), ('B', C.c_int32), ('W', C.c_int32),
        ('slot_of', _p), ('slot_pairs', _p), ('slot_negs', _p), ('slot_pos', _p),
        ('counts', _p * MAX_PEERS), ('req', _p * MAX_PEERS), ('grads', _p * MAX_PEERS), ('tables', _p * MAX_PEERS),
        ('fetched', _p), ('d', C.c_int32), ('ld', C.c_int32),
        ('table', _p), ('acc', _p), ('n_rows', C.c_int64), ('model', C.c_int32), ('optimizer', C.c_int32),
        ('lr', C.c_float), ('clip_norm', C.c_float),
        ('meta', _p), ('slot', _p), ('slot_row', _p), ('staging', _p), ('staging_rows', C.c_int64), ('segs', _p), ('counters', _p),
        ('dense_grads', _p), ('touched', _p),
    ]


class SvdArgs(C.Structure):
    _fields_ = [
        ('U', _p), ('V', _p), ('K', _p), ('n_users', C.c_int64), ('n_items', C.c_int64),
        ('d', C.c_int32), ('ld', C.c_int32), ('ldk', C.c_int32), ('reserved0', C.c_int32),
        ('pairs', _p), ('ratings', _p), ('B', C.c_int64), ('reg', C.c_float), ('reserved1', C.c_int32),
        ('gradU', _p), ('gradV', _p), ('gradK', _p), ('loss', _p), ('counters', _p),
    ]


class AlsArgs(C.Structure):
    _fields_ = [
        ('X', _p), ('Y', _p), ('n_x', C.c_int64), ('n_y', C.c_int64), ('d', C.c_int32), ('ldx', C.c_int32),
        ('ldy', C.c_int32), ('reserved', C.c_int32), ('indptr', _p), ('indices', _p), ('weight', C.c_float),
        ('reg', C.c_float), ('workspace', _p), ('workspace_bytes', C.c_int64),
    ]


class Csr(C.Structure):
    _fields_ = [('indptr', _p), ('indices', _p), ('rows', _p), ('values', _p),
                ('n_rows', C.c_int64), ('n_cols', C.c_int64), ('nnz', C.c_int64)]


TUPLE_PRIGP, TUPLE_CPLR = 0, 1


class TupleArgs(C.Structure):
    _fields_ = [('U', _p), ('V', _p), ('b', _p), ('n_users', C.c_int64), ('n_items', C.c_int64), ('d', C.c_int32), ('ld', C.c_int32),
                ('model', C.c_int32), ('reserved', C.c_int32), ('tuples', _p), ('coefs', _p), ('B', C.c_int64),
                ('alpha', C.c_float), ('beta', C.c_float), ('gamma', C.c_float), ('reg', C.c_float),
                ('gradU', _p), ('gradV', _p), ('gradb', _p), ('loss', _p), ('counters', _p)]


class TupleSampleArgs(C.Structure):
    _fields_ = [('train', Csr), ('coef', Csr), ('collab', Csr), ('eligible', _p), ('n_eligible', C.c_int64),
                ('seed', C.c_uint64), ('epoch', C.c_int64), ('batch0', C.c_int64), ('n_batches', C.c_int32), ('B', C.c_int32),
                ('model', C.c_int32), ('reserved', C.c_int32), ('out_tuples', _p), ('out_coefs', _p), ('flags', _p)]


class NeighborArgs(C.Structure):
    _fields_ = [('rows', Csr), ('cols', Csr), ('K', C.c_int32), ('tie_high_index_first', C.c_int32),
                ('out_idx', _p), ('out_sim', _p), ('norms', _p), ('scratch', _p), ('cand', _p), ('grid_rows', C.c_int64)]


class NeighborScoreArgs(C.Structure):
    _fields_ = [('train', Csr), ('users', _p), ('T', C.c_int32), ('K', C.c_int32), ('nbr_idx', _p), ('nbr_sim', _p),
                ('mode', C.c_int32), ('reserved', C.c_int32), ('out_scores', _p)]


class SampleArgs(C.Structure):
    _fields_ = [
        ('train', Csr), ('train_t', Csr), ('seed', C.c_uint64), ('epoch', C.c_int64), ('batch0', C.c_int64),
        ('n_batches', C.c_int32), ('B', C.c_int32), ('W', C.c_int32), ('G', C.c_int32),
        ('n_neg_rows', C.c_int32), ('shuffle', C.c_int32),
        ('out_pairs', _p), ('out_negs', _p), ('out_group', _p), ('out_ratings', _p), ('flags', _p),
        ('pair_set', _p), ('pair_set_bits', C.c_int32), ('reserved', C.c_int32),
    ]


class TopkArgs(C.Structure):
    _fields_ = [
        ('U', _p), ('V', _p), ('b', _p), ('n_users', C.c_int64), ('n_items', C.c_int64),
        ('d', C.c_int32), ('ld', C.c_int32), ('users', _p), ('T', C.c_int32), ('K', C.c_int32), ('kind', C.c_int32),
        ('train', Csr), ('out_idx', _p), ('out_val', _p), ('flags', _p),
        ('item_lo', C.c_int64), ('item_hi', C.c_int64),
    ]


_SIGNATURES = {
    'cf_last_error': (C.c_char_p, []),
    'cf_abi_version': (C.c_int, []),
    'cf_build_arch': (C.c_char_p, []),
    'cf_train_steps': (C.c_int, [C.POINTER(StepArgs), _p]),
    'cf_train_steps_profiled': (C.c_int, [C.POINTER(StepArgs), _p, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                        C.POINTER(C.c_float)]),
    'cf_step_staging_rows': (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    'cf_step_launches_per_batch': (C.c_int32, []),
    'cf_apply_rows': (C.c_int, [C.POINTER(ApplyArgs), _p]),
    'cf_exchange_route': (C.c_int, [C.POINTER(ExchangeArgs), _p]),
    'cf_exchange_prepare': (C.c_int, [C.POINTER(ExchangeArgs), _p]),
    'cf_exchange_apply': (C.c_int, [C.POINTER(ExchangeArgs), _p]),
    'cf_neighbors_concurrent_rows': (C.c_int64, []),
    'cf_neighbors': (C.c_int, [C.POINTER(NeighborArgs), _p]),
    'cf_neighbor_scores': (C.c_int, [C.POINTER(NeighborScoreArgs), _p]),
    'cf_topk_dense': (C.c_int, [_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _p, C.POINTER(Csr), _p, _p, _p]),
    'cf_tuple_grads': (C.c_int, [C.POINTER(TupleArgs), _p]),
    'cf_sample_tuples': (C.c_int, [C.POINTER(TupleSampleArgs), _p]),
    'cf_clip_rows': (C.c_int, [_p, C.c_int64, C.c_int32, C.c_int32, C.c_float, _p]),
    'cf_predict_pairs': (C.c_int, [_p, _p, _p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _p, C.c_int64, _p, _p, _p]),
    'cf_rating_metrics': (C.c_int, [_p, C.c_int32, _p, C.c_int64, C.c_double, C.c_double, _p, _p]),
    'cf_svd_grads': (C.c_int, [C.POINTER(SvdArgs), _p]),
    'cf_svd_predict_pairs': (C.c_int, [C.POINTER(SvdArgs), _p, _p]),
    'cf_apply_dense': (C.c_int, [C.POINTER(ApplyArgs), _p]),
    'cf_ipc_export': (C.c_int, [_p, _p, C.POINTER(C.c_int64)]),
    'cf_ipc_open': (C.c_int, [_p, C.POINTER(C.c_void_p)]),
    'cf_ipc_close': (C.c_int, [_p]),
    'cf_sample_ranking': (C.c_int, [C.POINTER(SampleArgs), _p]),
    'cf_sample_rating': (C.c_int, [C.POINTER(SampleArgs), _p]),
    'cf_pair_set_bits': (C.c_int32, [C.c_int64]),
    'cf_pair_set_build': (C.c_int, [C.POINTER(Csr), _p, C.c_int32, _p]),
    'cf_topk_exact': (C.c_int, [C.POINTER(TopkArgs), _p]),
    'cf_topk_tc_workspace_bytes': (C.c_int64, [C.POINTER(TopkArgs)]),
    'cf_topk_tc': (C.c_int, [C.POINTER(TopkArgs), _p, C.c_int64, _p, _p, _p]),
    'cf_scores': (C.c_int, [C.POINTER(TopkArgs), _p, _p]),
    'cf_topk_merge': (C.c_int, [_p, _p, C.c_int32, C.c_int32, C.c_int32, _p, _p, _p]),
    'cf_als_workspace_bytes': (C.c_int64, [C.c_int64]),
    'cf_als_half_sweep': (C.c_int, [C.POINTER(AlsArgs), _p]),
    'cf_als_gram': (C.c_int, [_p, C.c_int64, C.c_int32, C.c_int32, _p, _p, C.c_int64, _p]),
    'cf_als_solve_rows': (C.c_int, [C.POINTER(AlsArgs), _p, _p]),
    'cf_parse_triplets': (C.c_int, [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.POINTER(C.c_int64)),
                                  C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.POINTER(C.c_double))]),
    'cf_free_host': (None, [_p]),
    'cf_rank_metrics': (C.c_int, [_p, C.c_int32, C.c_int32, C.c_int32, _p, _p, _p, _p]),
}

_lib = None


class CudaLibraryError(RuntimeError):
    pass


def lib():
    """The loaded CUDA library; raises CudaLibraryError (never falls back) if it cannot be used."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CudaLibraryError(
            '%s is missing: build it with `python -m collaborativefilteringusingtensorflow_b200.build` '
            '(needs nvcc, targets sm_100a). There is no CPU fallback.' % LIB_PATH)
    try:
        h = C.CDLL(LIB_PATH)
    except OSError as e:
        raise CudaLibraryError('cannot load %s: %s' % (LIB_PATH, e))
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(h, name)
        except AttributeError:
            raise CudaLibraryError('%s does not export %s (stale build?)' % (LIB_PATH, name))
        fn.restype, fn.argtypes = res, args
    if h.cf_abi_version() != ABI_VERSION:
        raise CudaLibraryError('ABI mismatch: library %d, binding %d; rebuild' % (h.cf_abi_version(), ABI_VERSION))
    _lib = h
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc, what):
    if rc != 0:
        raise RuntimeError('%s failed (%d): %s' % (what, rc, lib().cf_last_error().decode()))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise CudaLibraryError('no CUDA device: this package runs on a B200 only (no CPU fallback)')
    return torch
