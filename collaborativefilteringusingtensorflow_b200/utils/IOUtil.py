"""reference src/utils/IOUtil.py:7-25: ``u<sep>i[<sep>r]`` text -> scipy lil_matrix, and the inverse writer.

The reference inserts one element at a time into a lil_matrix; here the file is parsed in one pass into COO arrays
(later entries for the same cell overwrite earlier ones, like the reference's repeated ``sR[u, i] = r``)."""
import numpy as np
from scipy.sparse import coo_matrix, lil_matrix


def loadTriplets(inFilePath):
    """(users, items, ratings) of every 2- or 3-field line, parsed by the native one-pass parser (cf_parse_triplets)."""
    import ctypes as C
    from .. import _lib
    lib = _lib.lib()
    n = C.c_int64(0)
    pu, pi, pr = C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)(), C.POINTER(C.c_double)()
    _lib.check(lib.cf_parse_triplets(str(inFilePath).encode(), C.byref(n), C.byref(pu), C.byref(pi), C.byref(pr)),
               'cf_parse_triplets')
    try:
        k = n.value
        u = np.ctypeslib.as_array(pu, shape=(k,)).copy() if k else np.zeros(0, np.int64)
        i = np.ctypeslib.as_array(pi, shape=(k,)).copy() if k else np.zeros(0, np.int64)
        r = np.ctypeslib.as_array(pr, shape=(k,)).copy() if k else np.zeros(0, np.float64)
    finally:
        for p in (pu, pi, pr):
            lib.cf_free_host(C.cast(p, C.c_void_p))
    return u, i, r


def loadSparseR(usernum, itemnum, inFilePath):
    u, i, r = loadTriplets(inFilePath)
    if len(u) and (u.min() < 0 or u.max() >= usernum or i.min() < 0 or i.max() >= itemnum):
        raise IndexError('row/column index out of bounds')
    key = u * itemnum + i
    _, last = np.unique(key[::-1], return_index=True)        # last write wins
    keep = len(key) - 1 - last
    sR = coo_matrix((r[keep], (u[keep], i[keep])), shape=(usernum, itemnum)).tolil()
    return lil_matrix(sR)


def saveTriads(triads, outFilePath, isRatingInt=False):
    with open(outFilePath, 'w') as outfile:
        for user, item, rating in triads:
            outfile.write(str(int(user)) + '\t' + str(int(item)) + ('\t%d\n' % rating if isRatingInt else '\t%.1f\n' % rating))
