"""utils/IOUtil.loadSparseR (native one-pass parser) vs the reference's own loader semantics: same matrix as the golden
ml-100k fixture (made by the reference's loadSparseR), separators / CRLF / 2-field lines / skipped lines / last-write-wins."""
import os

import numpy as np
import pytest

from collaborativefilteringusingtensorflow_b200.utils import IOUtil, Util


def _reference_loop(inFilePath):
    """The reference's per-line loop (utils/IOUtil.py:9-15), restated here as the checker of the native parser."""
    us, is_, rs = [], [], []
    with open(inFilePath, 'r') as infile:
        for line in infile:
            phs = Util.split_row(line)
            if len(phs) == 2:
                us.append(int(phs[0])); is_.append(int(phs[1])); rs.append(1.0)
            elif len(phs) == 3:
                us.append(int(phs[0])); is_.append(int(phs[1])); rs.append(float(phs[2]))
    return np.asarray(us, dtype=np.int64), np.asarray(is_, dtype=np.int64), np.asarray(rs, dtype=np.float64)


def test_parser_handles_separators_crlf_and_skips(tmp_path):
    p = tmp_path / 'r.txt'
    p.write_text('0\t1\t4.0\r\n2,3,5\n4;5;1.5\n6 7\n\nbad line with five fields x\n1\t1\t2.0\n0\t1\t3.0\n')
    u, i, r = IOUtil.loadTriplets(str(p))
    pu, pi, pr = _reference_loop(str(p))
    np.testing.assert_array_equal(u, pu)
    np.testing.assert_array_equal(i, pi)
    np.testing.assert_array_equal(r, pr)
    assert u.tolist() == [0, 2, 4, 6, 1, 0] and r.tolist() == [4.0, 5.0, 1.5, 1.0, 2.0, 3.0]
    m = IOUtil.loadSparseR(8, 8, str(p))
    assert m[0, 1] == 3.0 and m[6, 7] == 1.0 and m.nnz == 5          # later line overwrites (sR[u, i] = r)
    b = Util.matBinarize(m, 3)
    assert b.dtype == np.float32 and b.nnz == 1        # only (2, 3) = 5 is > 3 (the 4.0 at (0, 1) was overwritten by 3.0)
    with pytest.raises(IndexError):
        IOUtil.loadSparseR(4, 4, str(p))
    with pytest.raises(RuntimeError):
        IOUtil.loadTriplets(str(tmp_path / 'missing.txt'))


def test_roundtrip_of_ml100k_fixture(ml100k, tmp_path):
    u, i, r = ml100k['tra_raw']
    p = tmp_path / 'tra.txt'
    IOUtil.saveTriads(list(zip(u.tolist(), i.tolist(), r.astype(float).tolist())), str(p))
    m = IOUtil.loadSparseR(943, 1682, str(p))
    assert m.nnz == 80000
    b = Util.matBinarize(m, 3)
    assert b.nnz == ml100k['stats']['tra_pos'] == ml100k['tra'].nnz
    assert (b.tocsr() != ml100k['tra'].tocsr()).nnz == 0


@pytest.mark.skipif(not os.path.isdir('/root/reference/data/movielens/ml-100k'), reason='reference not mounted')
def test_same_matrix_as_reference_loader_on_bundled_file(ml100k):
    m = IOUtil.loadSparseR(943, 1682, '/root/reference/data/movielens/ml-100k/ratings__1_tra.txt')
    assert (Util.matBinarize(m, 3).tocsr() != ml100k['tra'].tocsr()).nnz == 0


@pytest.mark.parametrize('line', ['1\t2.5\t3.0', '1\t2\t4x', '12abc\t3\t1', 'a\tb'])
def test_malformed_numbers_are_rejected_like_python_int_float(tmp_path, line):
    """int('2.5') / float('4x') / int('12abc') raise ValueError in the reference's loop (IOUtil.py:12-15); a prefix parse
    (strtoll stopping at the dot) must not be accepted silently."""
    p = tmp_path / 'bad.txt'
    p.write_text('0\t1\t1.0\n' + line + '\n')
    with pytest.raises(RuntimeError, match='malformed line 2'):
        IOUtil.loadTriplets(str(p))
    with pytest.raises(ValueError):
        _reference_loop(str(p))


def test_empty_file_and_exponent_and_sign(tmp_path):
    p = tmp_path / 'e.txt'
    p.write_text('')
    u, i, r = IOUtil.loadTriplets(str(p))
    assert len(u) == len(i) == len(r) == 0
    assert IOUtil.loadSparseR(3, 3, str(p)).nnz == 0
    p.write_text('1 2 1e-1\n+2 0 -3.5\n')
    u, i, r = IOUtil.loadTriplets(str(p))
    pu, pi, pr = _reference_loop(str(p))
    assert u.tolist() == pu.tolist() == [1, 2] and i.tolist() == pi.tolist() == [2, 0] and r.tolist() == pr.tolist() == [0.1, -3.5]
