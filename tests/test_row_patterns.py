"""Row-pattern detection behind VEXB_FMT_PATTERNS (csrc/spmv.cu: find_row_patterns), host only: the Poisson matrices of
the reference benchmark (examples/benchmark.cpp:357-415) collapse to two patterns, and expanding the patterns again
reproduces the matrix exactly -- i.e. the strip can be handed to the CCSR kernel without changing a bit."""
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import ccsr
from vexcl_b200 import _lib as L, gen


def patterns(row, col, val, max_patterns=256):
    row, col = np.ascontiguousarray(row, np.int64), np.ascontiguousarray(col, np.int64)
    val = np.ascontiguousarray(val)
    n = row.size - 1
    m = C.c_size_t(0)
    idx = np.empty(n, np.int32)
    rc = L.lib().vexb_csr_row_patterns(n, row.ctypes.data, 8, col.ctypes.data, 8, val.ctypes.data,
                                       L.F64 if val.dtype == np.float64 else L.F32, max_patterns, C.byref(m), idx.ctypes.data)
    return rc, m.value, idx


@pytest.mark.parametrize("dim,n", [(2, 40), (3, 12)])
def test_poisson_has_two_row_patterns(built, dim, n):
    row, col, val = oracle.poisson(dim, n)
    rc, m, idx = patterns(row, col, val)
    assert rc == 0 and m == 2
    width = np.diff(row)
    assert np.array_equal(idx == idx[0], width == 1)                 # pattern of row 0 = the boundary rows
    # rebuild the CCSR arrays from (idx, first row of each pattern) and expand: the matrix comes back bit for bit
    N = row.size - 1
    firsts = [int(np.flatnonzero(idx == u)[0]) for u in range(m)]
    prow = np.concatenate([[0], np.cumsum([width[f] for f in firsts])])
    pcol = np.concatenate([col[row[f]:row[f + 1]] - f for f in firsts])
    pval = np.concatenate([val[row[f]:row[f + 1]] for f in firsts])
    erow, ecol, eval_ = ccsr.ccsr_to_csr(N, idx, prow, pcol, pval)
    assert np.array_equal(erow, row) and np.array_equal(ecol, col) and np.array_equal(eval_, val)
    x = oracle.uniform_real(5, N)
    assert np.array_equal(ccsr.ccsr_spmv(N, idx, prow, pcol, pval, x), oracle.csr_spmv(row, col, val, x))


def test_pattern_limit_values_and_types(built):
    row, col, val = oracle.tridiagonal(1000)
    rc, m, idx = patterns(row, col, val)
    assert rc == 0 and m == 3 and list(idx[:3]) == [0, 1, 1] and idx[-1] == 2
    val2 = val.copy()
    val2[row[500]] = np.nextafter(val2[row[500]], 10.0)             # one ulp: compared bit for bit, so a new pattern
    assert patterns(row, col, val2)[1] == 4
    val3 = val.copy()
    val3[row[500] + 1] = -0.0                                        # -0.0 is not +0.0
    val4 = val.copy()
    val4[row[500] + 1] = 0.0
    assert patterns(row, col, val3)[1] == 4 and patterns(row, col, val4)[1] == 4
    assert patterns(row, col, val.astype(np.float32))[1] == 3
    rrow, rcol, rval = oracle.random_matrix(2000, 2000, 8, seed=1)
    rc, m, _ = patterns(rrow, rcol, rval)
    assert rc == 4 and m == 0                                        # VEXB_ERR_UNSUPPORTED: more than 256 patterns
    rc, m, _ = patterns(rrow, rcol, rval, max_patterns=1 << 20)
    assert rc == 0 and 1500 < m <= 2000
    rc, m, _ = patterns(np.zeros(1, np.int64), np.zeros(0, np.int64), np.zeros(0))
    assert rc == 0 and m == 0
