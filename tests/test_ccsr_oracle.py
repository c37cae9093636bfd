"""Pins oracle/ccsr.py (CPU, no GPU): the reference's CCSR test matrix (tests/spmv.cpp:150-197, 3-D Poisson n=32 with
two unique rows) must give, bit for bit, the same product as (a) a literal transcription of the reference loop
(ccsr.hpp:41-47), (b) the CSR oracle on the expanded matrix, which is itself pinned against the reference generator
(examples/benchmark.cpp:357-415) in test_oracle_kat.py."""
import numpy as np

import oracle
from oracle import ccsr
from vexcl_b200 import gen


def test_ccsr_oracle_on_the_reference_test_matrix():
    n = 12
    N = n ** 3
    idx, row, col, val = gen.poisson_ccsr(n)
    x = oracle.uniform_real(7, N)
    y = ccsr.ccsr_spmv(N, idx, row, col, val, x)
    assert np.array_equal(y, ccsr.ccsr_spmv_loop(N, idx, row, col, val, x))
    # the same matrix as the CSR generator of the reference benchmark
    prow, pcol, pval = oracle.poisson(3, n)
    erow, ecol, eval_ = ccsr.ccsr_to_csr(N, idx, row, col, val)
    assert np.array_equal(erow, prow) and np.array_equal(ecol, pcol) and np.array_equal(eval_, pval)
    assert np.array_equal(y, oracle.csr_spmv(prow, pcol, pval, x))
    # boundary rows are identity, interior rows of a constant vector cancel
    ones = ccsr.ccsr_spmv(N, idx, row, col, val, np.ones(N))
    assert np.all(ones[idx == 0] == 1) and np.all(ones[idx == 1] == 0)
    # alpha / append
    y0 = oracle.uniform_real(8, N)
    assert np.array_equal(ccsr.ccsr_spmv(N, idx, row, col, val, x, y=y0, alpha=-2.0, append=True), y0 + (-2.0) * y)


def test_ccsr_oracle_general_unique_rows():
    rng = np.random.default_rng(3)
    N, m = 500, 6
    widths = rng.integers(0, 6, m)
    row = np.concatenate([[0], np.cumsum(widths)]).astype(np.uint64)
    col = rng.integers(-3, 4, int(row[-1])).astype(np.int64)
    val = rng.random(int(row[-1]))
    idx = rng.integers(0, m, N).astype(np.uint64)
    idx[:3] = idx[-3:] = int(np.argmin(widths)) if widths.min() == 0 else 0     # keep the ends inside the vector
    if widths.min() != 0:
        col[int(row[0]):int(row[1])] = 0
    x = rng.random(N)
    assert np.array_equal(ccsr.ccsr_spmv(N, idx, row, col, val, x), ccsr.ccsr_spmv_loop(N, idx, row, col, val, x))
