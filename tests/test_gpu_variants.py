"""Selectable kernel variants and storage encodings of the sparse strips: the CSR thread-per-row kernel
(spmv.kernel=3), hybrid ELL with 16-bit column offsets (spmv.col16) and row-pattern strips (VEXB_FMT_PATTERNS).
First seen green on a B200 in round 2 (profiles/r02_validate_unverified.log)."""
import numpy as np
import pytest

import oracle
import vexcl_b200 as vx

pytestmark = pytest.mark.gpu


# spmv.kernel: 0 = TMA-staged CTA tiles, 1 = persistent TMA pipeline, 2 = register-staged CTA tiles, 3 = thread per row,
# 4 = warp tiles, 5 = warp rings (per-warp TMA ring), 6 = CTA tiles with the x window in shared memory, -1 = the strip's own choice
@pytest.fixture(params=[3, 4, 5, 6, 0, 1, 2, -1])
def csr_kernel(request, built):
    vx.set_param("spmv.kernel", request.param)
    yield request.param
    vx.set_param("spmv.kernel", -1)


@pytest.fixture
def scalar_csr(built):
    vx.set_param("spmv.kernel", 3)                       # csr_scalar_kernel: one thread per row
    yield
    vx.set_param("spmv.kernel", -1)


def _irregular(n, seed, lo=0, hi=32, long_rows=()):
    """Rows of width U[lo, hi) (plus a few given long ones), sorted distinct columns."""
    rng = np.random.default_rng(seed)
    w = rng.integers(lo, hi, n)
    for i, L_ in long_rows:
        w[i] = L_
    row = np.concatenate([[0], np.cumsum(w)]).astype(np.int64)
    col = np.concatenate([np.sort(rng.choice(n, size=k, replace=False)) for k in w]).astype(np.int64) if row[-1] else np.empty(0, np.int64)
    return row, col, rng.random(int(row[-1])) - 0.5


@pytest.mark.parametrize("nparts", [1, 2, 3])
def test_csr_kernels_match_the_oracle(csr_kernel, ctx1, ctx2, ctx3, nparts):
    ctx = {1: ctx1, 2: ctx2, 3: ctx3}[nparts]
    cases = [oracle.poisson(2, 96), oracle.poisson(3, 20), oracle.random_matrix(3000, 3000, 16, seed=4), oracle.tridiagonal(1024),
             _irregular(5000, 1), _irregular(3000, 2, 20, 90), _irregular(2500, 3, 0, 8, long_rows=((7, 300), (1200, 2400), (2499, 257)))]
    for row, col, val in cases:
        n = row.size - 1
        xh = oracle.uniform_real(9, n)
        A = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_CSR)
        x, y = vx.vector(ctx, xh), vx.vector(ctx, n)
        y.assign(A * x)
        want = oracle.csr_spmv(row, col, val, xh)
        if nparts == 1 and csr_kernel == 3:
            assert np.array_equal(y.read(), want)        # storage-order sums without contraction: exact
        else:
            assert np.all(np.abs(y.read() - want) <= 1e-10 * oracle.csr_absrow(row, col, val, xh))
        y0 = oracle.uniform_real(10, n)
        y.assign(vx.vector(ctx, y0) - 2.0 * (A * x))
        assert np.all(np.abs(y.read() - (y0 - 2.0 * want)) <= 1e-10 * (np.abs(y0) + 2 * oracle.csr_absrow(row, col, val, xh)))


def test_csr_scalar_kernel_single_precision(scalar_csr, ctx1):
    row, col, val = oracle.random_matrix(2000, 2500, 12, seed=8)
    xh = oracle.uniform_real(3, 2500).astype(np.float32)
    A = vx.SpMat(ctx1, 2000, 2500, row, col, val.astype(np.float32), vx.FMT_CSR)
    x, y = vx.vector(ctx1, xh), vx.vector(ctx1, 2000, dtype=np.float32)
    y.assign(A * x)
    want = oracle.csr_spmv(row, col, val.astype(np.float32).astype(np.float64), xh.astype(np.float64))
    assert np.allclose(y.read(), want, rtol=2e-5, atol=1e-5)


@pytest.fixture(params=[1, 0])
def col16(request, built):
    vx.set_param("spmv.col16", request.param)            # hybrid ELL with 16-bit column offsets (read at construction; default on)
    yield
    vx.set_param("spmv.col16", 1)


@pytest.mark.parametrize("nparts", [1, 2, 3])
def test_hell_with_16_bit_columns_matches_the_oracle(col16, ctx1, ctx2, ctx3, nparts):
    ctx = {1: ctx1, 2: ctx2, 3: ctx3}[nparts]
    cases = [oracle.poisson(2, 200), oracle.poisson(3, 24), oracle.tridiagonal(5000),
             oracle.random_matrix(4000, 4000, 16, seed=6),             # band too wide in places: falls back to 32-bit
             oracle.random_matrix(70000, 70000, 8, seed=7)]            # columns further than 32767 from the diagonal: 32-bit
    for row, col, val in cases:
        n = row.size - 1
        xh = oracle.uniform_real(11, n)
        A = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_HELL)
        x, y = vx.vector(ctx, xh), vx.vector(ctx, n)
        y.assign(A * x)
        want = oracle.csr_spmv(row, col, val, xh)
        if nparts == 1:
            assert np.array_equal(y.read(), want)        # only the index encoding changes: same bits as 32-bit columns
        else:
            assert np.all(np.abs(y.read() - want) <= 1e-10 * oracle.csr_absrow(row, col, val, xh))
        y += 3.0 * (A * x)
        assert np.all(np.abs(y.read() - 4.0 * want) <= 1e-10 * 4 * oracle.csr_absrow(row, col, val, xh))


@pytest.mark.parametrize("nparts", [1, 2, 3])
def test_row_pattern_strips_match_the_oracle(built, ctx1, ctx2, ctx3, nparts):
    """VEXB_FMT_PATTERNS: strips with few distinct rows are multiplied by the CCSR kernel (same bits as ELL / CSR)."""
    ctx = {1: ctx1, 2: ctx2, 3: ctx3}[nparts]
    for (row, col, val), compressible in ((oracle.poisson(2, 200), True), (oracle.poisson(3, 24), True),
                                          (oracle.tridiagonal(5000), True), (oracle.random_matrix(3000, 3000, 8, seed=2), False)):
        n = row.size - 1
        xh = oracle.uniform_real(12, n)
        A = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_PATTERNS)
        assert (A.info().loc.fmt == vx.FMT_PATTERNS) == compressible
        if compressible:
            assert A.info().loc.n_tiles <= 3
        x, y = vx.vector(ctx, xh), vx.vector(ctx, n)
        y.assign(A * x)
        want = oracle.csr_spmv(row, col, val, xh)
        if nparts == 1:
            assert np.array_equal(y.read(), want)
        else:
            assert np.all(np.abs(y.read() - want) <= 1e-10 * oracle.csr_absrow(row, col, val, xh))
        y -= 0.5 * (A * x)
        assert np.all(np.abs(y.read() - 0.5 * want) <= 1e-10 * oracle.csr_absrow(row, col, val, xh))


def test_row_patterns_on_a_rectangular_strip_with_rows_left_of_the_diagonal(ctx1):
    """More rows than columns: rows i >= ncols whose entries all lie left of the diagonal, and empty rows, are valid CCSR
    rows (the reach check used to start from the diagonal and rejected them)."""
    n, m = 1000, 900
    w = np.where(np.arange(n) >= 100, 2, 0)
    row = np.concatenate([[0], np.cumsum(w)]).astype(np.int64)
    i = np.arange(100, n)
    col = np.stack([i - 100, np.minimum(i - 99, m - 1)], axis=1).ravel().astype(np.int64)
    val = np.tile([2.0, -1.0], n - 100)
    # the last row repeats column m-1 through the clamp: make it a separate pattern on purpose
    xh = oracle.uniform_real(5, m)
    A = vx.SpMat(ctx1, n, m, row, col, val, vx.FMT_PATTERNS)
    assert A.info().loc.fmt == vx.FMT_PATTERNS and A.info().loc.n_tiles <= 3
    x, y = vx.vector(ctx1, xh), vx.vector(ctx1, n)
    y.assign(7.0)
    y.assign(A * x)
    assert np.array_equal(y.read(), oracle.csr_spmv(row, col, val, xh))


@pytest.mark.parametrize("nparts", [1, 2, 3])
def test_sliced_ell_matches_the_oracle(built, ctx1, ctx2, ctx3, nparts):
    """VEXB_FMT_SELL (SELL-32-sigma): slices of 32 rows sorted by length inside windows; one lane per row, products added
    in storage order -> the same bits as the reference loop on one slice.  Also what AUTO picks for uneven rows."""
    ctx = {1: ctx1, 2: ctx2, 3: ctx3}[nparts]
    cases = [oracle.poisson(2, 96), oracle.random_matrix(3000, 3000, 16, seed=4), oracle.tridiagonal(1030),
             _irregular(5000, 1), _irregular(3001, 2, 20, 90), _irregular(2500, 3, 0, 8, long_rows=((7, 300), (1200, 2400), (2499, 257))),
             _irregular(33, 5, 0, 4), _irregular(40, 6, 0, 3)]
    for sigma in (1024, 32):
        vx.set_param("spmv.sell_sigma", sigma)
        for row, col, val in cases:
            n = row.size - 1
            xh = oracle.uniform_real(9, n)
            A = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_SELL)
            assert A.info().loc.fmt == vx.FMT_SELL or A.info().loc.nnz == 0
            x, y = vx.vector(ctx, xh), vx.vector(ctx, n)
            y.assign(7.0)
            y.assign(A * x)
            want = oracle.csr_spmv(row, col, val, xh)
            if nparts == 1:
                assert np.array_equal(y.read(), want)
            else:
                assert np.all(np.abs(y.read() - want) <= 1e-10 * oracle.csr_absrow(row, col, val, xh))
            y0 = oracle.uniform_real(10, n)
            y.assign(vx.vector(ctx, y0) - 2.0 * (A * x))
            assert np.all(np.abs(y.read() - (y0 - 2.0 * want)) <= 1e-10 * (np.abs(y0) + 2 * oracle.csr_absrow(row, col, val, xh)))
    vx.set_param("spmv.sell_sigma", 1024)
    # AUTO: even rows -> hybrid ELL, uneven rows -> sliced ELL
    row, col, val = oracle.poisson(2, 64)
    assert vx.SpMat(ctx1, row.size - 1, row.size - 1, row, col, val).info().loc.fmt == vx.FMT_HELL
    row, col, val = _irregular(4000, 7)
    assert vx.SpMat(ctx1, 4000, 4000, row, col, val).info().loc.fmt == vx.FMT_SELL
