"""Parity of the SpMV path with the oracle, through the C ABI.

Mirrors tests/spmv.cpp of the reference (vector_product :10-59, non_square_matrix :61,
non_default_types :89-114, empty_rows :116-146) and tests/sparse_matrices.cpp:153-193
(tridiagonal, all rows checked).  Tolerance 1e-10 relative = BOOST_CHECK_CLOSE(..., 1e-8 %);
the thread-per-row CSR stream path is additionally bit-exact against the unfused oracle.
"""
import ctypes as C

import numpy as np
import pytest

import oracle
import vexcl_b200 as vx
from vexcl_b200 import _lib as L

pytestmark = pytest.mark.gpu

FMTS = [L.FMT_CSR, L.FMT_HELL, L.FMT_AUTO]


def close(a, b, scale=None, tol=1e-10):
    scale = np.abs(b) if scale is None else scale
    return np.all(np.abs(a - b) <= tol * scale + 1e-300)


@pytest.mark.parametrize("fmt", FMTS)
def test_vector_product(ctx, fmt):
    n = 1024
    row, col, val = oracle.random_matrix(n, n, 16, seed=1)
    X = oracle.uniform_real(2, n)
    A = vx.SpMat(ctx, n, n, row, col, val, fmt)
    x, y = vx.vector(ctx, X), vx.vector(ctx, n)
    ref = oracle.csr_spmv(row, col, val, X)
    y.assign(A * x)
    assert close(y.read(), ref)
    y -= A * x
    assert np.all(np.abs(y.read()) <= 1e-10)                          # BOOST_CHECK_SMALL(a, 1e-8)
    y += 42 * (A * x)
    assert close(y.read(), 42 * ref, tol=2e-10)
    y.assign(x + A * x)
    assert close(y.read(), X + ref)
    y.assign(x - 0.5 * (A * x) + 2.0)
    assert close(y.read(), X + 2.0 - 0.5 * ref, scale=np.abs(X) + 2 + np.abs(ref))


@pytest.mark.parametrize("fmt", FMTS)
def test_non_square_and_index_types(ctx, fmt):
    n, m = 1024, 2048
    row, col, val = oracle.random_matrix(n, m, 16, seed=3)
    X = oracle.uniform_real(4, m)
    ref = oracle.csr_spmv(row, col, val, X)
    for rt, ct in ((np.uint64, np.uint64), (np.uint32, np.int32), (np.int32, np.int64)):
        A = vx.SpMat(ctx, n, m, row.astype(rt), col.astype(ct), val, fmt)
        x, y = vx.vector(ctx, X), vx.vector(ctx, n)
        y.assign(A * x)
        assert close(y.read(), ref)
        assert A.rows() == n and A.cols() == m and A.nonzeros() == row[-1]


@pytest.mark.parametrize("fmt", FMTS)
def test_empty_rows_and_empty_matrix(ctx, fmt):
    n, ne = 1024, 256
    row, col, val = oracle.random_matrix(ne, n, 16, seed=5)
    row = np.concatenate([row, np.full(n - ne, row[-1])])
    X = oracle.uniform_real(6, n)
    A = vx.SpMat(ctx, n, n, row, col, val, fmt)
    x, y = vx.vector(ctx, X), vx.vector(ctx, n)
    y.assign(7.0)
    y.assign(A * x)
    assert close(y.read(), oracle.csr_spmv(row, col, val, X))
    assert np.all(y.read()[ne:] == 0)
    # a matrix without any entry must still zero y on `=` (csr.inl:195-200) and leave it on `+=`
    Z = vx.SpMat(ctx, n, n, np.zeros(n + 1, np.int64), np.empty(0, np.int64), np.empty(0), fmt)
    y.assign(3.0)
    y += Z * x
    assert np.all(y.read() == 3.0)
    y.assign(Z * x)
    assert np.all(y.read() == 0.0)


@pytest.mark.parametrize("fmt", FMTS)
def test_tridiagonal_all_rows(ctx, fmt):
    """tests/sparse_matrices.cpp:155-192."""
    n = 1024
    row, col, val = oracle.tridiagonal(n)
    X = oracle.uniform_real(8, n)
    A = vx.SpMat(ctx, n, n, row, col, val, fmt)
    x, y = vx.vector(ctx, X), vx.vector(ctx, n)
    y.assign(A * x)
    got = y.read()
    ref = oracle.csr_spmv(row, col, val, X)
    assert close(got, ref, scale=oracle.csr_absrow(row, col, val, X))
    if ctx.nparts == 1 and fmt == L.FMT_CSR:
        assert np.array_equal(got, ref)                                  # same order, no contraction


@pytest.mark.parametrize("dim,n", [(2, 96), (3, 24)])
@pytest.mark.parametrize("fmt", FMTS)
def test_poisson_kat_and_random_x(ctx, fmt, dim, n):
    """examples/benchmark.cpp:357-473: Poisson with x == 1e-2; interior rows cancel to ~0,
    boundary rows give exactly x.  Then the parity run with U[0,1) x."""
    row, col, val = oracle.poisson(dim, n)
    N = row.size - 1
    A = vx.SpMat(ctx, N, N, row, col, val, fmt)
    for X in (np.full(N, 1e-2), oracle.uniform_real(7, N)):
        x, y = vx.vector(ctx, X), vx.vector(ctx, N)
        y.assign(A * x)
        got = y.read()
        ref = oracle.csr_spmv(row, col, val, X)
        mag = oracle.csr_absrow(row, col, val, X)
        assert close(got, ref, scale=mag)
        bnd = np.diff(row) == 1
        assert np.array_equal(got[bnd], X[bnd])
        y += A * x                                                       # benchmark form y += A*x
        assert close(y.read(), 2 * ref, scale=2 * mag)
        # res = sum((y - y_cpu)^2) as the reference benchmark reports (benchmark.cpp:467-473)
        assert np.sum((got - ref) ** 2) <= 1e-20 * np.sum(mag ** 2) + 1e-300


def test_long_rows_and_wide_matrices(ctx1):
    """Rows longer than a tile (CTA-wide path), rows that switch the tile to warp-per-row mode."""
    rng = np.random.default_rng(17)
    n, m = 300, 20000
    widths = np.concatenate([[5000, 0, 3, 2500], rng.integers(0, 200, n - 4)])
    row = np.concatenate([[0], np.cumsum(widths)]).astype(np.int64)
    col = np.concatenate([np.sort(rng.choice(m, w, replace=False)) for w in widths]).astype(np.int64)
    val = rng.random(col.size)
    X = oracle.uniform_real(19, m)
    ref = oracle.csr_spmv(row, col, val, X)
    for fmt in FMTS:
        A = vx.SpMat(ctx1, n, m, row, col, val, fmt)
        x, y = vx.vector(ctx1, X), vx.vector(ctx1, n)
        y.assign(A * x)
        assert close(y.read(), ref)


def test_float32_values(ctx2):
    n = 2048
    row, col, val = oracle.random_matrix(n, n, 12, seed=23)
    val32 = val.astype(np.float32)
    X = oracle.uniform_real(24, n).astype(np.float32)
    ref = oracle.csr_spmv(row, col, val32.astype(np.float64), X.astype(np.float64))
    for fmt in FMTS:
        A = vx.SpMat(ctx2, n, n, row, col, val32, fmt)
        x, y = vx.vector(ctx2, X), vx.vector(ctx2, n, np.float32)
        y.assign(A * x)
        assert np.allclose(y.read(), ref, rtol=2e-6, atol=1e-6)


def test_hell_layout_matches_reference_packing(ctx1):
    """hybrid_ell.inl:66-193: width heuristic, column-major ELL with pitch alignup(n,16), sentinel -1, CSR tail."""
    n = 1000
    row, col, val = oracle.random_matrix(n, n, 16, seed=31)
    want = oracle.hell_pack(row, col, val)
    h = C.c_void_p()
    lib = L.lib()
    L.check(lib.vexb_csr_create(0, ctx1.streams[0], n, n, row.ctypes.data, 8, col.ctypes.data, 8, val.ctypes.data,
                                L.F64, L.FMT_HELL, C.byref(h)))
    info = L.SpmatInfo()
    L.check(lib.vexb_spmat_get_info(h, C.byref(info)))
    assert (info.ell_width, info.ell_pitch, info.csr_tail_nnz) == (want["width"], want["pitch"], want["csr_col"].size)
    ec = np.empty(info.ell_pitch * info.ell_width, np.int32)
    ev = np.empty(info.ell_pitch * info.ell_width)
    tp = np.empty(n + 1, np.int64); tc = np.empty(info.csr_tail_nnz, np.int32); tv = np.empty(info.csr_tail_nnz)
    L.check(lib.vexb_spmat_hell_download(h, ec.ctypes.data, ev.ctypes.data, tp.ctypes.data, tc.ctypes.data, tv.ctypes.data))
    assert np.array_equal(ec, want["ell_col"].astype(np.int32)) and np.array_equal(ev, want["ell_val"])
    assert np.array_equal(tp, want["csr_row"]) and np.array_equal(tc, want["csr_col"]) and np.array_equal(tv, want["csr_val"])
    X = oracle.uniform_real(32, n)
    assert close(oracle.hell_spmv(want, X), oracle.csr_spmv(row, col, val, X))
    L.check(lib.vexb_spmat_destroy(h))


def test_split_tables_match_reference(ctx3):
    """csr.inl:70-112 split + spmat.hpp:291-378 exchange tables, bit-exact, on 3 slots."""
    n = 3000
    row, col, val = oracle.random_matrix(n, n, 10, seed=41)
    A = vx.SpMat(ctx3, n, n, row, col, val, L.FMT_CSR)
    part = oracle.partition(n, 3)
    assert np.array_equal(A.part, part)
    ex = oracle.setup_exchange(part, part, row, col)
    lib = L.lib()
    for d in range(3):
        info = A.info(d)
        lr, lc, lv, rr, rc, rv = oracle.split_strip(row, col, val, part[d], part[d + 1], part[d], part[d + 1], ex["ghost"][d])
        assert (info.n_ghost, info.loc_nnz, info.rem_nnz) == (ex["ghost"][d].size, lc.size, rc.size)
        glr = np.empty(lr.size, np.int64); glc = np.empty(lc.size, np.int64); glv = np.empty(lv.size)
        grr = np.empty(rr.size, np.int64); grc = np.empty(rc.size, np.int64); grv = np.empty(rv.size)
        L.check(lib.vexb_dspmat_download_split(A.parts[d], glr.ctypes.data, glc.ctypes.data, glv.ctypes.data,
                                               grr.ctypes.data, grc.ctypes.data, grv.ctypes.data))
        for g, w in ((glr, lr), (glc, lc), (glv, lv), (grr, rr), (grc, rc), (grv, rv)):
            assert np.array_equal(g, w)
    X = oracle.uniform_real(43, n)
    x, y = vx.vector(ctx3, X), vx.vector(ctx3, n)
    y.assign(A * x)
    assert close(y.read(), oracle.spmat_apply(part, part, row, col, val, X))


@pytest.mark.parametrize("fmt", [L.FMT_HELL, L.FMT_CSR])
def test_product_inlined_into_the_consumer_kernel(ctx1, fmt):
    """`y = x + A*x`, `y = z - 0.5*(A*x) + 2*(B*w)` and make_inline(A*x) inside any expression run as ONE generated kernel
    (VEXB_TERM_SPMV: sparse/product.hpp:45-130, spmat/inline_spmv.hpp:68-76); same bits as the unfused composition."""
    for row, col, val in (oracle.poisson(2, 70), oracle.random_matrix(3000, 3000, 11, seed=9), oracle.tridiagonal(2000)):
        n = row.size - 1
        X, Z = oracle.uniform_real(3, n), oracle.uniform_real(4, n)
        A = vx.SpMat(ctx1, n, n, row, col, val, fmt)
        x, z, y, y2 = vx.vector(ctx1, X), vx.vector(ctx1, Z), vx.vector(ctx1, n), vx.vector(ctx1, n)
        ax = oracle.csr_spmv(row, col, val, X)
        n0 = vx.launch_count()
        y.assign(x + A * x)
        assert vx.launch_count() - n0 == 1                                   # one kernel: no temporary, no second pass over y
        assert np.array_equal(y.read(), X + ax)
        ctx1.fuse_products = False
        try:
            y2.assign(x + A * x)                                             # the unfused path: vector part, then y += A*x
        finally:
            ctx1.fuse_products = True
        assert np.array_equal(y.read(), y2.read())
        n0 = vx.launch_count()
        y.assign(z - 0.5 * (A * x) + 2.0 * (A * z))
        assert vx.launch_count() - n0 == 1
        az = oracle.csr_spmv(row, col, val, Z)
        assert np.array_equal(y.read(), (Z + (-0.5) * ax) + 2.0 * az)
        y += x + A * x
        mag = oracle.csr_absrow(row, col, val, X) + np.abs(X) + np.abs(Z) + 2 * oracle.csr_absrow(row, col, val, Z)
        assert np.all(np.abs(y.read() - (((Z - 0.5 * ax) + 2.0 * az) + (X + ax))) <= 1e-10 * mag)
        # as a terminal of any expression
        n0 = vx.launch_count()
        y.assign(vx.sin(vx.make_inline(A * x)) * z + 1.0)
        assert vx.launch_count() - n0 == 1
        assert np.allclose(y.read(), np.sin(ax) * Z + 1.0, rtol=1e-14, atol=1e-14)
        # reductions evaluate the inlined expression into a temporary first
        s = vx.Reductor(ctx1, np.float64, L.SUM)(z - vx.make_inline(A * x))
        assert abs(s - oracle.kahan_sum(Z - ax)) <= 1e-10 * np.sum(np.abs(Z) + np.abs(ax))


def test_inline_product_falls_back_to_a_temporary_with_a_halo(ctx2):
    row, col, val = oracle.poisson(2, 60)
    n = row.size - 1
    X = oracle.uniform_real(5, n)
    A = vx.SpMat(ctx2, n, n, row, col, val)
    x, y = vx.vector(ctx2, X), vx.vector(ctx2, n)
    y.assign(x + A * x)
    want = oracle.csr_spmv(row, col, val, X)
    assert np.all(np.abs(y.read() - (X + want)) <= 1e-10 * (np.abs(X) + oracle.csr_absrow(row, col, val, X)))
    y.assign(2.0 * vx.make_inline(A * x) - x)
    assert np.all(np.abs(y.read() - (2.0 * want - X)) <= 1e-10 * (np.abs(X) + 2 * oracle.csr_absrow(row, col, val, X)))
