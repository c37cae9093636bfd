"""CG step (BASELINE configs[4]: SpMV + 2 axpy + 2 dot) built from the three hot paths; checked against the
oracle's composition on a small 3-D Poisson problem.  Covers device-resident scalars and CUDA-graph replay."""
import numpy as np
import pytest

import oracle
import vexcl_b200 as vx
from vexcl_b200 import _lib as L
from vexcl_b200.api import DeviceScalar
from vexcl_b200.solvers import CGDevice, cg_host_scalars

pytestmark = pytest.mark.gpu


def problem(n=20):
    """Symmetric positive definite 7-point Laplacian on an n^3 grid (Dirichlet neighbours simply dropped), so
    that CG converges; the benchmark generator's identity boundary rows make its matrix non-symmetric."""
    idx = np.arange(n ** 3).reshape(n, n, n)
    rows, cols, vals = [idx.ravel()], [idx.ravel()], [np.full(n ** 3, 6.0)]
    for ax in range(3):
        for sh in (-1, 1):
            src = [slice(None)] * 3; dst = [slice(None)] * 3
            src[ax] = slice(1, None) if sh < 0 else slice(None, -1)
            dst[ax] = slice(None, -1) if sh < 0 else slice(1, None)
            rows.append(idx[tuple(src)].ravel()); cols.append(idx[tuple(dst)].ravel()); vals.append(np.full(rows[-1].size, -1.0))
    r, c, v = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    order = np.lexsort((c, r))
    r, c, v = r[order], c[order], v[order]
    N = n ** 3
    row = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=N))]).astype(np.int64)
    b = oracle.uniform_real(3, N)
    return row, c.astype(np.int64), v, b, N


def test_device_scalars_in_expressions(ctx1):
    n = 10001
    X = oracle.uniform_real(1, n)
    x, y = vx.vector(ctx1, X), vx.vector(ctx1, n)
    s, t = DeviceScalar(ctx1, value=2.5), DeviceScalar(ctx1)
    vx.Reductor(ctx1, np.float64, L.SUM).device(x * x, t)
    assert abs(t.get() - np.dot(X, X)) <= 1e-10 * np.dot(X, X)
    z = vx.vector(ctx1, X[::-1].copy())
    y.assign(s * x + z)                                   # sweep kernel with a device-resident coefficient
    assert y.eval_path(L.SET, s * x + z) == "sweep:axpy"
    assert np.array_equal(y.read(), 2.5 * X + X[::-1])
    t.assign(s / 4.0 + 1.0)
    assert t.get() == 2.5 / 4.0 + 1.0
    y.assign(vx.sin(x) * t)                               # interpreter path
    assert np.allclose(y.read(), np.sin(X) * t.get(), rtol=1e-15)


@pytest.mark.parametrize("fmt", [L.FMT_CSR, L.FMT_HELL])
def test_cg_matches_oracle(ctx, fmt):
    row, col, val, b, N = problem()
    iters = 25
    xo, hist_o = oracle.cg(row, col, val, b, np.zeros(N), iters)
    A = vx.SpMat(ctx, N, N, row, col, val, fmt)
    bv, x = vx.vector(ctx, b), vx.vector(ctx, N)
    x.assign(0.0)
    hist = cg_host_scalars(A, bv, x, iters)
    assert hist[-1] < 1e-3 * hist[0]
    assert np.allclose(hist, hist_o, rtol=1e-8)
    assert np.allclose(x.read(), xo, rtol=1e-8, atol=1e-12)


def test_cg_device_scalars_and_graph(ctx1):
    row, col, val, b, N = problem()
    iters = 25
    xo, hist_o = oracle.cg(row, col, val, b, np.zeros(N), iters)
    A = vx.SpMat(ctx1, N, N, row, col, val)
    for use_graph in (False, True):
        bv, x = vx.vector(ctx1, b), vx.vector(ctx1, N)
        x.assign(0.0)
        cg = CGDevice(A, bv, x)
        if use_graph:
            cg.capture()                                   # performs iteration 1 while warming up
            cg.run(iters - 1)
        else:
            cg.run(iters)
        ctx1.finish()
        assert abs(cg.residual2() - hist_o[-1]) <= 1e-8 * hist_o[-1]
        assert np.allclose(x.read(), xo, rtol=1e-8, atol=1e-12)


def test_product_with_fused_dot(ctx1):
    ctx = ctx1
    """SpMat.apply_dot: y = alpha*A*x (+ y) and dot(w, y) in one launch when the strip is hybrid ELL on a single part
    (vexb_dspmat_apply_dot); the composition otherwise (several slots on one device, CSR).  Same y bits as apply."""
    row, col, val, b, N = problem(16)
    X, W, Y0 = oracle.uniform_real(4, N), oracle.uniform_real(5, N), oracle.uniform_real(6, N)
    for fmt in (L.FMT_HELL, L.FMT_CSR):
        A = vx.SpMat(ctx, N, N, row, col, val, fmt)
        x, w, y, y2 = vx.vector(ctx, X), vx.vector(ctx, W), vx.vector(ctx, Y0), vx.vector(ctx, Y0)
        d = DeviceScalar(ctx)
        for dot_with, alpha, append in ((None, 1.0, False), (w, -0.5, True)):
            fused = A.apply_dot(x, y, d, dot_with=dot_with, alpha=alpha, append=append)
            assert fused == (fmt == L.FMT_HELL and ctx.nparts == 1)
            A.apply(x, y2, alpha, append)
            got = y.read()
            assert np.array_equal(got, y2.read())
            ref = float(np.dot(X if dot_with is None else W, got))
            assert abs(d.get() - ref) <= 1e-10 * np.sum(np.abs((X if dot_with is None else W) * got))


def test_fused_cg_matches_oracle(ctx1):
    ctx = ctx1
    """CGFused: product with the dot partials in its epilogue (+ a one-block fold), r sweep + (r, r), x/p sweep -- four
    launches per iteration and GPU; same history as the oracle's composition, stream-launched and replayed as two
    alternating CUDA graphs."""
    from vexcl_b200.solvers import CGFused
    row, col, val, b, N = problem()
    iters = 25
    xo, hist_o = oracle.cg(row, col, val, b, np.zeros(N), iters)
    A = vx.SpMat(ctx, N, N, row, col, val)
    for use_graph in (False, True):
        bv, x = vx.vector(ctx, b), vx.vector(ctx, N)
        x.assign(0.0)
        cg = CGFused(A, bv, x)
        hist = []
        if use_graph:
            cg.capture()                                   # performs iterations 1 and 2 while warming up
            done = 2
        else:
            done = 0
        for _ in range(iters - done):
            cg.run(1)
            hist.append(cg.residual2())
        ctx.finish()
        assert np.allclose(hist, hist_o[done:], rtol=1e-8)
        assert np.allclose(x.read(), xo, rtol=1e-8, atol=1e-12)
        if ctx.nparts == 1:
            assert cg.fused_product
        n0 = vx.launch_count()
        cg.step()
        if ctx.nparts == 1:
            assert vx.launch_count() - n0 == 4             # product(+dot partials), fold, r sweep, x/p sweep
