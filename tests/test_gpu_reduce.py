"""Parity of vexb_reduce / vex::Reductor with the oracle (reference: tests/vector_arithmetics.cpp:66-99,
tests/threads.cpp:9-35).  SUM within 1e-10 relative (summation order differs by design),
MIN / MAX / MIN_MAX / integer SUM exact."""
import numpy as np
import pytest

import oracle
import vexcl_b200 as vx
from vexcl_b200 import _lib as L

pytestmark = pytest.mark.gpu

N = (1 << 20) + 13


def test_reduce_expression(ctx):
    """reduce_expression: sum / sum_Kahan vs a Kahan accumulator at 1e-8 % ; min / max exact."""
    X = oracle.uniform_real(77, N)
    x = vx.vector(ctx, X)
    ssum, skah = vx.Reductor(ctx, np.float64, L.SUM), vx.Reductor(ctx, np.float64, L.SUM_KAHAN)
    smin, smax, smm = (vx.Reductor(ctx, np.float64, k) for k in (L.MIN, L.MAX, L.MINMAX))
    ref = oracle.kahan_sum(X)
    assert abs(ssum(x) - ref) <= 1e-10 * abs(ref)
    assert abs(skah(x) - ref) <= 1e-10 * abs(ref)
    assert abs(ssum(x) - oracle.reduce(X, oracle.SUM)) <= 1e-10 * abs(ref)
    assert smin(x) == X.min() == oracle.reduce(X, oracle.MIN)
    assert smax(x) == X.max() == oracle.reduce(X, oracle.MAX)
    assert smm(x) == (X.min(), X.max())
    assert smax(vx.fabs(x - x)) == 0                         # vector_arithmetics.cpp:98
    # expression reductions
    Y = oracle.uniform_real(78, N)
    y = vx.vector(ctx, Y)
    ref = oracle.reduce_dot(X, Y, kahan=True)
    assert abs(ssum(x * y) - ref) <= 1e-10 * abs(ref)        # examples/benchmark.cpp:236-241
    assert abs(ssum(x * x) - oracle.reduce_dot(X, X, kahan=True)) <= 1e-10 * abs(ref)
    ref = oracle.kahan_sum(np.sin(X) * Y + 1.0)
    assert abs(ssum(vx.sin(x) * y + 1.0) - ref) <= 1e-10 * abs(ref)
    assert smax(vx.fabs(x - y)) == np.abs(X - Y).max()


def test_counting_reductions_exact(ctx):
    """vector_arithmetics.cpp:127-128, :143-144 and threads.cpp:34: integer-valued sums are exact."""
    n = 100003
    X = oracle.uniform_real(5, n)
    x = vx.vector(ctx, X)
    isum = vx.Reductor(ctx, np.int64, L.SUM)
    assert isum(x > 2.0) == 0
    assert isum(x >= 0.0) == n
    assert isum(x > 0.5) == int((X > 0.5).sum())
    ones = vx.vector(ctx, n, np.int32)
    ones.assign(1)
    assert vx.Reductor(ctx, np.int32, L.SUM)(ones * 2) == 2 * n
    usum = vx.Reductor(ctx, np.uint64, L.SUM)
    assert usum(vx.ElementIndex() * ones) == n * (n - 1) // 2


def test_float32_and_forced_interpreter(ctx1):
    n = 300001
    X = oracle.uniform_real(9, n).astype(np.float32)
    x = vx.vector(ctx1, X)
    s = vx.Reductor(ctx1, np.float32, L.SUM)
    ref = float(np.sum(X.astype(np.float64)))
    assert abs(float(s(x)) - ref) <= 1e-5 * ref
    assert vx.Reductor(ctx1, np.float32, L.MAX)(x) == X.max()
    Xd = oracle.uniform_real(10, n)
    xd = vx.vector(ctx1, Xd)
    sd = vx.Reductor(ctx1, np.float64, L.SUM)
    fast = sd(xd * xd)
    vx.set_param("eval.force_interp", 1)
    try:
        slow = sd(xd * xd)
    finally:
        vx.set_param("eval.force_interp", 0)
    assert abs(fast - slow) <= 1e-12 * abs(fast)


def test_reduce_is_deterministic(ctx1):
    X = oracle.uniform_real(123, N)
    x = vx.vector(ctx1, X)
    s = vx.Reductor(ctx1, np.float64, L.SUM)
    vals = {float(s(x)) for _ in range(5)}
    assert len(vals) == 1


def test_empty_expression_returns_initial(ctx1):
    """reductor.hpp:318-321."""
    x = vx.vector(ctx1, 0)
    assert vx.Reductor(ctx1, np.float64, L.SUM)(x) == 0
    assert vx.Reductor(ctx1, np.float64, L.MAX)(x) == np.finfo(np.float64).min
    assert vx.Reductor(ctx1, np.float64, L.MIN)(x) == np.finfo(np.float64).max


def test_combined_reductors(ctx):
    """vex::CombineReductors<R...> (reductor.hpp:132-280): several reductions of one expression, one pass over memory."""
    n = 250_007
    X = (oracle.uniform_real(21, n) - 0.25) * 1e3
    Y = oracle.uniform_real(22, n)
    x, y = vx.vector(ctx, X), vx.vector(ctx, Y)
    lo, hi, s = vx.Reductor(ctx, np.float64, [L.MIN, L.MAX, L.SUM])(x)
    assert lo == X.min() and hi == X.max()
    ref = oracle.kahan_sum(X)
    assert abs(s - ref) <= 1e-10 * np.sum(np.abs(X))
    r = vx.Reductor(ctx, np.float64, [L.SUM, L.SUM_KAHAN, L.MAX, L.MIN, L.MAX])(x * y + 1.0)
    E = X * Y + 1.0
    assert abs(r[0] - oracle.kahan_sum(E)) <= 1e-10 * np.sum(np.abs(E)) and abs(r[1] - oracle.kahan_sum(E)) <= 1e-10 * np.sum(np.abs(E))
    assert r[2] == E.max() and r[3] == E.min() and r[4] == E.max()
    cnt, any_ = vx.Reductor(ctx, np.int64, [L.SUM, L.MAX])(x > 0.0)
    assert cnt == int((X > 0).sum()) and any_ == 1
    for _ in range(3):                                           # the workspace slices are left reusable
        assert vx.Reductor(ctx, np.float64, [L.MAX, L.MIN])(x) == (X.max(), X.min())
