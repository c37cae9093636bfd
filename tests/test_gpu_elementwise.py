"""Parity of the elementwise path (vexb_eval) with the oracle, through the C ABI.

Mirrors tests/vector_arithmetics.cpp of the reference: assign_expression :33-48,
compound_assignment :50-64, builtin_functions :101, ternary_operator :238,
combine_expressions :271.  Bit-exact bar: the B200 path never contracts mul+add, so it must
equal numpy / the unfused oracle to the last bit; transcendental functions within 2 ulp.
"""
import numpy as np
import pytest

import oracle
import vexcl_b200 as vx
from vexcl_b200 import _lib as L

pytestmark = pytest.mark.gpu

N = 1000 * 1000 + 3          # config 1 size (N=1e6) plus a ragged tail


def rnd(seed, n=N, dtype=np.float64):
    return oracle.uniform_real(seed, n).astype(dtype)


def test_config1_a_eq_b_plus_c_times_d(ctx):
    """BASELINE config 1: a = b + c*d on vex::vector<double>, N=1e6, seed 42 family."""
    B, Cc, D = rnd(42), rnd(43), rnd(44)
    a, b, c, d = vx.vector(ctx, N), vx.vector(ctx, B), vx.vector(ctx, Cc), vx.vector(ctx, D)
    assert a.eval_path(L.SET, b + c * d) == "sweep:muladd"
    a.assign(b + c * d)
    ref = oracle.vec_muladd(np.zeros(N), B, Cc, D)
    assert np.array_equal(a.read(), ref)
    assert np.array_equal(a.read(), B + Cc * D)
    # benchmark form a += b + c*d (examples/benchmark.cpp:171-176)
    a += b + c * d
    assert np.array_equal(a.read(), oracle.vec_muladd(ref, B, Cc, D, accumulate=True))
    # commuted spelling takes the same kernel and gives the same bits
    a.assign(c * d + b)
    assert np.array_equal(a.read(), ref)


def test_interpreter_matches_sweep(ctx):
    B, Cc, D = rnd(1), rnd(2), rnd(3)
    a, b, c, d = vx.vector(ctx, N), vx.vector(ctx, B), vx.vector(ctx, Cc), vx.vector(ctx, D)
    a.assign(b + c * d)
    fast = a.read()
    vx.set_param("eval.force_interp", 1)
    try:
        assert a.eval_path(L.SET, b + c * d) == "interp"
        a.assign(b + c * d)
        slow = a.read()
        a += b + c * d
        slow2 = a.read()
    finally:
        vx.set_param("eval.force_interp", 0)
    assert np.array_equal(fast, slow)
    assert np.array_equal(slow2, fast + (B + Cc * D))


def test_saxpy_and_cg_updates(ctx):
    A0, B = rnd(5), rnd(6)
    alpha = float(oracle.uniform_real(7, 1)[0])
    a, b = vx.vector(ctx, A0), vx.vector(ctx, B)
    assert a.eval_path(L.SET, alpha * a + b) == "sweep:axpy"
    a.assign(alpha * a + b)                                   # examples/benchmark.cpp:102-107
    assert np.array_equal(a.read(), oracle.vec_saxpy(A0, alpha, B))
    r, p = vx.vector(ctx, A0), vx.vector(ctx, B)
    p.assign(r + alpha * p)
    assert np.array_equal(p.read(), A0 + alpha * B)
    r.assign(r - alpha * p)
    assert np.array_equal(r.read(), A0 - alpha * (A0 + alpha * B))
    r -= p
    assert np.array_equal(r.read(), (A0 - alpha * (A0 + alpha * B)) - (A0 + alpha * B))


def test_assign_expression_closed_form(ctx):
    """vector_arithmetics.cpp:41-47: x = 5*sin(y)+z with y=42, z=67 -> 5*sin(42)+67."""
    n = 1024
    x, y, z = vx.vector(ctx, n), vx.vector(ctx, n), vx.vector(ctx, n)
    y.assign(42)
    z.assign(67)
    x.assign(5 * vx.sin(y) + z)
    got = x.read()
    want = 5 * np.sin(42.0) + 67
    assert np.all(np.abs(got - want) <= 1e-14 * abs(want))    # BOOST_CHECK_CLOSE(..., 1e-12 %)


def test_compound_assignment_exact(ctx):
    """vector_arithmetics.cpp:50-64: x = 0; x += 1 -> 1; x -= 2 -> -1 (exact)."""
    n = 1024
    x = vx.vector(ctx, n)
    x.assign(0)
    x += 1
    assert np.all(x.read() == 1)
    x -= 2
    assert np.all(x.read() == -1)
    x *= 3
    assert np.all(x.read() == -3)
    x /= 2
    assert np.all(x.read() == -1.5)


@pytest.mark.parametrize("dtype", [np.float32, np.int32, np.uint32, np.int64, np.uint64])
def test_other_element_types(ctx1, dtype):
    n = 4099
    rng = np.random.default_rng(11)
    if np.issubdtype(dtype, np.floating):
        A, B = rng.random(n).astype(dtype), rng.random(n).astype(dtype)
    else:
        A, B = rng.integers(1, 100, n).astype(dtype), rng.integers(1, 100, n).astype(dtype)
    a, b, o = vx.vector(ctx1, A), vx.vector(ctx1, B), vx.vector(ctx1, n, dtype)
    o.assign(a + b * a)
    assert np.array_equal(o.read(), (A + B * A).astype(dtype))
    o.assign(a * 3 - b)
    assert np.array_equal(o.read(), (A * dtype(3) - B).astype(dtype))
    if not np.issubdtype(dtype, np.floating):
        o.assign((a ^ b) | (a & b))
        assert np.array_equal(o.read(), (A ^ B) | (A & B))
        o.assign(a % b + (a << 2) + (b >> 1))
        assert np.array_equal(o.read(), (A % B + (A << dtype(2)) + (B >> dtype(1))).astype(dtype))
        o.assign(a)
        o %= b
        assert np.array_equal(o.read(), A % B)
        o <<= 2
        assert np.array_equal(o.read(), ((A % B) << dtype(2)).astype(dtype))


def test_mixed_types_follow_c_promotion(ctx1):
    n = 2000
    rng = np.random.default_rng(3)
    I = rng.integers(-50, 50, n).astype(np.int32)
    F = rng.random(n).astype(np.float32)
    Dd = rng.random(n)
    i, f, d, o = vx.vector(ctx1, I), vx.vector(ctx1, F), vx.vector(ctx1, Dd), vx.vector(ctx1, n)
    o.assign(i * f + d)                      # (float)(i)*f in float, then + double
    assert np.array_equal(o.read(), (I.astype(np.float32) * F).astype(np.float64) + Dd)
    oi = vx.vector(ctx1, n, np.int32)
    oi.assign(d * 100)                       # double -> int truncation on store
    assert np.array_equal(oi.read(), (Dd * 100).astype(np.int32))
    oi.assign(i)
    oi += d * 10                             # int += double: computed in double, truncated
    assert np.array_equal(oi.read(), (I + Dd * 10).astype(np.int32))


def test_builtin_functions_and_ternary(ctx1):
    """vector_arithmetics.cpp:101-110 (sin^2+cos^2 ~ 1), :238 (ternary), pow/sqrt/fabs."""
    n = 5000
    X = rnd(9, n) * 4 - 2
    x, o = vx.vector(ctx1, X), vx.vector(ctx1, n)
    o.assign(vx.pow_(vx.sin(x), 2.0) + vx.pow_(vx.cos(x), 2.0))
    assert np.all(np.abs(o.read() - 1) <= 1e-8 / 100)
    o.assign(vx.if_else(x > 0, vx.sqrt(vx.fabs(x)), -x))
    want = np.where(X > 0, np.sqrt(np.abs(X)), -X)
    assert np.array_equal(o.read(), want)
    o.assign(vx.fmax(x, 0.5) + vx.fmin(x, -0.5) + vx.floor(x) + vx.exp(x))
    want = np.maximum(X, 0.5) + np.minimum(X, -0.5) + np.floor(X) + np.exp(X)
    mag = np.abs(np.maximum(X, 0.5)) + np.abs(np.minimum(X, -0.5)) + np.abs(np.floor(X)) + np.exp(X)
    assert np.all(np.abs(o.read() - want) <= 1e-15 * mag)          # exp() is within 1-2 ulp of libm
    o.assign(vx.fma(x, x, x))
    assert np.allclose(o.read(), X * X + X, rtol=1e-15)


def test_element_index_carries_part_start(ctx):
    n = 70001
    o = vx.vector(ctx, n)
    o.assign(vx.ElementIndex(5) * 2.0)
    assert np.array_equal(o.read(), (np.arange(n) + 5) * 2.0)


def test_empty_partitions_are_legal(ctx2):
    """tests/vector_create.cpp:189-194: n=1 over 2 devices (second slice empty)."""
    x = vx.vector(ctx2, 1)
    assert sorted([x.part_size(0), x.part_size(1)]) == [0, 1]
    x.assign(42)
    assert x.read()[0] == 42
    assert x[0] == 42
    s = vx.Reductor(ctx2, np.float64, L.SUM)
    assert s(x) == 42


def test_unaligned_slices_take_the_interpreter_and_agree(ctx1):
    """A raw C-ABI call with pointers offset by 8 bytes: the 256-bit sweep must not be chosen."""
    import ctypes as C
    n = 10007
    B = rnd(21, n + 1)
    b = vx.vector(ctx1, B)
    o = vx.vector(ctx1, n + 1)
    o.assign(0)
    e = L.Expr()
    e.n_terms, e.n_code = 1, 3
    e.term[0].kind, e.term[0].dtype = L.TERM_VEC, L.F64
    e.term[0].v.ptr = b.bufs[0].value + 8
    e.code[0].op, e.code[0].type, e.code[0].arg = L.OP["TERM"], L.F64, 0
    e.code[1].op, e.code[1].type, e.code[1].arg = L.OP["TERM"], L.F64, 0
    e.code[2].op, e.code[2].type = L.OP["MUL"], L.F64
    L.check(L.lib().vexb_eval(0, ctx1.streams[0], C.c_void_p(o.bufs[0].value + 8), L.F64, L.SET, C.byref(e), n, 0))
    got = o.read()
    assert got[0] == 0 and np.array_equal(got[1:], B[1:] * B[1:])


def test_errors_are_reported_not_thrown_across_the_abi(ctx1):
    import ctypes as C
    e = L.Expr()
    e.n_terms, e.n_code = 0, 1
    e.code[0].op, e.code[0].type = L.OP["ADD"], L.F64          # stack underflow
    o = vx.vector(ctx1, 16)
    rc = L.lib().vexb_eval(0, ctx1.streams[0], o.bufs[0], L.F64, L.SET, C.byref(e), 16, 0)
    assert rc == 2 and b"underflow" in L.lib().vexb_last_error()
    a, b = vx.vector(ctx1, 16), vx.vector(ctx1, 17)
    with pytest.raises(ValueError):
        o.assign(a + b)                                         # expression_size_check, vector_arithmetics.cpp:319-327
    of = vx.vector(ctx1, 16)
    with pytest.raises(vx.VexbError):
        of <<= 2                                                # shift on a floating vector
