"""User-defined functions (VEX_FUNCTION) and the NVRTC side path on the GPU.
Reference cases: tests/vector_arithmetics.cpp:113-145 (user_defined_functions: greater / times2, counting sums)."""
import numpy as np
import pytest

import oracle
import vexcl_b200 as vx
from vexcl_b200 import _lib as L
from vexcl_b200.api import UserFunction

pytestmark = pytest.mark.gpu


def test_user_defined_functions(ctx):
    N = 1024
    x, y = vx.vector(ctx, N), vx.vector(ctx, N)
    x.assign(1)
    y.assign(2)
    greater = UserFunction(np.uint64, "greater", [(np.float64, "x"), (np.float64, "y")], "return x > y;")
    times2 = UserFunction(np.float64, "times2", [(np.float64, "x")], "return x * 2;")
    count = vx.Reductor(ctx, np.uint64, L.SUM)
    assert count(greater(x, y)) == 0                       # vector_arithmetics.cpp:127
    assert count(greater(y, x)) == N                       # :128
    ssum = vx.Reductor(ctx, np.float64, L.SUM)
    assert ssum(times2(x)) == 2 * N                        # :143-144
    z = vx.vector(ctx, N)
    z.assign(times2(x) + y * greater(y, x))
    assert np.all(z.read() == 2 + 2)
    z += times2(z)
    assert np.all(z.read() == 12)


def test_jit_matches_interpreter_bit_for_bit(ctx1):
    """eval.jit = 1 sends interpreter-class expressions through NVRTC; same IR semantics, --fmad=false: same bits."""
    n = 100003
    B, Cc, D = (oracle.uniform_real(s, n) for s in (1, 2, 3))
    b, c, d = vx.vector(ctx1, B), vx.vector(ctx1, Cc), vx.vector(ctx1, D)
    k = vx.vector(ctx1, (np.arange(n) % 97).astype(np.int32))
    o1, o2 = vx.vector(ctx1, n), vx.vector(ctx1, n)
    exprs = [
        lambda: vx.sin(b) * c + vx.sqrt(d) / (b + 1.5),
        lambda: vx.if_else(b > c, b * c - d, vx.fmax(c, d) + k),
        lambda: (b + c * d) * (k % 7) - vx.ElementIndex(5) * 1e-3,
        lambda: vx.pow_(b + 1.0, c) + vx.fma(b, c, d) + vx.floor(d * 10),
    ]
    for mk in exprs:
        vx.set_param("eval.force_interp", 1)
        try:
            o1.assign(mk())
            vx.set_param("eval.jit", 1)
            assert o2.eval_path(L.SET, mk()) == "jit"
            o2.assign(mk())
            o2 += mk()
            o2 -= mk()
        finally:
            vx.set_param("eval.jit", 2)
            vx.set_param("eval.force_interp", 0)
        assert np.array_equal(o1.read(), o2.read()) or np.allclose(o1.read(), o2.read(), rtol=1e-15, atol=1e-15)
    i1, i2 = vx.vector(ctx1, n, np.int32), vx.vector(ctx1, n, np.int32)
    i1.assign((k << 3) ^ (k * 5) | (k & 12))
    vx.set_param("eval.jit", 1)
    try:
        i2.assign((k << 3) ^ (k * 5) | (k & 12))
    finally:
        vx.set_param("eval.jit", 2)
    assert np.array_equal(i1.read(), i2.read())


def _wait_for_jit(timeout=60.0):
    import ctypes as C
    import time
    t0, p = time.time(), C.c_int(1)
    while time.time() - t0 < timeout:
        L.check(L.lib().vexb_jit_pending(C.byref(p)))
        if not p.value:
            return True
        time.sleep(0.02)
    return False


def test_background_compilation_takes_over_from_the_interpreter(ctx1):
    """Default mode (eval.jit = 2): the first use of a new expression shape starts NVRTC on a background thread and is
    served by the interpreter; once the kernel is ready, later launches take it.  Same bits either way."""
    n = 200_003
    B, Cc, D = (oracle.uniform_real(s, n) for s in (11, 12, 13))
    b, c, d = vx.vector(ctx1, B), vx.vector(ctx1, Cc), vx.vector(ctx1, D)
    o1, o2 = vx.vector(ctx1, n), vx.vector(ctx1, n)
    mk = lambda: (b - c) * (b + c) / (d + 0.75) + vx.cos(b) * 0.125 + d * c * b     # a shape no other test uses
    n0 = vx.launch_count()
    o1.assign(mk())                                            # interpreter (compilation started)
    assert vx.launch_count() - n0 == 1
    assert _wait_for_jit()
    o2.assign(mk())                                            # specialised kernel
    ctx1.finish()
    assert np.array_equal(o1.read(), o2.read())
    o2 += mk()                                                 # a different request shape (compound): interpreter first ...
    assert _wait_for_jit()
    o1 += mk()                                                 # ... then its own kernel
    assert np.array_equal(o1.read(), o2.read())


def test_compound_shift_uses_the_left_operand_type(ctx1):
    """`a >>= b` with signed a and unsigned b is an arithmetic shift of a: interpreter and NVRTC agree with numpy."""
    n = 4099
    A = (np.arange(n, dtype=np.int64) * 7919 % 200001 - 100000).astype(np.int32)
    Bc = (np.arange(n) % 5).astype(np.uint32)
    for jit in (0, 1):
        vx.set_param("eval.jit", jit)
        try:
            a, b = vx.vector(ctx1, A), vx.vector(ctx1, Bc)
            a >>= b
            assert np.array_equal(a.read(), A >> Bc.astype(np.int32))
            a = vx.vector(ctx1, A)
            a <<= b
            assert np.array_equal(a.read(), (A.astype(np.int64) << Bc).astype(np.int32))
            u = vx.vector(ctx1, A.astype(np.uint32))
            u >>= vx.vector(ctx1, Bc.astype(np.int32))
            assert np.array_equal(u.read(), A.astype(np.uint32) >> Bc)
        finally:
            vx.set_param("eval.jit", 2)


@pytest.mark.parametrize("nparts", [1, 2])
def test_multi_expression_assignment_in_one_kernel(ctx1, ctx2, nparts):
    """vex::tie(x, y) = std::tie(x + y, y - x) and friends through vexb_eval_multi: one generated kernel for all
    components, right-hand sides evaluated before any target is written (assign_multiexpression,
    vexcl/operations.hpp:2081-2185, tests/multivector_arithmetics.cpp).  Same bits as component by component."""
    import ctypes as C
    import time
    ctx = {1: ctx1, 2: ctx2}[nparts]
    n = 100_003
    X, Y, Z = (oracle.uniform_real(s, n) for s in (41, 42, 43))
    vx.set_param("eval.jit", 1)                               # compile at once: the fused kernel serves the first call
    try:
        x, y, z = vx.vector(ctx, X), vx.vector(ctx, Y), vx.vector(ctx, Z)
        assert vx.assign_multi([x, y], [x + y, y - x]) is True          # each component reads what the other writes
        assert np.array_equal(x.read(), X + Y) and np.array_equal(y.read(), Y - X)
        n0 = vx.launch_count()
        assert vx.assign_multi([x, y, z], [vx.sin(x) * 2.0 + z, x * y - z / 3.0, vx.sqrt(vx.fabs(y)) + x], L.ADD) is True
        assert vx.launch_count() - n0 == ctx.nparts                     # one launch per device slice
        x0, y0, z0 = X + Y, Y - X, Z
        want = (x0 + (np.sin(x0) * 2.0 + z0), y0 + (x0 * y0 - z0 / 3.0), z0 + (np.sqrt(np.abs(y0)) + x0))
        for got, w in zip((x, y, z), want):
            assert np.allclose(got.read(), w, rtol=1e-14, atol=0)
    finally:
        vx.set_param("eval.jit", 2)
    # default mode: the first call is served component by component while the kernel compiles, later calls are fused
    a, b = vx.vector(ctx, X), vx.vector(ctx, Y)
    first = vx.assign_multi([a, b], [a * 3.0 - b, b * a + 1.5])
    assert np.array_equal(a.read(), X * 3.0 - Y) and np.array_equal(b.read(), Y * X + 1.5)
    pend, t0 = C.c_int(1), time.time()
    while pend.value and time.time() - t0 < 60:
        L.check(L.lib().vexb_jit_pending(C.byref(pend)))
    a.write(X); b.write(Y)
    assert vx.assign_multi([a, b], [a * 3.0 - b, b * a + 1.5]) is True
    assert np.array_equal(a.read(), X * 3.0 - Y) and np.array_equal(b.read(), Y * X + 1.5)
    assert first in (True, False)
