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
