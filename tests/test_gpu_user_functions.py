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
            vx.set_param("eval.jit", 0)
            vx.set_param("eval.force_interp", 0)
        assert np.array_equal(o1.read(), o2.read()) or np.allclose(o1.read(), o2.read(), rtol=1e-15, atol=1e-15)
    i1, i2 = vx.vector(ctx1, n, np.int32), vx.vector(ctx1, n, np.int32)
    i1.assign((k << 3) ^ (k * 5) | (k & 12))
    vx.set_param("eval.jit", 1)
    try:
        i2.assign((k << 3) ^ (k * 5) | (k & 12))
    finally:
        vx.set_param("eval.jit", 0)
    assert np.array_equal(i1.read(), i2.read())
