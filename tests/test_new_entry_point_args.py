"""Argument checks of the CCSR / stencil entry points that happen before any device work, so they can run without a
GPU: a matrix that would read outside x is rejected (the reference reads unchecked, ccsr.hpp:195), bad index values,
bad stencil geometry.  With a GPU present the same calls are covered by tests/test_gpu_ccsr.py / test_gpu_stencil.py."""
import ctypes as C

import numpy as np
import pytest

from vexcl_b200 import _lib as L, gen


def _create(idx, row, col, val, n=None):
    h = C.c_void_p()
    idx, row = np.ascontiguousarray(idx, np.uint64), np.ascontiguousarray(row, np.uint64)
    col, val = np.ascontiguousarray(col, np.int32), np.ascontiguousarray(val, np.float64)
    rc = L.lib().vexb_ccsr_create(0, None, idx.size if n is None else n, row.size - 1, idx.ctypes.data, 8, row.ctypes.data, 8,
                                  col.ctypes.data, 4, val.ctypes.data, L.F64, C.byref(h))
    return rc, L.lib().vexb_last_error().decode(), h


def test_ccsr_create_validates_before_touching_the_device(built):
    idx, row, col, val = gen.poisson_ccsr(8)
    bad = idx.copy()
    bad[0] = 1
    rc, msg, _ = _create(bad, row, col, val)
    assert rc == 2 and "reaches outside the vector" in msg            # VEXB_ERR_INVALID
    bad[0] = 7
    rc, msg, _ = _create(bad, row, col, val)
    assert rc == 2 and "names no unique row" in msg
    rc, msg, _ = _create(idx, [0, 2, 1], col[:2], val[:2])
    assert rc == 2 and "non-decreasing" in msg
    rc, msg, _ = _create(idx, [1, 1, 8], col, val)
    assert rc == 2 and "start at 0" in msg
    lib = L.lib()
    h = C.c_void_p()
    assert lib.vexb_ccsr_create(0, None, 8, 2, idx.ctypes.data, 2, row.ctypes.data, 8, col.ctypes.data, 4, val.ctypes.data, L.F64,
                                C.byref(h)) == 2                      # 16-bit host indices are not accepted
    assert lib.vexb_ccsr_create(0, None, 8, 2, idx.ctypes.data, 8, row.ctypes.data, 8, col.ctypes.data, 4, val.ctypes.data, L.I32,
                                C.byref(h)) == 2                      # integer values are not accepted
    assert lib.vexb_ccsr_destroy(None) == 0


def test_stencil_apply_validates_geometry(built):
    lib = L.lib()
    one = C.c_void_p(0x1000)
    assert lib.vexb_stencil_apply(0, None, L.F64, one, 0, 0, one, 16, None, None, one, 1.0, 0) == 2       # width 0
    assert lib.vexb_stencil_apply(0, None, L.F64, one, 5, 5, one, 16, None, None, one, 1.0, 0) == 2       # center outside
    assert lib.vexb_stencil_apply(0, None, L.I32, one, 5, 2, one, 16, None, None, one, 1.0, 0) == 2       # integer stencil
    assert b"width" in lib.vexb_last_error() or b"float" in lib.vexb_last_error()
    assert lib.vexb_stencil_apply(0, None, L.F64, one, 5, 2, one, 0, None, None, one, 1.0, 0) == 0        # empty slice: nothing to do
    assert lib.vexb_copy_peer(0, one, 0, one, 0, None) == 0                                               # zero bytes: nothing to do


def test_multi_expression_and_precompile_entry_points_validate_arguments(built):
    """vexb_eval_multi / vexb_jit_precompile (round 2): argument checks and the "not handled" answers that need no device."""
    import vexcl_b200 as vx
    from vexcl_b200 import api
    lib = L.lib()

    class FakeCtx:
        nparts, local, is_distributed = 1, [0], False
        def partition(self, n): return vx.partition(n, 1)

    def fake_vec(n, addr):
        v = api.vector.__new__(api.vector)
        v.ctx, v.n, v.np_dtype, v.dtype, v.part, v.bufs = FakeCtx(), n, np.dtype(np.float64), api._vdt(np.float64), vx.partition(n, 1), {0: C.c_void_p(addr)}
        return v

    x, y = fake_vec(64, 0x1000), fake_vec(64, 0x2000)
    lows = []
    for e in (x + y, y - x):
        low = api._Lowering(0, 0); low.size = 64; low.lower(api.wrap(e)); lows.append(low)
    es = (C.POINTER(L.Expr) * 2)(*[C.pointer(l.e) for l in lows])
    out = (C.c_void_p * 2)(x.bufs[0], y.bufs[0])
    handled = C.c_int(7)
    assert lib.vexb_eval_multi(0, None, 2, out, L.F64, L.SET, es, 0, 0, C.byref(handled)) == 0 and handled.value == 1   # empty slice: done
    assert lib.vexb_eval_multi(0, None, 1, out, L.F64, L.SET, es, 64, 0, C.byref(handled)) == 0 and handled.value == 0  # one component: not here
    assert lib.vexb_eval_multi(0, None, 2, out, L.F64, L.SET, es, 64, 0, None) == 2                                       # handled is NULL
    assert lib.vexb_eval_multi(0, None, 2, out, 99, L.SET, es, 64, 0, C.byref(handled)) == 2                              # bad dtype
    assert lib.vexb_eval_multi(0, None, 2, None, L.F64, L.SET, es, 64, 0, C.byref(handled)) == 2                          # no targets
    assert lib.vexb_jit_precompile(L.F64, L.SET, None, 0) == 2
    assert lib.vexb_jit_precompile(99, L.SET, C.byref(lows[0].e), 0) == 2
    assert lib.vexb_jit_precompile(L.F64, L.SET, C.byref(lows[0].e), 0) == 0                                              # NVRTC, no device needed
    # the fused multi-expression kernel compiles for sm_100a without a device: components as functions, reads before writes
    n = C.c_size_t(0)
    assert lib.vexb_jit_source_multi(L.F64, L.ADD, 2, es, None, C.byref(n), 0) == 0 and n.value > 100
    buf = C.create_string_buffer(n.value + 256)
    cap = C.c_size_t(len(buf))
    assert lib.vexb_jit_source_multi(L.F64, L.ADD, 2, es, buf, C.byref(cap), 1) == 0, lib.vexb_last_error()
    src = buf.value.decode()
    assert "NVRTC: ok" in src and "vexb_elem_0(" in src and "vexb_elem_1(" in src and "struct multi_j { terms_j c[2]; double *lhs[2]; };" in src
    assert src.index("const double a1 = vexb_elem_1(") < src.index("mt.lhs[0][i] = a0;")      # every read before the first write
    assert lib.vexb_jit_source_multi(L.F64, L.SET, 1, es, None, C.byref(n), 0) == 2
