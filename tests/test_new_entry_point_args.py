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
