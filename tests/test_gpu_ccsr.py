"""vex::SpMatCCSR on the GPU against oracle/ccsr.py -- the reference's ccsr_vector_product test (tests/spmv.cpp:148-231)
and the benchmark matrix (examples/benchmark.cpp:481-606), bit-exact because neither side contracts a*b+c."""
import numpy as np
import pytest

import oracle
from oracle import ccsr
import vexcl_b200 as vx
from vexcl_b200 import gen

pytestmark = pytest.mark.gpu


# "jit" = the matrix-specialised NVRTC kernel (ccsr.jit): the unique rows become code.
VARIANTS = [1, 2, 3, "jit"]


@pytest.fixture(params=VARIANTS, autouse=True)
def kernel_variant(request, built):
    """Every case runs under each kernel variant (csrc/ccsr.cu: ccsr.kernel, ccsr.jit)."""
    if request.param == "jit":
        vx.set_param("ccsr.jit", 1)
    else:
        vx.set_param("ccsr.kernel", request.param)
    yield request.param
    vx.set_param("ccsr.kernel", 0)          # 0 = not set: back to the built-in default
    vx.set_param("ccsr.jit", 0)


def test_ccsr_vector_product(ctx1):
    n = 32
    N = n ** 3
    idx, row, col, val = gen.poisson_ccsr(n)
    A = vx.SpMatCCSR(ctx1, N, idx, row, col, val)
    info = A.info()
    assert (info.nrows, info.unique_rows, info.nnz, info.idx_bytes, info.table_in_smem) == (N, 2, 8, 1, 1)
    xh = oracle.uniform_real(1, N)
    x, y = vx.vector(ctx1, xh), vx.vector(ctx1, N)
    want = ccsr.ccsr_spmv(N, idx, row, col, val, xh)
    y.assign(A * x)
    assert np.array_equal(y.read(), want)
    y.assign(x + A * x)                                      # spmv.cpp:208-218
    assert np.array_equal(y.read(), xh + want)
    y -= 0.5 * (A * x)
    assert np.array_equal(y.read(), (xh + want) + (-0.5) * want)
    y.assign(vx.sin(x) - A * x)
    assert np.allclose(y.read(), np.sin(xh) - want, rtol=1e-13, atol=1e-9)


@pytest.mark.parametrize("val_dtype,col_dtype,idx_dtype", [(np.float64, np.int64, np.uint64), (np.float32, np.int32, np.uint32)])
def test_ccsr_types_and_many_unique_rows(ctx1, val_dtype, col_dtype, idx_dtype):
    rng = np.random.default_rng(5)
    N, m = 100_003, 700                                     # m > 256: 2-byte idx on the device
    widths = rng.integers(0, 9, m)
    widths[0] = 0
    row = np.concatenate([[0], np.cumsum(widths)]).astype(idx_dtype)
    col = rng.integers(-40, 41, int(row[-1])).astype(col_dtype)
    val = rng.random(int(row[-1])).astype(val_dtype)
    idx = rng.integers(0, m, N).astype(idx_dtype)
    idx[:40] = 0
    idx[-40:] = 0
    xh = rng.random(N).astype(val_dtype)
    A = vx.SpMatCCSR(ctx1, N, idx, row, col, val)
    assert A.info().idx_bytes == 2
    x, y = vx.vector(ctx1, xh), vx.vector(ctx1, N, dtype=val_dtype)
    y.assign(A * x)
    assert np.array_equal(y.read(), ccsr.ccsr_spmv(N, idx, row, col, val, xh))


def test_ccsr_table_too_large_for_shared_memory(ctx1):
    rng = np.random.default_rng(6)
    N, m = 20_000, 70_000                                   # 4-byte idx, table read from global memory
    row = np.arange(m + 1, dtype=np.uint64)
    col = np.zeros(m, dtype=np.int32)
    val = rng.random(m)
    idx = rng.integers(0, m, N).astype(np.uint64)
    xh = rng.random(N)
    A = vx.SpMatCCSR(ctx1, N, idx, row, col, val)
    assert (A.info().idx_bytes, A.info().table_in_smem) == (4, 0)
    x, y = vx.vector(ctx1, xh), vx.vector(ctx1, N)
    y.assign(A * x)
    assert np.array_equal(y.read(), val[idx.astype(np.int64)] * xh)


def test_ccsr_rejects_what_the_reference_would_read_out_of_bounds(ctx1):
    idx, row, col, val = gen.poisson_ccsr(8)
    bad = idx.copy()
    bad[0] = 1                                              # the stencil row at the first grid point reaches x[-64]
    with pytest.raises(vx.VexbError, match="reaches outside"):
        vx.SpMatCCSR(ctx1, 512, bad, row, col, val)
    bad[0] = 5
    with pytest.raises(vx.VexbError, match="names no unique row"):
        vx.SpMatCCSR(ctx1, 512, bad, row, col, val)
    with pytest.raises(TypeError):
        vx.SpMatCCSR(ctx1, 512, idx, row, col.astype(np.uint32), val)


def test_ccsr_benchmark_size_matches_hell(ctx1):
    """benchmark.cpp:481-606 at n=128: the CCSR product equals the SpMat product of the same matrix, bit for bit."""
    n = 128
    N = n ** 3
    idx, row, col, val = gen.poisson_ccsr(n)
    prow, pcol, pval = gen.poisson_strip(3, n)
    xh = oracle.uniform_real(3, N)
    x, y1, y2 = vx.vector(ctx1, xh), vx.vector(ctx1, N), vx.vector(ctx1, N)
    y1.assign(vx.SpMatCCSR(ctx1, N, idx, row, col, val) * x)
    y2.assign(vx.SpMat(ctx1, N, N, prow, pcol, pval) * x)
    assert np.array_equal(y1.read(), y2.read())
