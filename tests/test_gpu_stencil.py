"""vex::stencil on the GPU against oracle/stencil.py (tests/stencil.cpp, examples/benchmark.cpp:281-349): bit-exact,
because both accumulate the taps in order without contraction.  One slice, and several slices on one device (the
reference fixture's trick) so that the halo copies run."""
import numpy as np
import pytest

import oracle
from oracle import stencil as ost
import vexcl_b200 as vx

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 2, 3])
def anyctx(request, ctx1, ctx2, ctx3):
    return {1: ctx1, 2: ctx2, 3: ctx3}[request.param]


@pytest.mark.parametrize("n,width,center", [(1024, 21, 10), (1024, 1, 0), (1000, 64, 0), (1000, 64, 63), (4099, 9, 4),
                                            (128, 57, 31), (65536, 2048, 700), (1, 5, 2), (40, 41, 20), (3000, 8, 3), (3000, 16, 15)])
def test_stencil_convolution(anyctx, n, width, center):
    s = oracle.uniform_real(width, width)
    xh = oracle.uniform_real(n + 1, n)
    S = vx.stencil(anyctx, s, center)
    x, y = vx.vector(anyctx, xh), vx.vector(anyctx, n)
    want = ost.convolve(s, center, xh)
    y.assign(x * S)
    assert np.array_equal(y.read(), want)
    y.assign(1.0)
    y += S * x                                             # stencil.cpp:32-33
    assert np.array_equal(y.read(), 1.0 + want)
    y.assign(42 * (x * S))                                 # stencil.cpp:47
    assert np.array_equal(y.read(), 42.0 * want)
    y.assign(x * S + x * S)                                # stencil.cpp:69
    assert np.array_equal(y.read(), want + want)
    y.assign(x - 0.5 * (x * S))
    assert np.array_equal(y.read(), xh + (-0.5) * want)


def test_stencil_single_precision_and_benchmark_size(ctx1):
    n = 1 << 20                                            # benchmark.cpp:286-295: N = 1M, 21 taps of 1/21
    s = np.full(21, 1 / 21, dtype=np.float32)
    xh = oracle.uniform_real(5, n).astype(np.float32)
    S = vx.stencil(ctx1, s, 10, dtype=np.float32)
    x, y = vx.vector(ctx1, xh), vx.vector(ctx1, n, dtype=np.float32)
    y.assign(x * S)
    assert np.array_equal(y.read(), ost.convolve(s, 10, xh))
    sd = np.full(21, 1 / 21)
    xd = oracle.uniform_real(6, n)
    Sd = vx.stencil(ctx1, sd, 10)
    x, y = vx.vector(ctx1, xd), vx.vector(ctx1, n)
    y.assign(x * Sd)
    assert np.array_equal(y.read(), ost.convolve(sd, 10, xd))


def test_stencil_argument_checks(ctx1):
    with pytest.raises(ValueError):
        vx.stencil(ctx1, [1.0, 2.0], 2)
    with pytest.raises(ValueError):
        vx.stencil(ctx1, [], 0)


@pytest.mark.parametrize("nparts", [1, 2])
def test_pipelined_stencil_kernel_same_bits(ctx1, ctx2, nparts):
    """stencil.kernel = 0: persistent blocks with the next window arriving by cp.async (csrc/stencil.cu).  Opt-in (measured
    slower than one block per tile), but selectable, so it is checked: more tiles than resident blocks, ragged end, both
    precisions, `=` and `+=`, one and two slices (halo buffers as cp.async sources)."""
    ctx = {1: ctx1, 2: ctx2}[nparts]
    vx.set_param("stencil.kernel", 0)
    try:
        for n, width, center, dt in ((3_000_017, 21, 10, np.float64), (2_500_003, 5, 0, np.float64), (2_400_001, 33, 32, np.float32)):
            s = oracle.uniform_real(width, width).astype(dt)
            xh = oracle.uniform_real(n % 97 + 1, n).astype(dt)
            S = vx.stencil(ctx, s, center, dtype=dt)
            x, y = vx.vector(ctx, xh), vx.vector(ctx, n, dtype=dt)
            want = ost.convolve(s, center, xh)
            y.assign(x * S)
            assert np.array_equal(y.read(), want)
            y.assign(1.0)
            y += S * x
            assert np.array_equal(y.read(), (dt(1.0) + want).astype(dt))
    finally:
        vx.set_param("stencil.kernel", 1)
