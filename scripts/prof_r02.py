"""Driver for ncu captures (round 2): a few launches of each hot kernel on its bench workload.

    python scripts/prof_r02.py hell csr_poisson csr_irregular ccsr ccsr_jit patterns stencil vec cg fused_product multi_rhs

Every target warms its kernel twice and then launches it three times, so `ncu -k regex:<name> -s 2 -c 3` lands on steady
launches.  Never a bench: numbers taken under a profiler are not throughput figures."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import vexcl_b200 as vx
from vexcl_b200 import gen, _lib as L

what = sys.argv[1:] or ["hell", "csr_poisson", "csr_irregular", "ccsr", "stencil"]
ctx = vx.Context([0])
REPS = 5


def poisson2():
    row, col, val = gen.poisson_strip(2, 3162)
    N = row.size - 1
    x, y = vx.vector(ctx, N), vx.vector(ctx, N)
    x.assign(vx.ElementIndex() * (1.0 / N) + 0.5)
    return row, col, val, N, x, y


if {"hell", "csr_poisson", "patterns", "fused_product", "multi_rhs"} & set(what):
    row, col, val, N, x, y = poisson2()
    if "hell" in what:
        A = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_AUTO)
        for _ in range(REPS):
            A.apply(x, y, 1.0, False)
        ctx.finish()
        if "fused_product" in what:
            for _ in range(REPS + 4):                       # the NVRTC kernel takes over after the first uses
                y.assign(x + A * x)
                ctx.finish()
        if "multi_rhs" in what:
            xs = [vx.vector(ctx, N) for _ in range(4)]
            ys = [vx.vector(ctx, N) for _ in range(4)]
            for r, v in enumerate(xs):
                v.assign(vx.ElementIndex() * (1.0 / N) + 0.25 * r)
            for _ in range(REPS):
                A.apply_multi(xs, ys, 1.0, False)
            ctx.finish()
            del xs, ys
        del A
    if "csr_poisson" in what:
        A = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_CSR)
        for k in (3, 4, 0):
            vx.set_param("spmv.kernel", k)
            for _ in range(REPS):
                A.apply(x, y, 1.0, False)
        vx.set_param("spmv.kernel", -1)
        ctx.finish()
        del A
    if "patterns" in what:
        for jit in (0, 1):
            vx.set_param("ccsr.jit", jit)
            A = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_PATTERNS)
            for _ in range(REPS):
                A.apply(x, y, 1.0, False)
            ctx.finish()
            del A
        vx.set_param("ccsr.jit", 0)
    del row, col, val, x, y

if "csr_irregular" in what:
    n = 4_000_000
    row, col, val = gen.irregular_rows(n, 0, 32, seed=1)
    xi, yi = vx.vector(ctx, n), vx.vector(ctx, n)
    xi.assign(vx.ElementIndex() * (1.0 / n) + 0.5)
    Ai = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_CSR)
    for k in (4, 3, 0):
        vx.set_param("spmv.kernel", k)
        for _ in range(REPS):
            Ai.apply(xi, yi, 1.0, False)
    vx.set_param("spmv.kernel", -1)
    ctx.finish()
    del Ai, xi, yi

if "sell" in what:
    n = 4_000_000
    row, col, val = gen.irregular_rows(n, 0, 32, seed=1)
    xi, yi = vx.vector(ctx, n), vx.vector(ctx, n)
    xi.assign(vx.ElementIndex() * (1.0 / n) + 0.5)
    Ai = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_AUTO)          # uneven rows: sliced ELL
    for _ in range(REPS):
        Ai.apply(xi, yi, 1.0, False)
    ctx.finish()
    del Ai, xi, yi

if "ccsr" in what or "ccsr_jit" in what:
    n = 256
    N = n ** 3
    idx, row, col, val = gen.poisson_ccsr(n)
    x, y = vx.vector(ctx, N), vx.vector(ctx, N)
    x.assign(vx.ElementIndex() * (1.0 / N) + 0.5)
    y.assign(0.0)
    for jit in ([0] if "ccsr" in what else []) + ([1] if "ccsr_jit" in what else []):
        vx.set_param("ccsr.jit", jit)
        A = vx.SpMatCCSR(ctx, N, idx, row, col, val)
        for _ in range(REPS):
            A.apply(x, y, 1.0, True)
        ctx.finish()
        del A
    vx.set_param("ccsr.jit", 0)
    del x, y

if "stencil" in what:
    n, width = 1 << 26, 21
    S = vx.stencil(ctx, np.full(width, 1.0 / width), width // 2)
    a, b = vx.vector(ctx, n), vx.vector(ctx, n)
    a.assign(vx.ElementIndex() * (1.0 / n) + 0.5)
    for _ in range(REPS):
        S.apply(a, b, 1.0, False)
    ctx.finish()
    del a, b, S

if "vec" in what:
    m = 100_000_000
    a, b, c, d = (vx.vector(ctx, m) for _ in range(4))
    for v in (a, b, c, d):
        v.assign(vx.ElementIndex() * 1e-8 + 0.25)
    s = vx.Reductor(ctx, np.float64, L.SUM)
    vx.set_param("eval.jit", 0)
    for _ in range(REPS):
        a.assign(b + c * d)
        s(a * b)
        a.assign((b - c) * (b + c) / d + b * 0.5 + c * d)      # interpreter
    ctx.finish()
    vx.set_param("eval.jit", 2)
    del a, b, c, d

if "cg" in what:
    from vexcl_b200.solvers import CGDevice, CGFused
    n = 256
    N = n ** 3
    row, col, val = gen.poisson_strip(3, n, spd=True)
    A = vx.SpMat(ctx, N, N, row, col, val)
    del row, col, val
    b, x = vx.vector(ctx, N), vx.vector(ctx, N)
    for cls in (CGFused, CGDevice):
        b.assign(vx.ElementIndex() * (1.0 / N) + 0.5)
        x.assign(0.0)
        cg = cls(A, b, x)
        cg.run(REPS)
        ctx.finish()
        del cg
