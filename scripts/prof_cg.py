"""Driver for an ncu launch list: a few fused / unfused CG iterations on the SPD 256^3 operator (configs[4], 1 GPU) and the
CSR kernels on the irregular matrix.   ncu --metrics gpu__time_duration.sum --csv --log-file out.csv python scripts/prof_cg.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import vexcl_b200 as vx
from vexcl_b200 import gen
from vexcl_b200.solvers import CGDevice, CGFused

what = sys.argv[1:] or ["cg", "csr"]
ctx = vx.Context([0])
if "cg" in what:
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
        cg.run(4)
        ctx.finish()
        del cg
    y = vx.vector(ctx, N)
    for _ in range(3):
        A.apply(b, y, 1.0, False)
    ctx.finish()
    del A, b, x, y
if "csr" in what:
    n = 4_000_000
    row, col, val = gen.irregular_rows(n, 0, 32, seed=1)
    xi, yi = vx.vector(ctx, n), vx.vector(ctx, n)
    xi.assign(vx.ElementIndex() * (1.0 / n) + 0.5)
    Ai = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_CSR)
    for k in (4, 3):
        vx.set_param("spmv.kernel", k)
        for _ in range(3):
            Ai.apply(xi, yi, 1.0, False)
    ctx.finish()
    vx.set_param("spmv.kernel", -1)
    row, col, val = gen.poisson_strip(2, 3162)
    N = row.size - 1
    xp, yp = vx.vector(ctx, N), vx.vector(ctx, N)
    xp.assign(vx.ElementIndex() * (1.0 / N) + 0.5)
    Ap = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_CSR)
    for k in (4, 3):
        vx.set_param("spmv.kernel", k)
        for _ in range(3):
            Ap.apply(xp, yp, 1.0, False)
    ctx.finish()
