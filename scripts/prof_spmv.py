"""Small driver for ncu: 2-D Poisson 3162^2, a few passes of each SpMV kernel (and the vector kernels)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import vexcl_b200 as vx
from vexcl_b200 import gen, _lib as L

ctx = vx.Context([0])
what = sys.argv[1:] or ["csr", "hell"]
n = 3162
row, col, val = gen.poisson_strip(2, n)
N = row.size - 1
x, y = vx.vector(ctx, N), vx.vector(ctx, N)
x.assign(vx.ElementIndex() * 1e-9 + 0.5)
for w in what:
    if w in ("csr", "csr1", "hell"):
        vx.set_param("spmv.pipeline", 1 if w == "csrpipe" else 0)
        A = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_HELL if w == "hell" else vx.FMT_CSR)
        for _ in range(4):
            A.apply(x, y, 1.0, False)
        ctx.finish()
        del A
    if w == "vec":
        m = 100_000_000
        a, b, c, d = (vx.vector(ctx, m) for _ in range(4))
        for v in (a, b, c, d):
            v.assign(vx.ElementIndex() * 1e-8 + 0.25)
        s = vx.Reductor(ctx, np.float64, L.SUM)
        for _ in range(3):
            a.assign(b + c * d)
            a += b + c * d
            s(a * b)
        ctx.finish()
