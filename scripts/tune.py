"""Parameter sweeps on one GPU (run under gpurun): SpMV kernel variants / tile sizes / formats, vector kernels."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import vexcl_b200 as vx
from vexcl_b200 import gen, _lib as L
from vexcl_b200.api import Event

ctx = vx.Context([0])
VARIANT = {0: "csr-tma", 1: "csr-pipe", 2: "csr-direct"}


def timeit(fn, reps=100, warm=5):
    for _ in range(warm):
        fn()
    ctx.finish()
    e0, e1 = Event(ctx), Event(ctx)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.sync()
    return e0.elapsed_ms(e1) / reps


def spmv_sweep(row, col, val, label, cfgs=None):
    N = row.size - 1
    nnz = int(row[-1])
    nbytes = gen.spmv_bytes(N, N, nnz)
    x, y = vx.vector(ctx, N), vx.vector(ctx, N)
    x.assign(vx.ElementIndex() * 1e-9 + 0.5)
    if cfgs is None:
        cfgs = [("hell", 0, 0, 0)]
        cfgs += [("csr", 0, tn, tr) for (tn, tr) in ((2048, 512),)]
        cfgs += [("csr", 2, tn, tr) for (tn, tr) in ((1024, 256), (1536, 384), (2048, 512), (2048, 1024))]
    for fmt, variant, tn, tr in cfgs:
        if fmt == "csr":
            vx.set_param("spmv.tile_nnz", tn); vx.set_param("spmv.tile_rows", tr); vx.set_param("spmv.kernel", variant)
        A = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_HELL if fmt == "hell" else vx.FMT_CSR)
        ms = min(timeit(lambda: A.apply(x, y, 1.0, False)) for _ in range(3))
        ms_app = min(timeit(lambda: A.apply(x, y, 1.0, True)) for _ in range(2))
        r = dict(case=label, fmt="hell" if fmt == "hell" else VARIANT[variant], tile_nnz=tn, tile_rows=tr, ms=ms, gbs=nbytes / ms / 1e6,
                 gbs_append=(nbytes + 8 * N) / ms_app / 1e6)
        print(json.dumps(r), flush=True)
        del A


def random_rows(n, m, avg, seed):
    """Irregular matrix: row widths U[0, 2*avg], sorted random columns within a band of 64*avg around the diagonal."""
    rng = np.random.default_rng(seed)
    w = rng.integers(0, 2 * avg + 1, n)
    row = np.concatenate([[0], np.cumsum(w)]).astype(np.int64)
    base = np.repeat(np.arange(n, dtype=np.int64), w)
    off = rng.integers(-32 * avg, 32 * avg + 1, base.size)
    col = np.clip(base + off, 0, m - 1)
    # sort columns within rows
    key = base * (m + 1) + col
    col = col[np.argsort(key, kind="stable")]
    return row, col, rng.random(col.size)


def vec_sweep():
    n = 100_000_000
    a, b, c, d = (vx.vector(ctx, n) for _ in range(4))
    for v in (a, b, c, d):
        v.assign(vx.ElementIndex() * 1e-8 + 0.25)
    s = vx.Reductor(ctx, np.float64, L.SUM)
    for name, fn, bpe in (("a=b+c*d", lambda: a.assign(b + c * d), 32), ("a+=b+c*d", lambda: a.__iadd__(b + c * d), 40),
                          ("a=alpha*a+b", lambda: a.assign(0.5 * a + b), 24), ("sum(a*b)", lambda: s(a * b), 16)):
        ms = timeit(fn, reps=30)
        print(json.dumps(dict(case=name, ms=ms, gbs=bpe * n / ms / 1e6)), flush=True)
    vx.set_param("eval.force_interp", 1)
    for name, fn, bpe in (("interp a=b+c*d", lambda: a.assign(b + c * d), 32), ("interp sum(a*b)", lambda: s(a * b), 16),
                          ("interp sin(b)*c+sqrt(d)", lambda: a.assign(vx.sin(b) * c + vx.sqrt(d)), 32)):
        ms = timeit(fn, reps=20)
        print(json.dumps(dict(case=name, ms=ms, gbs=bpe * n / ms / 1e6)), flush=True)
    vx.set_param("eval.force_interp", 0)


if __name__ == "__main__":
    what = sys.argv[1:] or ["spmv2d", "vec"]
    if "spmv2d" in what:
        spmv_sweep(*gen.poisson_strip(2, 3162), "poisson2d_3162")
    if "spmv3d" in what:
        spmv_sweep(*gen.poisson_strip(3, 256), "poisson3d_256")
    if "irregular" in what:
        spmv_sweep(*random_rows(4_000_000, 4_000_000, 12, 1), "random_avg12")
        spmv_sweep(*random_rows(1_000_000, 1_000_000, 60, 2), "random_avg60")
    if "vec" in what:
        vec_sweep()
