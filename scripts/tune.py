"""Parameter sweeps on one GPU (run under gpurun): SpMV tile sizes / formats, sweep grid sizes."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import vexcl_b200 as vx
from vexcl_b200 import gen, _lib as L
from vexcl_b200.api import Event

ctx = vx.Context([0])


def timeit(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    ctx.finish()
    e0, e1 = Event(ctx), Event(ctx)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.sync()
    return e0.elapsed_ms(e1) / reps


def spmv_sweep(dim, nx, label):
    row, col, val = gen.poisson_strip(dim, nx)
    N = row.size - 1
    nnz = int(row[-1])
    nbytes = gen.spmv_bytes(N, N, nnz)
    x, y = vx.vector(ctx, N), vx.vector(ctx, N)
    x.assign(vx.ElementIndex() * 1e-9 + 0.5)
    res = []
    cfgs = [(vx.FMT_HELL, 0, 0, 0, 0, 1)]
    cfgs += [(vx.FMT_CSR, tn, tr, 0, 0, 0) for (tn, tr) in ((1536, 384), (2048, 512), (2560, 512), (3072, 768))]
    cfgs += [(vx.FMT_CSR, tn, tr, st, 0, 1) for (tn, tr) in ((1024, 256), (2048, 512)) for st in (2, 3)]
    for fmt, tn, tr, stages, cps, pipe in cfgs:
        if fmt == vx.FMT_CSR:
            vx.set_param("spmv.tile_nnz", tn); vx.set_param("spmv.tile_rows", tr); vx.set_param("spmv.pipeline", pipe)
            vx.set_param("spmv.stages", stages or 4); vx.set_param("spmv.ctas_per_sm", cps)
        A = vx.SpMat(ctx, N, N, row, col, val, fmt)
        ms = timeit(lambda: A.apply(x, y, 1.0, False))
        ms_app = timeit(lambda: A.apply(x, y, 1.0, True))
        r = dict(case=label, fmt="hell" if fmt == vx.FMT_HELL else ("csr-pipe" if pipe else "csr-1shot"), tile_nnz=tn, tile_rows=tr, stages=stages, cps=cps, ms=ms, gbs=nbytes / ms / 1e6,
                 ms_append=ms_app, gbs_append=(nbytes + 8 * N) / ms_app / 1e6, n_tiles=int(A.info().loc.n_tiles))
        print(json.dumps(r), flush=True)
        res.append(r)
        del A
    return res


def vec_sweep():
    n = 100_000_000
    a, b, c, d = (vx.vector(ctx, n) for _ in range(4))
    for v in (a, b, c, d):
        v.assign(vx.ElementIndex() * 1e-8 + 0.25)
    s = vx.Reductor(ctx, np.float64, L.SUM)
    for bps in (2, 4, 6, 8, 12, 16, 32):
        for persistent in (1, 0):
            vx.set_param("sweep.blocks_per_sm", bps); vx.set_param("sweep.persistent", persistent)
            ms = timeit(lambda: a.assign(b + c * d), reps=30)
            ms2 = timeit(lambda: a.assign(0.5 * a + b), reps=30)
            print(json.dumps(dict(case="a=b+c*d", bps=bps, persistent=persistent, ms=ms, gbs=32 * n / ms / 1e6, saxpy_gbs=24 * n / ms2 / 1e6)), flush=True)
            if not persistent:
                break
    vx.set_param("sweep.blocks_per_sm", 8); vx.set_param("sweep.persistent", 1)
    for bps in (2, 4, 8, 16):
        vx.set_param("reduce.blocks_per_sm", bps)
        ms = timeit(lambda: s(a * b), reps=30)
        print(json.dumps(dict(case="sum(a*b)", bps=bps, ms=ms, gbs=16 * n / ms / 1e6)), flush=True)
    vx.set_param("reduce.blocks_per_sm", 4)
    vx.set_param("eval.force_interp", 1)
    for bps in (4, 6, 8):
        vx.set_param("interp.blocks_per_sm", bps)
        ms = timeit(lambda: a.assign(b + c * d), reps=20)
        print(json.dumps(dict(case="interp a=b+c*d", bps=bps, ms=ms, gbs=32 * n / ms / 1e6)), flush=True)
    ms = timeit(lambda: s(a * b), reps=20)
    print(json.dumps(dict(case="interp sum(a*b)", ms=ms, gbs=16 * n / ms / 1e6)), flush=True)
    ms = timeit(lambda: a.assign(vx.sin(b) * c + vx.sqrt(d)), reps=20)
    print(json.dumps(dict(case="interp sin(b)*c+sqrt(d)", ms=ms, gbs=32 * n / ms / 1e6)), flush=True)
    vx.set_param("eval.force_interp", 0)


if __name__ == "__main__":
    what = sys.argv[1:] or ["spmv2d", "vec"]
    if "spmv2d" in what:
        spmv_sweep(2, 3162, "poisson2d_3162")
    if "spmv3d" in what:
        spmv_sweep(3, 256, "poisson3d_256")
    if "vec" in what:
        vec_sweep()
