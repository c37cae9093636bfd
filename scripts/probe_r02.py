"""Round-2 timing probe (one GPU, CUDA events): the kernels changed since the last bench line.
    python scripts/probe_r02.py > gpurun_out/r02_probe.json"""
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
PEAK = 6584.5
out = {}


def timeit(fn, steps=30, warm=5):
    for _ in range(warm):
        fn()
    ctx.finish()
    e0, e1 = Event(ctx), Event(ctx)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); e1.sync()
    return e0.elapsed_ms(e1) / steps


what = sys.argv[1:] or ["ring", "multi", "ccsr", "p3d", "cg", "stencil"]

if "ring" in what:
    for label, (row, col, val) in (("irregular 4M U[0,32) gaps U[1,64)", gen.irregular_rows(4_000_000, 0, 32, seed=1)),
                                   ("poisson2d 3162^2", gen.poisson_strip(2, 3162))):
        n = row.size - 1
        nb = gen.spmv_bytes(n, n, int(row[-1]))
        x, y, y4 = vx.vector(ctx, n), vx.vector(ctx, n), vx.vector(ctx, n)
        x.assign(vx.ElementIndex() * (1.0 / n) + 0.5)
        A = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_CSR)
        res = {}
        vx.set_param("spmv.kernel", 4)
        A.apply(x, y4, 1.0, False)
        ref = y4.read()
        for k in (3, 4):
            vx.set_param("spmv.kernel", k)
            ms = timeit(lambda: A.apply(x, y, 1.0, False))
            res[f"kernel={k}"] = {"ms": ms, "frac": nb / ms / 1e6 / PEAK}
        vx.set_param("spmv.kernel", 5)
        for warps in (8, 4):
            for stages in (2, 3, 4):
                for ctas in (0, 1, 2, 3, 4):
                    vx.set_param("spmv.ring_warps", warps); vx.set_param("spmv.ring_stages", stages); vx.set_param("spmv.ctas_per_sm", ctas)
                    try:
                        y.assign(0.0)
                        A.apply(x, y, 1.0, False)
                        same = bool(np.array_equal(y.read(), ref))
                        ms = timeit(lambda: A.apply(x, y, 1.0, False))
                        res[f"kernel=5 warps={warps} stages={stages} ctas_per_sm={ctas}"] = {"ms": ms, "frac": nb / ms / 1e6 / PEAK, "same_bits_as_warp_tiles": same}
                    except vx.VexbError as e:
                        res[f"kernel=5 warps={warps} stages={stages} ctas_per_sm={ctas}"] = {"error": str(e)[:200]}
        vx.set_param("spmv.ctas_per_sm", 0); vx.set_param("spmv.kernel", -1)
        out[label] = res
        del A, x, y, y4

if "window" in what:
    # csr_window_kernel (spmv.kernel=6): tile size and window size are read when the matrix is built
    for label, gaps in (("irregular 4M U[0,32) gaps U[1,64)", 64), ("irregular 4M U[0,32) gaps U[1,8)", 8)):
        row, col, val = gen.irregular_rows(4_000_000, 0, 32, seed=1, max_gap=gaps)
        n = row.size - 1
        nb = gen.spmv_bytes(n, n, int(row[-1]))
        x, y, y4 = vx.vector(ctx, n), vx.vector(ctx, n), vx.vector(ctx, n)
        x.assign(vx.ElementIndex() * (1.0 / n) + 0.5)
        res = {}
        ref = None
        for tn, xw in ((2048, 2048), (1024, 2048), (4096, 4096), (2048, 4096), (1024, 1024)):
            vx.set_param("spmv.tile_nnz", tn); vx.set_param("spmv.xwin", xw)
            A = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_CSR)
            if ref is None:
                vx.set_param("spmv.kernel", 4)
                A.apply(x, y4, 1.0, False)
                ref = y4.read()
                for k in (3, 4):
                    vx.set_param("spmv.kernel", k)
                    ms = timeit(lambda: A.apply(x, y, 1.0, False))
                    res[f"kernel={k}"] = {"ms": ms, "frac": nb / ms / 1e6 / PEAK}
                vx.set_param("spmv.kernel", -1)
                res["default_variant_ms"] = timeit(lambda: A.apply(x, y, 1.0, False))
            vx.set_param("spmv.kernel", 6)
            y.assign(0.0)
            A.apply(x, y, 1.0, False)
            got = y.read()
            err = float(np.max(np.abs(got - ref) / (np.abs(ref) + 1e-300)))
            ms = timeit(lambda: A.apply(x, y, 1.0, False))
            res[f"kernel=6 tile_nnz={tn} xwin={xw}"] = {"ms": ms, "frac": nb / ms / 1e6 / PEAK, "max_rel_diff_vs_warp_tiles": err}
            vx.set_param("spmv.kernel", -1)
            del A
        vx.set_param("spmv.tile_nnz", 2048); vx.set_param("spmv.xwin", 2048)
        Ah = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_AUTO)
        ms = timeit(lambda: Ah.apply(x, y, 1.0, False))
        res["spmat_auto"] = {"ms": ms, "frac": nb / ms / 1e6 / PEAK, "fmt": int(Ah.info().loc.fmt)}
        del Ah
        out[label] = res
        del x, y, y4

if "multi" in what:
    row, col, val = gen.poisson_strip(2, 3162)
    N = row.size - 1
    A = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_AUTO)
    xs = [vx.vector(ctx, N) for _ in range(4)]
    ys = [vx.vector(ctx, N) for _ in range(4)]
    for r, v in enumerate(xs):
        v.assign(vx.ElementIndex() * (1.0 / N) + 0.25 * r)
    one = timeit(lambda: A.apply(xs[0], ys[0], 1.0, False))
    four = timeit(lambda: A.apply_multi(xs, ys, 1.0, False))
    res = {"ms_one": one, "ms_four": four, "ratio": four / one}
    out["multi_rhs"] = res
    del A, xs, ys

if "p3d" in what:
    row, col, val = gen.poisson_strip(3, 256)
    N = row.size - 1
    nb = gen.spmv_bytes(N, N, int(row[-1]))
    x, y = vx.vector(ctx, N), vx.vector(ctx, N)
    x.assign(vx.ElementIndex() * (1.0 / N) + 0.5)
    res = {}
    for c16 in (0, 1):
        vx.set_param("spmv.col16", c16)
        A = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_AUTO)
        ms = timeit(lambda: A.apply(x, y, 1.0, False))
        info = A.info().loc
        res[f"col16={c16}"] = {"ms": ms, "gbs_canonical": nb / ms / 1e6, "device_bytes": int(info.device_bytes), "ell_width": int(info.ell_width)}
        del A
    vx.set_param("spmv.col16", 1)
    out["poisson3d 256^3 hell"] = res
    del x, y

if "sell" in what:
    for label, gaps in (("irregular 4M U[0,32) gaps U[1,64)", 64), ("irregular 4M U[0,32) gaps U[1,8)", 8)):
        row, col, val = gen.irregular_rows(4_000_000, 0, 32, seed=1, max_gap=gaps)
        n = row.size - 1
        nb = gen.spmv_bytes(n, n, int(row[-1]))
        x, y, yr = vx.vector(ctx, n), vx.vector(ctx, n), vx.vector(ctx, n)
        x.assign(vx.ElementIndex() * (1.0 / n) + 0.5)
        res = {}
        A = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_CSR)
        vx.set_param("spmv.kernel", 3)
        A.apply(x, yr, 1.0, False)                            # thread per row: storage order
        ref = yr.read()
        vx.set_param("spmv.kernel", 4)
        ms = timeit(lambda: A.apply(x, y, 1.0, False))
        res["csr warp tiles"] = {"ms": ms, "frac": nb / ms / 1e6 / PEAK}
        vx.set_param("spmv.kernel", -1)
        del A
        A = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_HELL)
        ms = timeit(lambda: A.apply(x, y, 1.0, False))
        res["hybrid ell"] = {"ms": ms, "frac": nb / ms / 1e6 / PEAK, "device_mb": int(A.info().loc.device_bytes) / 1e6}
        del A
        for sigma in (1024, 256, 8192):
            vx.set_param("spmv.sell_sigma", sigma)
            A = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_SELL)
            y.assign(0.0)
            A.apply(x, y, 1.0, False)
            same = bool(np.array_equal(y.read(), ref))
            ms = timeit(lambda: A.apply(x, y, 1.0, False))
            res[f"sliced ell sigma={sigma}"] = {"ms": ms, "frac": nb / ms / 1e6 / PEAK, "same_bits_as_thread_per_row": same,
                                                "device_mb": int(A.info().loc.device_bytes) / 1e6}
            del A
        vx.set_param("spmv.sell_sigma", 1024)
        A = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_AUTO)
        ms = timeit(lambda: A.apply(x, y, 1.0, False))
        res["auto"] = {"ms": ms, "frac": nb / ms / 1e6 / PEAK, "fmt": int(A.info().loc.fmt)}
        del A, x, y, yr
        out[label] = res

if "small" in what:
    # one GPU's share of the strong-scaled configs at N = 8, WITHOUT a halo: what the product costs when nothing is exchanged
    res = {}
    for label, dim, dims in (("configs[2]/8: 3162 x 396", 2, (3162, 396, None)), ("configs[3]/8: 256 x 256 x 32", 3, (256, 256, 32))):
        row, col, val = gen.poisson_strip(dim, *[d for d in dims if d is not None])
        N = row.size - 1
        A = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_AUTO)
        x, y = vx.vector(ctx, N), vx.vector(ctx, N)
        x.assign(vx.ElementIndex() * (1.0 / N) + 0.5)
        ms = timeit(lambda: A.apply(x, y, 1.0, False), steps=200, warm=20)
        from vexcl_b200.api import Graph
        g = Graph(ctx, lambda: [A.apply(x, y, 1.0, False) for _ in range(10)])
        msg = timeit(g.launch, steps=40, warm=5) / 10
        res[label] = {"rows": N, "us_stream": ms * 1e3, "us_graph_of_10": msg * 1e3, "format_mb": int(A.info().loc.device_bytes) / 1e6}
        del g, A, x, y
    out["one eighth of the strong configs, no halo"] = res

if "ccsr" in what:
    n = 256
    N = n ** 3
    idx, row, col, val = gen.poisson_ccsr(n)
    x, y = vx.vector(ctx, N), vx.vector(ctx, N)
    x.assign(vx.ElementIndex() * (1.0 / N) + 0.5)
    y.assign(0.0)
    res = {}
    for jit in (0, 1):
        vx.set_param("ccsr.jit", jit)
        A = vx.SpMatCCSR(ctx, N, idx, row, col, val)
        for app in (True, False):
            ms = timeit(lambda: A.apply(x, y, 1.0, app))
            comp = N * (1 + 16 + (8 if app else 0))
            res[f"jit={jit} append={app}"] = {"ms": ms, "gbs_compulsory": comp / ms / 1e6, "frac": comp / ms / 1e6 / PEAK}
        del A
    vx.set_param("ccsr.jit", 1)
    out["ccsr 256^3"] = res
    del x, y

if "cg" in what:
    from vexcl_b200.solvers import CGFused
    n = 256
    N = n ** 3
    row, col, val = gen.poisson_strip(3, n, spd=True)
    A = vx.SpMat(ctx, N, N, row, col, val)
    b, x = vx.vector(ctx, N), vx.vector(ctx, N)
    b.assign(((vx.ElementIndex() * 2654435761) % 1000003) * (1.0 / 1000003) - 0.5)
    x.assign(0.0)
    cg = CGFused(A, b, x)
    cg.capture()
    ms = timeit(lambda: cg.run(1), steps=40)
    out["cg fused 256^3 graph"] = {"ms_per_iteration": ms, "residual2": cg.residual2()}
    del cg, A, b, x

if "stencil" in what:
    n, width = 1 << 26, 21
    S = vx.stencil(ctx, np.full(width, 1.0 / width), width // 2)
    a, b = vx.vector(ctx, n), vx.vector(ctx, n)
    a.assign(vx.ElementIndex() * (1.0 / n) + 0.5)
    res = {}
    for k, nm in ((1, "one block per tile"), (0, "pipelined persistent blocks")):
        vx.set_param("stencil.kernel", k)
        ms = timeit(lambda: S.apply(a, b, 1.0, False))
        res[nm] = {"ms": ms, "gbs_compulsory": 16 * n / ms / 1e6, "frac": 16 * n / ms / 1e6 / PEAK}
    vx.set_param("stencil.kernel", 1)
    out["stencil 2^26 w21"] = res

print(json.dumps(out, indent=1))
