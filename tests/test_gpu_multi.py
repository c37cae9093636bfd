"""Multi-GPU paths (skipped on a single-GPU box): NCCL halo exchange and all-reduce, one process
driving several devices, and the one-process-per-GPU launch that bench.py uses."""
import ctypes as C
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle
import vexcl_b200 as vx
from vexcl_b200 import _lib as L

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def ndev():
    n = C.c_int(0)
    L.lib().vexb_device_count(C.byref(n))
    return n.value


@pytest.fixture(scope="module", params=["nccl", "peer"])
def ctxn(request, built):
    """Several GPUs driven by one process: halo over NCCL send/recv, or pushed through peer memory inside the product
    kernel (the default on distinct devices)."""
    n = ndev()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    if request.param == "nccl":
        return vx.Context(list(range(min(n, 4))), use_nccl=True, peer_halo=False)
    return vx.Context(list(range(min(n, 4))), use_peer=True)


def test_nccl_reduce_and_halo_single_process(ctxn):
    nd = ctxn.nparts
    n = 40000
    X = oracle.uniform_real(3, n)
    x = vx.vector(ctxn, X)
    s = vx.Reductor(ctxn, np.float64, L.SUM)
    ref = oracle.kahan_sum(X)
    assert abs(s(x) - ref) <= 1e-10 * abs(ref)
    assert vx.Reductor(ctxn, np.float64, L.MAX)(x) == X.max()
    assert vx.Reductor(ctxn, np.float64, L.MIN)(x) == X.min()
    assert vx.Reductor(ctxn, np.float64, L.MINMAX)(x) == (X.min(), X.max())
    assert vx.Reductor(ctxn, np.int64, L.SUM)(x > 0.5) == int((X > 0.5).sum())
    for dim, m in ((2, 200), (3, 34)):
        row, col, val = oracle.poisson(dim, m)
        N = row.size - 1
        Xv = oracle.uniform_real(7, N)
        part = oracle.partition(N, nd)
        want = oracle.spmat_apply(part, part, row, col, val, Xv)
        for fmt in (vx.FMT_CSR, vx.FMT_HELL):
            A = vx.SpMat(ctxn, N, N, row, col, val, fmt)
            xv, y = vx.vector(ctxn, Xv), vx.vector(ctxn, N)
            for _ in range(3):                                   # repeated applies reuse the halo buffers
                y.assign(A * xv)
            mag = oracle.csr_absrow(row, col, val, Xv)
            assert A.peer_halo == ctxn.peer_halo
            assert np.all(np.abs(y.read() - want) <= 1e-10 * mag)
            if A.peer_halo and fmt == vx.FMT_HELL:
                n0 = vx.launch_count()
                A.apply(xv, y, 1.0, False)
                assert vx.launch_count() - n0 == nd           # ONE kernel per device and product
            y += 2 * (A * xv)
            assert np.all(np.abs(y.read() - 3 * want) <= 3e-10 * mag)
    row, col, val = oracle.random_matrix(5000, 5000, 12, seed=77)
    Xv = oracle.uniform_real(8, 5000)
    A = vx.SpMat(ctxn, 5000, 5000, row, col, val)
    xv, y = vx.vector(ctxn, Xv), vx.vector(ctxn, 5000)
    y.assign(A * xv)
    want = oracle.csr_spmv(row, col, val, Xv)
    assert np.all(np.abs(y.read() - want) <= 1e-10 * np.abs(want) + 1e-300)


def test_fused_reduce_allreduce_over_peer_memory(built):
    """vexb_reduce_all: the last block of each rank's reduction kernel exchanges partials through peer
    mailboxes (no NCCL); every rank must end with the same bits, and the sum must match the oracle."""
    n = ndev()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = vx.Context(list(range(min(n, 8))), use_peer=True)
    N = 1_000_003
    X = oracle.uniform_real(5, N)
    x = vx.vector(ctx, X)
    ref = oracle.kahan_sum(X)
    s = vx.Reductor(ctx, np.float64, L.SUM)
    for _ in range(5):                                           # epochs advance, parity alternates
        assert abs(s(x) - ref) <= 1e-10 * abs(ref)
    assert vx.Reductor(ctx, np.float64, L.MAX)(x) == X.max()
    assert vx.Reductor(ctx, np.float64, L.MINMAX)(x) == (X.min(), X.max())
    assert vx.Reductor(ctx, np.int64, L.SUM)(x > 0.25) == int((X > 0.25).sum())
    from vexcl_b200.api import DeviceScalar
    d = DeviceScalar(ctx)
    s.device(x * x, d)
    ctx.finish()
    vals = []
    for k in ctx.local:                                          # identical bits on every device
        h = np.empty(1)
        L.check(L.lib().vexb_d2h(ctx.devs[k], h.ctypes.data, d.bufs[k], 8, ctx.streams[k], 1))
        vals.append(h[0])
    assert len(set(vals)) == 1 and abs(vals[0] - np.dot(X, X)) <= 1e-10 * np.dot(X, X)
    # an empty slice still takes part (n = 1 over several devices)
    one = vx.vector(ctx, 1)
    one.assign(42.0)
    assert s(one) == 42.0
    err = __import__("ctypes").c_uint64(7)
    L.check(L.lib().vexb_peer_error(ctx.peers[0], __import__("ctypes").byref(err)))
    assert err.value == 0


def test_peer_halo_many_products_graph_and_fused_cg(built):
    """Epoch parity, acknowledgements and CUDA-graph replay of the fused product over many iterations; then the fused CG
    iteration (product + dot combined across the GPUs inside the kernel) against the oracle."""
    n = ndev()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    from vexcl_b200.api import Graph
    from vexcl_b200.solvers import CGFused
    ctx = vx.Context(list(range(min(n, 8))), use_peer=True)
    row, col, val = oracle.poisson(3, 40)
    N = row.size - 1
    Xv = oracle.uniform_real(7, N)
    A = vx.SpMat(ctx, N, N, row, col, val)
    assert A.peer_halo
    x, y = vx.vector(ctx, Xv), vx.vector(ctx, N)
    want = oracle.csr_spmv(row, col, val, Xv)
    mag = oracle.csr_absrow(row, col, val, Xv)
    for _ in range(7):
        A.apply(x, y, 1.0, False)
    assert np.all(np.abs(y.read() - want) <= 1e-10 * mag)
    g = Graph(ctx, lambda: A.apply(x, y, 1.0, False))
    for it in range(40):                                         # x changes between replays: stale ghosts would show
        x.assign(x * 0.5 + 0.25)
        Xv = Xv * 0.5 + 0.25
        g.launch()
    ctx.finish()
    assert np.all(np.abs(y.read() - oracle.csr_spmv(row, col, val, Xv)) <= 1e-10 * oracle.csr_absrow(row, col, val, Xv))
    # fused CG on an SPD problem
    from vexcl_b200 import gen
    row, col, val = gen.poisson_strip(3, 24, 20, 8 * ctx.nparts, spd=True)
    N = row.size - 1
    b = oracle.uniform_real(3, N)
    xo, hist_o = oracle.cg(row, col, val, b, np.zeros(N), 20)
    A = vx.SpMat(ctx, N, N, row, col, val)
    bv, xs = vx.vector(ctx, b), vx.vector(ctx, N)
    xs.assign(0.0)
    cg = CGFused(A, bv, xs)
    cg.capture()
    hist = []
    for _ in range(18):
        cg.run(1)
        hist.append(cg.residual2())
    assert cg.fused_product
    assert np.allclose(hist, hist_o[2:], rtol=1e-8)
    assert np.allclose(xs.read(), xo, rtol=1e-8, atol=1e-12)
    e = C.c_uint64(7)
    L.check(L.lib().vexb_peer_fault(C.byref(e), 0))
    assert e.value == 0


def test_copy_engine_halo_without_nccl(built):
    if ndev() < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = vx.Context([0, 1], use_nccl=False, peer_halo=False)
    row, col, val = oracle.poisson(2, 150)
    N = row.size - 1
    Xv = oracle.uniform_real(7, N)
    A = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_CSR)
    xv, y = vx.vector(ctx, Xv), vx.vector(ctx, N)
    y.assign(A * xv)
    want = oracle.csr_spmv(row, col, val, Xv)
    assert np.all(np.abs(y.read() - want) <= 1e-10 * oracle.csr_absrow(row, col, val, Xv))


def test_bench_one_process_per_gpu(built):
    n = ndev()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(ROOT / "bench.py"), "--gpus", "2", "--steps", "20", "--warmup", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["n_gpus"] == 2 and out["value"] > 0 and out["gpu_launches"] >= 20
