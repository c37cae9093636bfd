"""N>1 host logic with one process per part, world_size 2 and 3 over gloo (no GPU).

Each rank holds only its row strip (global column ids), finds its ghost columns with the C ABI,
all-gathers the lists, builds the halo plan, and must arrive at exactly the tables the oracle's
single-process restatement of SpMat::setup_exchange produces (spmat.hpp:291-378).  The halo exchange
itself is then played over gloo send/recv with the plan's pairwise counts, with the oracle doing the
strip arithmetic, and must reproduce the full product; the Reductor combine is an all-reduce of the
per-rank partials.  On the GPU box the same plan drives ncclSend/ncclRecv (vexb_halo_exchange).
"""
import ctypes as C
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import oracle
    import vexcl_b200 as vx
    from vexcl_b200 import gen, _lib as L
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = L.lib()
    # 2-D Poisson n x (n*world) grid (the weak-scaling layout of bench.py) + a random matrix
    for case in ("poisson", "random"):
        if case == "poisson":
            N = n * n * world
            part = vx.partition(N, world)
            row, col, val = gen.poisson_strip(2, n, n * world, r0=int(part[rank]), r1=int(part[rank + 1]))
            frow, fcol, fval = gen.poisson_strip(2, n, n * world)
        else:
            N = 700
            part = vx.partition(N, world)
            frow, fcol, fval = oracle.random_matrix(N, N, 9, seed=99)
            a, b = int(part[rank]), int(part[rank + 1])
            row, col, val = frow[a:b + 1] - frow[a], fcol[frow[a]:frow[b]], fval[frow[a]:frow[b]]
        nloc = int(part[rank + 1] - part[rank])
        row = np.ascontiguousarray(row); col = np.ascontiguousarray(col)
        cnt = C.c_size_t(0)
        L.check(lib.vexb_strip_ghost_cols(nloc, row.ctypes.data, 8, col.ctypes.data, 8, int(part[rank]), int(part[rank + 1]), None, C.byref(cnt)))
        g = np.empty(cnt.value, np.int64)
        cap = C.c_size_t(cnt.value)
        L.check(lib.vexb_strip_ghost_cols(nloc, row.ctypes.data, 8, col.ctypes.data, 8, int(part[rank]), int(part[rank + 1]), g.ctypes.data, C.byref(cap)))
        allg = [None] * world
        dist.all_gather_object(allg, g)
        off = np.zeros(world + 1, np.uint64)
        for d in range(world):
            off[d + 1] = off[d] + len(allg[d])
        cat = np.ascontiguousarray(np.concatenate(allg), dtype=np.int64) if off[-1] else np.empty(0, np.int64)
        cp = (C.c_size_t * (world + 1))(*[int(x) for x in part])
        go = (C.c_size_t * (world + 1))(*[int(x) for x in off])
        plan = C.c_void_p()
        L.check(lib.vexb_halo_plan_create(world, cp, cat.ctypes.data, go, C.byref(plan)))
        # reference tables from the whole matrix, single process
        ex = oracle.setup_exchange(part, part, frow, fcol)
        assert np.array_equal(g, ex["ghost"][rank])
        tot = C.c_size_t()
        L.check(lib.vexb_halo_plan_ref_sizes(plan, C.byref(tot)))
        cts = np.empty(tot.value, np.int64)
        cidx = (C.c_size_t * (world + 1))()
        L.check(lib.vexb_halo_plan_ref_tables(plan, cts.ctypes.data, cidx))
        assert np.array_equal(cts, ex["cols_to_send"]) and list(cidx) == list(ex["cidx"])
        sc, rc = (C.c_size_t * world)(), (C.c_size_t * world)()
        L.check(lib.vexb_halo_plan_counts(plan, rank, sc, rc))
        send_cols = np.empty(sum(sc), np.int64)
        if send_cols.size:
            L.check(lib.vexb_halo_plan_send_cols(plan, rank, send_cols.ctypes.data))
        # play SpMat::apply: pack, exchange over gloo, local + remote products with the oracle
        x_full = oracle.uniform_real(11, N)
        x_loc = x_full[part[rank]:part[rank + 1]]
        packed = torch.from_numpy(x_loc[send_cols].copy())
        ghost_vals = torch.empty(len(g), dtype=torch.float64)
        reqs, so, ro = [], 0, 0
        for p in range(world):
            if sc[p]:
                reqs.append(dist.isend(packed[so:so + sc[p]], dst=p))
            if rc[p]:
                reqs.append(dist.irecv(ghost_vals[ro:ro + rc[p]], src=p))
            so += sc[p]; ro += rc[p]
        for r in reqs:
            r.wait()
        assert np.array_equal(ghost_vals.numpy(), x_full[g])
        lr, lc, lv, rr, rcl, rv = oracle.split_strip(row, col, val, 0, nloc, int(part[rank]), int(part[rank + 1]), g)
        y = oracle.csr_spmv(lr, lc, lv, x_loc) if lc.size else np.zeros(nloc)
        if rcl.size:
            y = oracle.csr_spmv(rr, rcl, rv, ghost_vals.numpy(), y, 1.0, True)
        want = oracle.spmat_apply(part, part, frow, fcol, fval, x_full)[part[rank]:part[rank + 1]]
        assert np.array_equal(y, want)
        # Reductor combine: per-rank partial, all-reduce (ncclAllReduce on the GPU box)
        partial = torch.tensor([oracle.reduce_dot(y, y)], dtype=torch.float64)
        dist.all_reduce(partial)
        full = oracle.spmat_apply(part, part, frow, fcol, fval, x_full)
        ref = oracle.reduce_dot(full, full, kahan=True)
        assert abs(partial.item() - ref) <= 1e-10 * abs(ref)
        L.check(lib.vexb_halo_plan_destroy(plan))
    dist.barrier()
    dist.destroy_process_group()
    (Path(out_dir) / f"ok{rank}").write_text("ok")


@pytest.mark.parametrize("world", [2, 3])
def test_strip_wise_setup_and_exchange_over_gloo(built, tmp_path, world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, 24, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
