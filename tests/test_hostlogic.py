"""CPU-side checks of the product: the C-ABI library loads without a GPU and exports everything
include/vexb200.h declares; its host-only entry points (partition, ghost columns, halo plan) are
bit-exact against the oracle's literal restatement of the reference; compute entry points fail
loudly when no device is present; the product never touches oracle/."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import oracle

ROOT = Path(__file__).resolve().parent.parent
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def L(built):
    from vexcl_b200 import _lib
    _lib.lib()
    return _lib


def test_library_exports_every_declared_symbol(L):
    lib = L.lib()
    declared = L.declared_symbols()
    assert len(declared) >= 60
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/vexb200.h but not exported"
    assert set(lib._signatures) == set(declared), "ctypes table and header disagree"
    assert lib.vexb_abi_version() == 1
    # ABI structs have the layout the header promises
    assert C.sizeof(L.Term) == 16 and C.sizeof(L.Instr) == 4 and C.sizeof(L.Expr) == 8 + 16 * 16 + 4 * 64


def test_header_opcode_table_matches_python(L):
    text = (ROOT / "include" / "vexb200.h").read_text()
    body = text[text.index("typedef enum {\n    VEXB_OP_TERM"):text.index("} vexb_opcode;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"VEXB_OP_([A-Z0-9]+)\b(?!_)", body)
    names = [n for n in names if n != "COUNT"]
    assert names == list(L._OPS)


def test_partition_bit_exact(L):
    import vexcl_b200 as vx
    for n in (0, 1, 15, 16, 17, 1000, 4097, 10**6 + 3, 9998244, 16777216, 134217728):
        for nd in (1, 2, 3, 4, 7, 8):
            assert np.array_equal(vx.partition(n, nd), oracle.partition(n, nd))
    w = [0.5, 2.0, 1.25]
    assert np.array_equal(vx.partition(12345, 3, w), oracle.partition(12345, 3, w))
    part = (C.c_size_t * 2)()
    assert L.lib().vexb_partition(10, 0, None, part) == 2
    assert b"bad arguments" in L.lib().vexb_last_error()


def _plan(L, nd, col_part, ghosts):
    off = np.zeros(nd + 1, np.uint64)
    for d in range(nd):
        off[d + 1] = off[d] + len(ghosts[d])
    cat = np.ascontiguousarray(np.concatenate(ghosts) if off[-1] else np.empty(0, np.int64), dtype=np.int64)
    cp = (C.c_size_t * (nd + 1))(*[int(x) for x in col_part])
    go = (C.c_size_t * (nd + 1))(*[int(x) for x in off])
    plan = C.c_void_p()
    L.check(L.lib().vexb_halo_plan_create(nd, cp, cat.ctypes.data, go, C.byref(plan)))
    return plan


def _ghosts(L, nd, part, col_part, row, col, rb=8, cb=8):
    out = []
    lib = L.lib()
    for d in range(nd):
        prow = np.ascontiguousarray(row[part[d]:part[d + 1] + 1])
        pcol = np.ascontiguousarray(col[row[part[d]]:])
        cnt = C.c_size_t(0)
        L.check(lib.vexb_strip_ghost_cols(int(part[d + 1] - part[d]), prow.ctypes.data, rb, pcol.ctypes.data, cb,
                                          int(col_part[d]), int(col_part[d + 1]), None, C.byref(cnt)))
        g = np.empty(cnt.value, np.int64)
        cap = C.c_size_t(cnt.value)
        L.check(lib.vexb_strip_ghost_cols(int(part[d + 1] - part[d]), prow.ctypes.data, rb, pcol.ctypes.data, cb,
                                          int(col_part[d]), int(col_part[d + 1]), g.ctypes.data, C.byref(cap)))
        out.append(g)
    return out


@pytest.mark.parametrize("nd", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("shape", [(600, 600), (500, 900), (64, 4000)])
def test_halo_plan_tables_bit_exact(L, nd, shape):
    """setup_exchange (spmat.hpp:291-378): ghost sets, cols_to_send, cidx, cols_to_recv."""
    n, m = shape
    row, col, val = oracle.random_matrix(n, m, 9, seed=n + m + nd)
    part, cpart = oracle.partition(n, nd), oracle.partition(m, nd)
    ex = oracle.setup_exchange(part, cpart, row, col)
    ghosts = _ghosts(L, nd, part, cpart, row, col)
    for d in range(nd):
        assert np.array_equal(ghosts[d], ex["ghost"][d] if nd > 1 else np.empty(0, np.int64))
    plan = _plan(L, nd, cpart, ghosts)
    lib = L.lib()
    tot = C.c_size_t()
    L.check(lib.vexb_halo_plan_ref_sizes(plan, C.byref(tot)))
    assert tot.value == ex["cols_to_send"].size
    cts = np.empty(tot.value, np.int64)
    cidx = (C.c_size_t * (nd + 1))()
    L.check(lib.vexb_halo_plan_ref_tables(plan, cts.ctypes.data, cidx))
    assert np.array_equal(cts, ex["cols_to_send"])
    assert list(cidx) == list(ex["cidx"])
    for d in range(nd):
        rc = np.empty(len(ghosts[d]), np.int64)
        if rc.size:
            L.check(lib.vexb_halo_plan_ref_recv(plan, d, rc.ctypes.data))
        assert np.array_equal(rc, ex["cols_to_recv"][d])
    # pairwise form: what d receives from o is the run of d's ghosts owned by o; o sends exactly those
    send_cols, send_counts, recv_counts = [], [], []
    for d in range(nd):
        sc, rc_ = (C.c_size_t * nd)(), (C.c_size_t * nd)()
        L.check(lib.vexb_halo_plan_counts(plan, d, sc, rc_))
        send_counts.append(list(sc)); recv_counts.append(list(rc_))
        buf = np.empty(sum(sc), np.int64)
        if buf.size:
            L.check(lib.vexb_halo_plan_send_cols(plan, d, buf.ctypes.data))
        send_cols.append(buf)
    for d in range(nd):
        assert sum(recv_counts[d]) == len(ghosts[d]) and recv_counts[d][d] == 0
        pos = 0
        for o in range(nd):
            assert recv_counts[d][o] == send_counts[o][d]
            seg = ghosts[d][pos:pos + recv_counts[d][o]]
            assert np.all((seg >= cpart[o]) & (seg < cpart[o + 1]))
            so = sum(send_counts[o][:d])
            assert np.array_equal(send_cols[o][so:so + send_counts[o][d]] + cpart[o], seg)
            pos += recv_counts[d][o]
    L.check(lib.vexb_halo_plan_destroy(plan))


def test_halo_plan_against_golden_fixture(L):
    g = np.load(GOLD / "exchange_600x3.npz")
    part = g["part"]
    ghosts = _ghosts(L, 3, part, part, g["row"], g["col"])
    for d in range(3):
        assert np.array_equal(ghosts[d], g[f"ghost{d}"])
    plan = _plan(L, 3, part, ghosts)
    cts = np.empty(g["cols_to_send"].size, np.int64)
    cidx = (C.c_size_t * 4)()
    L.check(L.lib().vexb_halo_plan_ref_tables(plan, cts.ctypes.data, cidx))
    assert np.array_equal(cts, g["cols_to_send"]) and list(cidx) == list(g["cidx"])
    L.check(L.lib().vexb_halo_plan_destroy(plan))


def test_ghost_cols_index_widths_and_errors(L):
    row, col, val = oracle.random_matrix(200, 300, 7, seed=9)
    part, cpart = oracle.partition(200, 2), oracle.partition(300, 2)
    a = _ghosts(L, 2, part, cpart, row, col)
    b = _ghosts(L, 2, part, cpart, row.astype(np.uint32), col.astype(np.int32), 4, 4)
    for d in range(2):
        assert np.array_equal(a[d], b[d])
    lib = L.lib()
    cnt = C.c_size_t(0)
    assert lib.vexb_strip_ghost_cols(10, row.ctypes.data, 3, col.ctypes.data, 8, 0, 10, None, C.byref(cnt)) == 2
    # a ghost list that contains a local column is rejected
    bad = [np.array([int(cpart[0])], np.int64), np.empty(0, np.int64)]
    off = (C.c_size_t * 3)(0, 1, 1)
    cp = (C.c_size_t * 3)(*[int(x) for x in cpart])
    plan = C.c_void_p()
    assert lib.vexb_halo_plan_create(2, cp, bad[0].ctypes.data, off, C.byref(plan)) == 2
    assert b"not a sorted set of remote columns" in lib.vexb_last_error()


def test_generators_match_oracle():
    from vexcl_b200 import gen
    for dim, n in ((2, 37), (3, 11)):
        r, c, v = oracle.poisson(dim, n)
        r2, c2, v2 = gen.poisson_strip(dim, n)
        assert np.array_equal(r, r2) and np.array_equal(c, c2) and np.array_equal(v, v2)
        N = r.size - 1
        a, b = N // 3, 2 * N // 3 + 5
        rs, cs, vs = gen.poisson_strip(dim, n, r0=a, r1=b)
        assert np.array_equal(rs, r[a:b + 1] - r[a]) and np.array_equal(cs, c[r[a]:r[b]]) and np.array_equal(vs, v[r[a]:r[b]])
        assert gen.poisson_nnz(dim, n) == (N, int(r[-1]))
    assert gen.spmv_bytes(9998244, 9998244, 49940644) == 799252612
    assert gen.spmv_bytes(16777216, 16777216, 115099600) == 1716739524
    assert gen.spmv_bytes(134217728, 134217728, 930123728) == 13845839300


def test_python_lowering_and_eval_path_without_gpu(L):
    """Expression lowering (type promotion, CVT insertion) and shape recognition are host logic."""
    import vexcl_b200 as vx
    from vexcl_b200 import api

    class FakeCtx:
        nparts, local, devs, streams, weights = 1, [0], {0: 0}, {0: None}, None
        def partition(self, n): return vx.partition(n, 1)

    def fake_vec(n, dt, addr):
        v = api.vector.__new__(api.vector)
        v.ctx, v.n, v.np_dtype, v.dtype, v.part, v.bufs = FakeCtx(), n, np.dtype(dt), api._vdt(dt), vx.partition(n, 1), {0: C.c_void_p(addr)}
        return v

    a, b, c, d = (fake_vec(1024, np.float64, 0x1000 * (i + 1)) for i in range(4))
    assert a.eval_path(L.SET, b + c * d) == "sweep:muladd"
    assert a.eval_path(L.ADD, c * d + b) == "sweep:muladd"
    assert a.eval_path(L.SET, 0.5 * a + b) == "sweep:axpy"
    assert a.eval_path(L.SET, 2 * a + b) == "sweep:axpy"            # int scalar folded to double on the host
    assert a.eval_path(L.SET, b - 0.25 * c) == "sweep:xmay"
    assert a.eval_path(L.SET, 3.0) == "sweep:fill"
    assert a.eval_path(L.SET, b * b) == "sweep:sqr"
    assert a.eval_path(L.MUL, b) == "interp"
    assert a.eval_path(L.SET, vx.sin(b) + c) == "interp"
    f = fake_vec(1024, np.float32, 0x9000)
    assert a.eval_path(L.SET, f + b) == "interp"                     # mixed element types
    assert f.eval_path(L.SET, f * f) == "sweep:sqr"
    low = api._Lowering(0, 0)
    low.size = 1024
    i = fake_vec(1024, np.int32, 0xa000)
    low.lower(i * f + b)
    ops = [(L._OPS[low.e.code[k].op], low.e.code[k].type) for k in range(low.e.n_code)]
    assert ops == [("TERM", L.I32), ("CVT", L.F32), ("TERM", L.F32), ("MUL", L.F32), ("CVT", L.F64), ("TERM", L.F64), ("ADD", L.F64)]
    with pytest.raises(ValueError):
        a.eval_path(L.SET, b + fake_vec(1000, np.float64, 0xb000))


def test_compute_entry_points_fail_loudly_without_a_device(L):
    """No CPU fallback: on a box without a GPU, vexb_init and compute calls return an error code."""
    lib = L.lib()
    n = C.c_int(-1)
    rc = lib.vexb_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    assert lib.vexb_init() != 0
    assert b"cuda" in lib.vexb_last_error().lower()
    import vexcl_b200 as vx
    with pytest.raises(vx.VexbError):
        vx.Context([0])


def test_product_never_uses_the_oracle():
    """oracle/ is test infrastructure: nothing under vexcl_b200/ or include/ may import, link or call it."""
    pat = re.compile(r"\boracle\b|liboracle|orc_[a-z_]+\s*\(")
    offenders = []
    for base in ("vexcl_b200", "include"):
        for p in (ROOT / base).rglob("*"):
            if p.is_file() and p.suffix in {".py", ".cu", ".cuh", ".hpp", ".h", ".cpp"}:
                for ln, line in enumerate(p.read_text(errors="replace").splitlines(), 1):
                    line = re.split(r"//|#", line)[0]                      # comments may mention the checker
                    if pat.search(line) and "build_oracle" not in line and "ORACLE_" not in line and p.name != "build.py":
                        offenders.append(f"{p.relative_to(ROOT)}:{ln}: {line.strip()}")
    assert not offenders, "\n".join(offenders)


def test_sliced_ell_layout_host_only(built):
    """VEXB_FMT_SELL (csrc/spmv.cu): windows of sigma rows sorted by length (longest first, ties in row order), slices of 32
    lanes as wide as their longest row.  Host logic: checked here without a GPU against a numpy restatement."""
    import ctypes as C
    from vexcl_b200 import _lib as L
    lib = L.lib()
    rng = np.random.default_rng(11)
    for n, sigma in ((5000, 1024), (1000, 32), (33, 64), (1, 32), (4096, 4096), (0, 1024)):
        w = rng.integers(0, 32, n)
        if n > 7:
            w[7] = 300                                                  # one long row
        row = np.concatenate([[0], np.cumsum(w)]).astype(np.int64)
        ns, slots = C.c_size_t(0), C.c_size_t(0)
        L.check(lib.vexb_csr_sell_layout(n, row.ctypes.data, 8, sigma, C.byref(ns), C.byref(slots), None, None))
        assert ns.value == (n + 31) // 32
        perm = np.full(ns.value * 32, -7, np.int32)
        sptr = np.full(ns.value + 1, -7, np.int32)
        L.check(lib.vexb_csr_sell_layout(n, row.ctypes.data, 8, sigma, C.byref(ns), C.byref(slots), perm.ctypes.data, sptr.ctypes.data))
        # numpy restatement
        want = np.full(ns.value * 32, -1, np.int64)
        for w0 in range(0, n, sigma):
            idx = np.arange(w0, min(n, w0 + sigma))
            want[w0:w0 + idx.size] = idx[np.argsort(-w[idx], kind="stable")]
        assert np.array_equal(perm, want)
        assert sorted(perm[perm >= 0]) == list(range(n))                # a permutation of the rows
        widths = np.array([max([w[r] for r in perm[s * 32:(s + 1) * 32] if r >= 0], default=0) for s in range(ns.value)], dtype=np.int64)
        assert np.array_equal(sptr, np.concatenate([[0], np.cumsum(widths * 32)]))
        assert slots.value == int(widths.sum() * 32)
        if n == 5000:                                                   # sorting keeps the padding small: < 10 % on U[0,32) + one long row
            assert slots.value <= 1.10 * row[-1] + 300 * 32
    # 32-bit row pointers, decreasing pointers rejected
    row32 = np.array([0, 3, 5, 9], np.int32)
    L.check(lib.vexb_csr_sell_layout(3, row32.ctypes.data, 4, 32, C.byref(ns), C.byref(slots), None, None))
    assert (ns.value, slots.value) == (1, 4 * 32)
    bad = np.array([0, 3, 2], np.int64)
    assert lib.vexb_csr_sell_layout(2, bad.ctypes.data, 8, 32, C.byref(ns), C.byref(slots), None, None) == 2
