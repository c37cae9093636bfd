"""CPU oracle for the VexCL hot paths -- TEST INFRASTRUCTURE, never part of the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference``
legs may import this package.  vexcl_b200/ never does (tests/test_no_oracle_in_product.py
checks the sources and the library's dynamic dependencies).

Two halves:
  * oracle.c  (liboracle.so, gcc + OpenMP): floating-point loops -- elementwise chunks,
    Reductor work-group model, CSR / hybrid-ELL products, generators, RNG, partition.
  * this file (numpy / plain Python): the *index tables* of the multi-device SpMat,
    written the way the reference writes them (std::set -> Python set / np.unique),
    small sizes only.

Reference citations are relative to /root/reference.  The reference cannot be built in
this image (no Boost, no OpenCL headers), and it stores no golden vectors; the oracle is
pinned against the closed-form known answers of the reference's own tests in
tests/test_oracle_kat.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"
_lib = None

SUM, SUM_KAHAN, MAX, MIN = 0, 1, 2, 3


def build(force: bool = False) -> Path:
    src = _HERE / "oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["gcc", "-O3", "-march=native", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
                        "-std=c11", "-Wall", "-o", str(_LIB_PATH), str(src), "-lm"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        dp, ip, sz = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_size_t
        L.orc_num_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_set_threads(usable_cpus())      # respect the container's CPU quota (cgroup cpu.max), not just the CPU count
        L.orc_partition.argtypes = [sz, C.c_int, dp, C.POINTER(sz)]
        L.orc_uniform_real.argtypes = [C.c_uint32, sz, dp]
        L.orc_poisson_sizes.argtypes = [C.c_int, sz, C.POINTER(sz), C.POINTER(sz)]
        L.orc_poisson.argtypes = [C.c_int, sz, ip, ip, dp]
        L.orc_vec_muladd.argtypes = [dp, dp, dp, dp, sz, C.c_int, C.c_int, C.c_int]
        L.orc_vec_saxpy.argtypes = [dp, C.c_double, dp, sz, C.c_int, C.c_int]
        L.orc_reduce.argtypes = [dp, sz, C.c_int, C.c_int]
        L.orc_reduce.restype = C.c_double
        L.orc_reduce_dot.argtypes = [dp, dp, sz, C.c_int, C.c_int]
        L.orc_reduce_dot.restype = C.c_double
        L.orc_kahan_sum.argtypes = [dp, sz]
        L.orc_kahan_sum.restype = C.c_double
        L.orc_csr_spmv.argtypes = [sz, ip, ip, dp, dp, dp, C.c_double, C.c_int, C.c_int, C.c_int]
        L.orc_csr_absrow.argtypes = [sz, ip, ip, dp, dp, dp]
        L.orc_hell_spmv.argtypes = [sz, sz, sz, ip, dp, ip, ip, dp, dp, dp, C.c_double, C.c_int]
        L.orc_hell_width.argtypes = [ip, sz]
        L.orc_hell_width.restype = sz
        L.orc_cpp_vec.argtypes = [dp, dp, dp, dp, sz]
        L.orc_cpp_dot.argtypes = [dp, dp, sz]
        L.orc_cpp_dot.restype = C.c_double
        L.orc_cpp_spmv.argtypes = [sz, ip, ip, dp, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def usable_cpus() -> int:
    """Host threads this process can really run: the smaller of the visible CPUs (affinity mask) and the
    cgroup CPU quota (cpu.max = "quota period"), e.g. 16 on a 128-thread host leased with a 16-CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = Path(path).read_text().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, -(-int(txt[0]) // int(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    n = min(n, max(1, -(-q // int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text()))))
            break
        except (OSError, ValueError, IndexError):
            continue
    return n


def num_threads() -> int:
    return lib().orc_num_threads()


def default_groups() -> int:
    """8 * compute units (vexcl/backend/opencl/kernel.hpp:166-171); a CPU device exposes one
    compute unit per usable hardware thread."""
    return 8 * usable_cpus()


# ---------------------------------------------------------------------------- partition
def partition(n: int, nparts: int, weights=None) -> np.ndarray:
    """vexcl/vector.hpp:131-167 (alignup 16: util.hpp:91-93).  weights=None -> equal_weights."""
    part = (C.c_size_t * (nparts + 1))()
    w = None if weights is None else _dp(_f64(weights))
    lib().orc_partition(n, nparts, w, part)
    return np.array(list(part), dtype=np.int64)


# ---------------------------------------------------------------------------- inputs
def uniform_real(seed: int, n: int) -> np.ndarray:
    """std::default_random_engine(seed) + uniform_real_distribution<double>(0,1) (libstdc++),
    the generator of tests/random_vector.hpp:12-19 and examples/benchmark.cpp:72-80."""
    out = np.empty(n, dtype=np.float64)
    lib().orc_uniform_real(seed & 0xFFFFFFFF, n, _dp(out))
    return out


def poisson(dim: int, n: int):
    """examples/benchmark.cpp:357-415 (dim=3) or its 2-D 5-point analogue.  Returns int64 row, col and f64 val."""
    nrows, nnz = C.c_size_t(), C.c_size_t()
    lib().orc_poisson_sizes(dim, n, C.byref(nrows), C.byref(nnz))
    row = np.empty(nrows.value + 1, dtype=np.int64)
    col = np.empty(nnz.value, dtype=np.int64)
    val = np.empty(nnz.value, dtype=np.float64)
    lib().orc_poisson(dim, n, _ip(row), _ip(col), _dp(val))
    return row, col, val


def tridiagonal(n: int):
    """-1 / 2 / -1 matrix of tests/sparse_matrices.cpp:155-175."""
    row = [0]; col = []; val = []
    for i in range(n):
        if i > 0:
            col.append(i - 1); val.append(-1.0)
        col.append(i); val.append(2.0)
        if i + 1 < n:
            col.append(i + 1); val.append(-1.0)
        row.append(len(col))
    return np.array(row, np.int64), np.array(col, np.int64), np.array(val, np.float64)


def random_matrix(n: int, m: int, nnz_per_row: int, seed: int):
    """Shape of tests/random_matrix.hpp:9-40: row width U[0, nnz_per_row-1], sorted unique random
    columns, U[0,1) values.  (The reference seeds from time(0); a fixed numpy stream is used here.)"""
    rng = np.random.default_rng(seed)
    row = [0]; cols = []
    for _ in range(n):
        w = int(rng.integers(0, nnz_per_row))
        cs = np.sort(rng.choice(m, size=min(w, m), replace=False)) if w else np.empty(0, np.int64)
        cols.append(cs.astype(np.int64))
        row.append(row[-1] + len(cs))
    col = np.concatenate(cols) if cols else np.empty(0, np.int64)
    val = rng.random(len(col))
    return np.array(row, np.int64), col, val


# ---------------------------------------------------------------------------- elementwise
def vec_muladd(a, b, c, d, accumulate=False, fma=False, groups=None) -> np.ndarray:
    """a = b + c*d  or  a += b + c*d, chunked work-group model."""
    a = _f64(a).copy(); b, c, d = _f64(b), _f64(c), _f64(d)
    lib().orc_vec_muladd(_dp(a), _dp(b), _dp(c), _dp(d), a.size, int(accumulate), int(fma), groups or default_groups())
    return a


def vec_saxpy(a, alpha, b, fma=False, groups=None) -> np.ndarray:
    a = _f64(a).copy(); b = _f64(b)
    lib().orc_vec_saxpy(_dp(a), float(alpha), _dp(b), a.size, int(fma), groups or default_groups())
    return a


# ---------------------------------------------------------------------------- reductions
def reduce(x, op=SUM, groups=None) -> float:
    """vexcl/reductor.hpp:302-439 on a CPU device; x is the evaluated expression."""
    x = _f64(x)
    return lib().orc_reduce(_dp(x), x.size, op, groups or default_groups())


def reduce_dot(a, b, kahan=False, groups=None) -> float:
    a, b = _f64(a), _f64(b)
    return lib().orc_reduce_dot(_dp(a), _dp(b), a.size, int(kahan), groups or default_groups())


def kahan_sum(x) -> float:
    x = _f64(x)
    return lib().orc_kahan_sum(_dp(x), x.size)


# ---------------------------------------------------------------------------- SpMV
def csr_spmv(row, col, val, x, y=None, alpha=1.0, append=False, fma=False, groups=None) -> np.ndarray:
    """vexcl/spmat/csr.inl:163-170.  Returns y (new array)."""
    row, col, val, x = _i64(row), _i64(col), _f64(val), _f64(x)
    n = row.size - 1
    y = np.zeros(n) if y is None else _f64(y).copy()
    lib().orc_csr_spmv(n, _ip(row), _ip(col), _dp(val), _dp(x), _dp(y), float(alpha), int(append), int(fma),
                       groups or default_groups())
    return y


def csr_absrow(row, col, val, x) -> np.ndarray:
    row, col, val, x = _i64(row), _i64(col), _f64(val), _f64(x)
    out = np.empty(row.size - 1)
    lib().orc_csr_absrow(row.size - 1, _ip(row), _ip(col), _dp(val), _dp(x), _dp(out))
    return out


def hell_width(widths) -> int:
    w = _i64(widths)
    return int(lib().orc_hell_width(_ip(w), w.size))


def hell_pack(row, col, val):
    """Pack one strip part into hybrid ELL exactly as vexcl/spmat/hybrid_ell.inl:132-193 does:
    column-major ELL with pitch alignup(n,16), sentinel -1, overflow to a CSR tail."""
    row, col, val = _i64(row), _i64(col), _f64(val)
    n = row.size - 1
    w = hell_width(np.diff(row))
    pitch = (n + 15) // 16 * 16
    ell_col = np.full(pitch * w, -1, np.int64)
    ell_val = np.zeros(pitch * w)
    t_row = [0]; t_col = []; t_val = []
    for i in range(n):
        cnt = 0
        for j in range(row[i], row[i + 1]):
            if cnt < w:
                ell_col[i + pitch * cnt] = col[j]; ell_val[i + pitch * cnt] = val[j]; cnt += 1
            else:
                t_col.append(col[j]); t_val.append(val[j])
        t_row.append(len(t_col))
    return dict(width=w, pitch=pitch, ell_col=ell_col, ell_val=ell_val,
                csr_row=np.array(t_row, np.int64), csr_col=np.array(t_col, np.int64), csr_val=np.array(t_val, np.float64))


def hell_spmv(h, x, y=None, alpha=1.0, append=False) -> np.ndarray:
    n = h["csr_row"].size - 1
    y = np.zeros(n) if y is None else _f64(y).copy()
    x = _f64(x)
    has_tail = h["csr_col"].size > 0
    lib().orc_hell_spmv(n, h["width"], h["pitch"], _ip(h["ell_col"]), _dp(h["ell_val"]),
                        _ip(h["csr_row"]) if has_tail else None, _ip(h["csr_col"]) if has_tail else None,
                        _dp(h["csr_val"]) if has_tail else None, _dp(x), _dp(y), float(alpha), int(append))
    return y


# ---------------------------------------------------------------------------- multi-device SpMat tables
def setup_exchange(part, col_part, row, col):
    """vexcl/spmat.hpp:291-378, literally: per-device ghost sets, their sorted union
    `cols_to_send`, owner offsets `cidx`, per-device `cols_to_recv` positions, and the
    owner-relative send lists."""
    nd = len(part) - 1
    ghost = []
    for d in range(nd):
        j0, j1 = row[part[d]], row[part[d + 1]]
        c = np.asarray(col[j0:j1])
        g = np.unique(c[(c < col_part[d]) | (c >= col_part[d + 1])]) if nd > 1 else np.empty(0, np.int64)
        ghost.append(g.astype(np.int64))
    cols_to_send = np.unique(np.concatenate(ghost)) if nd > 1 and sum(len(g) for g in ghost) else np.empty(0, np.int64)
    cols_to_recv = [np.searchsorted(cols_to_send, g).astype(np.int64) for g in ghost]
    cidx = np.searchsorted(cols_to_send, np.asarray(col_part, dtype=np.int64), side="left").astype(np.int64)
    send_local = cols_to_send.copy()
    for d in range(nd):
        send_local[cidx[d]:cidx[d + 1]] -= col_part[d]
    return dict(ghost=ghost, cols_to_send=send_local, cols_to_send_global=cols_to_send, cidx=cidx, cols_to_recv=cols_to_recv)


def split_strip(row, col, val, r0, r1, col_begin, col_end, ghost):
    """vexcl/spmat/csr.inl:70-112: local part with columns shifted by col_begin, remote part
    renumbered by rank in the sorted ghost set; storage order kept."""
    lrow = [0]; lcol = []; lval = []; rrow = [0]; rcol = []; rval = []
    r2l = {int(c): k for k, c in enumerate(ghost)}
    for i in range(r0, r1):
        for j in range(row[i], row[i + 1]):
            c = int(col[j])
            if col_begin <= c < col_end:
                lcol.append(c - col_begin); lval.append(val[j])
            else:
                rcol.append(r2l[c]); rval.append(val[j])
        lrow.append(len(lcol)); rrow.append(len(rcol))
    f = lambda a, t: np.array(a, dtype=t)
    return (f(lrow, np.int64), f(lcol, np.int64), f(lval, np.float64),
            f(rrow, np.int64), f(rcol, np.int64), f(rval, np.float64))


def spmat_apply(part, col_part, row, col, val, x, y=None, alpha=1.0, append=False):
    """SpMat::apply on nd devices (vexcl/spmat.hpp:120-185): gather vals_to_send, local product,
    host rx shuffle, remote product `+=`.  Returns the full y."""
    nd = len(part) - 1
    n = part[-1]
    x = _f64(x)
    y = np.zeros(n) if y is None else _f64(y).copy()
    ex = setup_exchange(part, col_part, row, col)
    rx = np.zeros(len(ex["cols_to_send"]))
    for d in range(nd):
        xs = x[col_part[d]:col_part[d + 1]]
        rx[ex["cidx"][d]:ex["cidx"][d + 1]] = xs[ex["cols_to_send"][ex["cidx"][d]:ex["cidx"][d + 1]]]
    for d in range(nd):
        if part[d + 1] == part[d]:
            continue
        lr, lc, lv, rr, rc, rv = split_strip(row, col, val, part[d], part[d + 1], col_part[d], col_part[d + 1], ex["ghost"][d])
        xs = x[col_part[d]:col_part[d + 1]]
        ys = y[part[d]:part[d + 1]]
        if lc.size or append:
            ys = csr_spmv(lr, lc, lv, xs, ys, alpha, append) if lc.size else ys
        else:
            ys = np.zeros_like(ys)                                   # csr.inl:195-200
        if rc.size:
            ys = csr_spmv(rr, rc, rv, rx[ex["cols_to_recv"][d]], ys, alpha, True)
        y[part[d]:part[d + 1]] = ys
    return y


# ---------------------------------------------------------------------------- CG (composition check)
def cg(row, col, val, b, x0, iters):
    """Conjugate gradient from the oracle's own pieces, the way vexcl/external/viennacl.hpp:36-64 lets
    ViennaCL compose it: prod -> csr_spmv, inner_prod -> reduce_dot, vector updates in plain numpy."""
    x = _f64(x0).copy()
    r = _f64(b) - csr_spmv(row, col, val, x)
    p = r.copy()
    rho = reduce_dot(r, r, kahan=True)
    hist = []
    for _ in range(iters):
        q = csr_spmv(row, col, val, p)
        alpha = rho / reduce_dot(p, q, kahan=True)
        x = x + alpha * p
        r = r - alpha * q
        rho_new = reduce_dot(r, r, kahan=True)
        p = r + (rho_new / rho) * p
        rho = rho_new
        hist.append(rho)
    return x, hist
