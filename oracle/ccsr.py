"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's CCSR product,
vexcl/spmat/ccsr.hpp:41-47 (the loop in the class comment) and :191-199 (the generated device function):

    for i in range(n):
        sum = 0
        for j in range(row[idx[i]], row[idx[i] + 1]):
            sum += val[j] * x[i + col[j]]
        y[i] = sum

numpy, vectorised over the rows that share a unique row; the accumulation over j stays sequential and every product
and sum is rounded separately, exactly like the scalar loop compiled without FMA contraction.

Pinned by tests/test_oracle_kat.py::test_ccsr_*: on the reference's own CCSR test matrix (tests/spmv.cpp:150-197, 3-D
Poisson n=32) it must agree bit for bit with the CSR oracle applied to the expanded matrix, and with a pure-Python
transcription of the loop above.
"""
import numpy as np


def ccsr_spmv(n, idx, row, col, val, x, y=None, alpha=1.0, append=False):
    idx = np.asarray(idx).astype(np.int64)
    row = np.asarray(row).astype(np.int64)
    col = np.asarray(col).astype(np.int64)
    val = np.asarray(val)
    x = np.asarray(x)
    out = np.zeros(n, dtype=val.dtype)
    for u in range(row.size - 1):
        rows = np.nonzero(idx == u)[0]
        if rows.size == 0:
            continue
        s = np.zeros(rows.size, dtype=val.dtype)
        for j in range(int(row[u]), int(row[u + 1])):
            s = s + val[j] * x[rows + col[j]]
        out[rows] = s
    out = val.dtype.type(alpha) * out
    if append:
        return np.asarray(y) + out
    return out


def ccsr_spmv_loop(n, idx, row, col, val, x):
    """The reference loop, literally (small cases only)."""
    y = np.zeros(n, dtype=np.asarray(val).dtype)
    for i in range(n):
        s = y.dtype.type(0)
        for j in range(int(row[int(idx[i])]), int(row[int(idx[i]) + 1])):
            s = s + val[j] * x[i + int(col[j])]
        y[i] = s
    return y


def ccsr_to_csr(n, idx, row, col, val):
    """Expand to plain CSR with absolute columns (for cross-checks against the CSR oracle)."""
    idx = np.asarray(idx).astype(np.int64)
    row = np.asarray(row).astype(np.int64)
    width = (row[1:] - row[:-1])[idx]
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(width, out=ptr[1:])
    c = np.empty(int(ptr[-1]), dtype=np.int64)
    v = np.empty(int(ptr[-1]), dtype=np.asarray(val).dtype)
    for u in range(row.size - 1):
        rows = np.nonzero(idx == u)[0]
        for t, j in enumerate(range(int(row[u]), int(row[u + 1]))):
            c[ptr[rows] + t] = rows + int(col[j])
            v[ptr[rows] + t] = val[j]
    return ptr, c, v
