"""Regenerate tests/golden/*.npz.

The reference holds no stored vectors (its tests seed from time(0), tests/context_setup.hpp:19-22)
and cannot be built in this image, so these fixtures are produced by the oracle AFTER it has been
pinned to the reference's closed-form known answers (tests/test_oracle_kat.py).  They freeze the
oracle's outputs on fixed seeds so that GPU runs are compared against files, and so that a later
change to oracle/ that alters any result is caught.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    # config 1 plumbing case, reduced to 4096 elements: a = b + c*d, seeds 42..44
    n = 4096
    b, c, d = (oracle.uniform_real(s, n) for s in (42, 43, 44))
    np.savez_compressed(OUT / "axpy_seed42.npz", b=b, c=c, d=d, a=oracle.vec_muladd(np.zeros(n), b, c, d),
                        a_acc=oracle.vec_muladd(b, b, c, d, accumulate=True),
                        dot_bc=np.array([oracle.reduce_dot(b, c, kahan=True)]), min_b=np.array([b.min()]), max_b=np.array([b.max()]))
    # SpMV: 2-D Poisson 48^2 and 3-D Poisson 12^3 with x == 1e-2 (the benchmark's KAT) and U[0,1) seed 7
    for dim, m in ((2, 48), (3, 12)):
        row, col, val = oracle.poisson(dim, m)
        N = row.size - 1
        xr = oracle.uniform_real(7, N)
        np.savez_compressed(OUT / f"poisson{dim}d_{m}.npz", row=row, col=col, val=val, x_rand=xr,
                            y_const=oracle.csr_spmv(row, col, val, np.full(N, 1e-2)), y_rand=oracle.csr_spmv(row, col, val, xr))
    # multi-device tables for a random 600 x 600 matrix on 3 parts
    row, col, val = oracle.random_matrix(600, 600, 8, seed=2024)
    part = oracle.partition(600, 3)
    ex = oracle.setup_exchange(part, part, row, col)
    np.savez_compressed(OUT / "exchange_600x3.npz", row=row, col=col, val=val, part=part, cols_to_send=ex["cols_to_send"],
                        cidx=ex["cidx"], recv0=ex["cols_to_recv"][0], recv1=ex["cols_to_recv"][1], recv2=ex["cols_to_recv"][2],
                        ghost0=ex["ghost"][0], ghost1=ex["ghost"][1], ghost2=ex["ghost"][2])
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))


if __name__ == "__main__":
    main()
