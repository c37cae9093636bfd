"""Pins oracle/stencil.py on CPU: literal loop (benchmark.cpp:318-327), closed forms of tests/stencil.cpp."""
import numpy as np

import oracle
from oracle import stencil


def test_stencil_oracle_against_the_literal_loop():
    rng = np.random.default_rng(11)
    for n, w in ((1, 1), (5, 9), (128, 64), (300, 21), (64, 1)):
        for center in sorted({0, w // 2, w - 1}):
            s, x = rng.random(w), rng.random(n)
            assert np.array_equal(stencil.convolve(s, center, x), stencil.convolve_loop(s, center, x))
    y0 = rng.random(300)
    s, x = rng.random(21), rng.random(300)
    assert np.array_equal(stencil.convolve(s, 10, x, y=y0, alpha=42.0, append=True), y0 + 42.0 * stencil.convolve(s, 10, x))


def test_stencil_oracle_closed_forms():
    # a constant vector is reproduced times the sum of the taps (accumulated in tap order)
    s = oracle.uniform_real(3, 21)
    want = 0.0
    for v in s:
        want = want + v * 2.0
    assert np.all(stencil.convolve(s, 10, np.full(50, 2.0)) == want)
    # tests/stencil.cpp:58-75 (two_stencils): zeros in, zeros out
    assert np.all(stencil.convolve(np.ones(5), 3, np.zeros(32)) == 0)
    # a one-tap stencil with center 0 is a scaling; a shift stencil clamps at the ends
    x = np.arange(10.0)
    assert np.array_equal(stencil.convolve([3.0], 0, x), 3 * x)
    assert np.array_equal(stencil.convolve([1.0, 0.0], 1, x), np.concatenate([[0.0], x[:-1]]))   # y[i] = x[i-1], clamped
    assert np.array_equal(stencil.convolve([0.0, 0.0, 1.0], 0, x), np.concatenate([x[2:], [9.0, 9.0]]))
