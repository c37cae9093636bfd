"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of vex::stencil convolution.

Reference semantics: the generated kernel of vexcl/stencil.hpp:232-330 (local-memory window with clamped ends) and, in
plain form, the CPU check of examples/benchmark.cpp:318-327 and tests/stencil.cpp:8-16,36-45:

    for i in range(n):
        sum = 0
        for k in range(len(s)):
            sum += s[k] * x[min(n - 1, max(0, i + k - center))]
        y[i] = sum

numpy, vectorised over i; the loop over taps stays sequential and every product and sum is rounded separately, like the
scalar loop compiled without FMA contraction.  Pinned by tests/test_stencil_oracle.py against a literal transcription
of the loop and the closed-form cases of the reference tests (constant vector -> sum of taps; two_stencils -> zeros).
"""
import numpy as np


def convolve(s, center, x, y=None, alpha=1.0, append=False):
    s = np.asarray(s)
    x = np.asarray(x)
    n = x.size
    i = np.arange(n, dtype=np.int64)
    acc = np.zeros(n, dtype=x.dtype)
    for k in range(s.size):
        acc = acc + s[k] * x[np.clip(i + (k - center), 0, n - 1)]
    out = x.dtype.type(alpha) * acc
    return np.asarray(y) + out if append else out


def convolve_loop(s, center, x):
    """The reference loop, literally (small cases only)."""
    n = len(x)
    y = np.zeros(n, dtype=np.asarray(x).dtype)
    for i in range(n):
        acc = y.dtype.type(0)
        for k in range(len(s)):
            acc = acc + s[k] * x[min(n - 1, max(0, i + k - center))]
        y[i] = acc
    return y
