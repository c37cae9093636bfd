"""Runs the C++ front-end tests (tests/cpp/*.cpp: the reference's own hot-path test cases compiled
against include/vexcl, no Boost) on the GPU.  Each binary is run twice: with the reference fixture's
duplicated queue (two slices on one device) and with a single slice."""
import os
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
BIN = Path(__file__).resolve().parent / "cpp" / "bin"


@pytest.mark.parametrize("parts", ["2", "1"])
@pytest.mark.parametrize("name", ["test_vector_arithmetics", "test_spmv", "test_sparse_matrices", "test_multivector"])
def test_cpp_front_end(built, name, parts):
    from vexcl_b200 import build
    build.build_cpp_tests()
    exe = BIN / name
    assert exe.exists(), f"{exe} was not built"
    env = dict(os.environ, VEXCL_TEST_PARTS=parts)
    r = subprocess.run([str(exe), "12345"], capture_output=True, text=True, env=env, timeout=300)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0, f"{name} failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
    assert " 0 failures" in r.stdout
