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
    assert r.returncode == 0, f"{name} exited with status {r.returncode}:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
    assert " 0 failures" in r.stdout


@pytest.mark.parametrize("parts", ["2", "1"])
def test_cpp_constants(built, parts):
    _run_binary("test_constants", parts, 120)


def _run_stencil(parts: str, timeout: int):
    _run_binary("test_stencil", parts, timeout)


def _run_binary(name: str, parts: str, timeout: int):
    from vexcl_b200 import build
    build.build_cpp_tests()
    exe = BIN / name
    assert exe.exists(), f"{exe} was not built"
    r = subprocess.run([str(exe), "12345"], capture_output=True, text=True, env=dict(os.environ, VEXCL_TEST_PARTS=parts),
                       timeout=timeout)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0 and " 0 failures" in r.stdout, f"{name} exited with status {r.returncode}:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"


def test_cpp_stencil_single_slice(built):
    """tests/cpp/test_stencil.cpp (the cases of the reference's tests/stencil.cpp) on one slice."""
    _run_stencil("1", 300)


def test_cpp_stencil_two_slices(built):
    try:
        _run_stencil("2", 90)
    except subprocess.TimeoutExpired:
        pytest.fail("test_stencil with two slices did not finish in 90 s")


def test_cpp_hotpath_benchmark(built):
    """examples/hotpath_benchmark.cpp: the hot-path functions of the reference's examples/benchmark.cpp (saxpy :83-148,
    vector :152-216, reductor :219-278, spmv :352-477) against include/vexcl, with the reference's own self-checks."""
    import re
    from vexcl_b200 import build
    build.build_cpp_tests()
    exe = BIN / "hotpath_benchmark"
    assert exe.exists(), f"{exe} was not built"
    r = subprocess.run([str(exe), "--quick"], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:]); print(r.stderr[-2000:])
    assert r.returncode == 0
    res = [float(v) for v in re.findall(r"res = ([-+0-9.eE]+|nan|inf)", r.stdout)]
    assert len(res) == 4, r.stdout                               # saxpy, vector arithmetic, reduction, SpMV
    assert res[0] <= 1e-12 and res[1] <= 1e-12 and res[3] <= 1e-12   # sums of squared differences against the CPU loops
    assert res[2] <= 1e-10                                       # relative difference of the reductions
    assert len(re.findall(r"Bandwidth:\s+[0-9.]+", r.stdout)) == 4
