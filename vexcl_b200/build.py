"""Build the native pieces in-tree (no JIT cache: the .so files travel with the repo).

  vexcl_b200/libvexb200.so   CUDA kernels + C ABI   (nvcc, sm_100a only)
  oracle/liboracle.so        CPU restatement (gcc, OpenMP)   -- test infrastructure
  tests/cpp/bin/*            C++ front-end tests (g++), linked against libvexb200.so

Each target is rebuilt only when a source is newer than the output.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "vexcl_b200" / "csrc"
LIB = ROOT / "vexcl_b200" / "libvexb200.so"
ORACLE_SRC = ROOT / "oracle" / "oracle.c"
ORACLE_LIB = ROOT / "oracle" / "liboracle.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _stale(out: Path, srcs) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in srcs)


def _run(cmd, **kw):
    print("+", " ".join(str(c) for c in cmd), flush=True)
    subprocess.run([str(c) for c in cmd], check=True, **kw)


def build_lib(force: bool = False, verbose_ptxas: bool = False) -> Path:
    cus = sorted(CSRC.glob("*.cu"))
    deps = [p for p in CSRC.glob("*") if p.is_file()] + [ROOT / "include" / "vexb200.h"]
    if not force and not _stale(LIB, deps):
        return LIB          # e.g. on the GPU box: the .so travels with the repo, the object files do not
    objdir = CSRC / "obj"
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for cu in cus:
        obj = objdir / (cu.stem + ".o")
        objs.append(obj)
        if force or _stale(obj, deps):
            cmd = [_nvcc(), *NVCC_FLAGS, "-c", cu, "-o", obj, "-I", ROOT / "include"]
            if verbose_ptxas:
                cmd += ["-Xptxas", "-v"]
            print("+", " ".join(str(c) for c in cmd), flush=True)
            procs.append((cu, subprocess.Popen([str(c) for c in cmd])))
    for cu, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed on {cu}")
    if force or procs or _stale(LIB, objs):
        _run([_nvcc(), "-shared", "-o", LIB, *objs, "-cudart", "static", "-ldl", "-lpthread"])
    return LIB


def build_oracle(force: bool = False) -> Path:
    if force or _stale(ORACLE_LIB, [ORACLE_SRC]):
        _run(["gcc", "-O3", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=c11",
              "-Wall", "-o", ORACLE_LIB, ORACLE_SRC, "-lm"])
    return ORACLE_LIB


def build_cpp_tests(force: bool = False):
    src_dir = ROOT / "tests" / "cpp"
    bin_dir = src_dir / "bin"
    outs = []
    if not src_dir.exists():
        return outs
    headers = list((ROOT / "include").rglob("*.h*")) + list(src_dir.glob("*.hpp"))
    for cpp in sorted(src_dir.glob("*.cpp")) + sorted((ROOT / "examples").glob("*.cpp")):
        bin_dir.mkdir(exist_ok=True)
        out = bin_dir / cpp.stem
        outs.append(out)
        if force or _stale(out, [cpp, *headers, LIB]):
            _run(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unused-function", "-rdynamic", "-I", ROOT / "include", cpp, "-o", out,
                  "-L", LIB.parent, "-lvexb200", f"-Wl,-rpath,{LIB.parent}", "-Wl,-rpath,$ORIGIN/../../../vexcl_b200",
                  "-lpthread", "-ldl"])
    return outs


def build_all(force: bool = False):
    build_lib(force)
    build_oracle(force)
    build_cpp_tests(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
