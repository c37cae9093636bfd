"""The oracle is test infrastructure: nothing the product ships may import, link or execute it.

Checked here: the Python package (everything except build.py, which only COMPILES the checker), the C++ front end, the
examples, the CUDA sources, and the dynamic dependencies of libvexb200.so."""
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _sources():
    for pattern in ("vexcl_b200/*.py", "vexcl_b200/csrc/*.cu", "vexcl_b200/csrc/*.cuh", "vexcl_b200/csrc/*.hpp",
                    "include/**/*.h", "include/**/*.hpp", "examples/*.cpp"):
        yield from ROOT.glob(pattern)


def test_product_sources_never_use_the_oracle():
    offenders = []
    for f in _sources():
        if f.name == "build.py":                       # compiles oracle/liboracle.so for the tests; never loads it
            continue
        text = f.read_text(errors="replace")
        # comments may cite the oracle as the checker of a kernel; code may not import, include, link or load it
        if re.search(r"^\s*(import|from)\s+oracle\b|liboracle|#\s*include[^\n]*oracle|dlopen\([^)]*oracle|CDLL\([^)]*oracle|oracle\.[a-z_]+\(", text, re.M):
            offenders.append(str(f.relative_to(ROOT)))
    assert not offenders, f"product sources refer to the oracle: {offenders}"


def test_build_script_only_compiles_the_oracle():
    text = (ROOT / "vexcl_b200" / "build.py").read_text()
    assert "import oracle" not in text and "from oracle" not in text and "CDLL" not in text


def test_library_does_not_link_the_oracle(built):
    lib = ROOT / "vexcl_b200" / "libvexb200.so"
    out = subprocess.run(["ldd", str(lib)], capture_output=True, text=True).stdout
    assert "oracle" not in out
    syms = subprocess.run(["nm", "-D", "--undefined-only", str(lib)], capture_output=True, text=True).stdout
    assert "oracle" not in syms
