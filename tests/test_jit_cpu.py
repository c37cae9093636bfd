"""The NVRTC side path, as far as it can be checked without a GPU: user functions are registered, the IR is
printed as CUDA C, and NVRTC compiles it for sm_100a (NVRTC needs no device).  Mirrors the user-function cases of
the reference's tests/vector_arithmetics.cpp:113-145 at the source level; the numerical checks are in
tests/test_gpu_user_functions.py."""
import ctypes as C

import numpy as np
import pytest

ROOTS = {}


@pytest.fixture(scope="module")
def env(built):
    import vexcl_b200 as vx
    from vexcl_b200 import api, _lib as L

    class FakeCtx:
        nparts, local, devs, streams, weights = 1, [0], {0: 0}, {0: None}, None
        def partition(self, n): return vx.partition(n, 1)

    def fake_vec(n, dt, addr):
        v = api.vector.__new__(api.vector)
        v.ctx, v.n, v.np_dtype, v.dtype, v.part, v.bufs = FakeCtx(), n, np.dtype(dt), api._vdt(dt), vx.partition(n, 1), {0: C.c_void_p(addr)}
        return v
    return vx, api, L, fake_vec


def jit_source(api, L, lhs, op, expr, compile=True):
    low = api._Lowering(0, 0)
    low.size = lhs.n
    low.lower(api.wrap(expr))
    n = C.c_size_t(0)
    L.check(L.lib().vexb_jit_source(lhs.dtype, op, C.byref(low.e), None, C.byref(n), 0))
    buf = C.create_string_buffer(n.value + 4096)
    cap = C.c_size_t(len(buf))
    L.check(L.lib().vexb_jit_source(lhs.dtype, op, C.byref(low.e), buf, C.byref(cap), int(compile)))
    return buf.value.decode()


def test_user_function_source_and_nvrtc_compile(env):
    vx, api, L, fake_vec = env
    x, y, z = (fake_vec(1024, np.float64, 0x1000 * (k + 1)) for k in range(3))
    greater = api.UserFunction(np.int32, "greater", [(np.float64, "x"), (np.float64, "y")], "return x > y;")
    times2 = api.UserFunction(np.float64, "times2", [(np.float64, "v")], "return v * 2;")
    assert api.UserFunction(np.float64, "times2", [(np.float64, "v")], "return v * 2;").id == times2.id      # same definition, same id
    assert z.eval_path(L.SET, times2(x) + y) == "jit"
    src = jit_source(api, L, z, L.SET, times2(x) + greater(x, y) * y)
    assert "__device__ __forceinline__ double times2_" in src and "const double v = prm1;" in src
    assert "__device__ __forceinline__ int greater_" in src
    assert "vexb_jit_kernel" in src and "NVRTC: ok" in src
    # compound assignment and mixed types go through the same printer
    i = fake_vec(1024, np.int32, 0x9000)
    src = jit_source(api, L, i, L.ADD, greater(x, 0.5) + (i << 2))
    assert "return (int)((int)lhs[i] + (int)" in src and "NVRTC: ok" in src
    assert "a3 = vexb_elem(tt, lhs, i + 768ull, off)" in src               # four elements per thread in flight, one contiguous chunk per block
    # every operator family compiles
    f = fake_vec(1024, np.float32, 0xa000)
    e = vx.if_else(x > y, vx.sin(x) * vx.pow_(y, 2.0), vx.fmin(x, y)) + vx.fma(x, y, z) - vx.fabs(-x) + f * i + vx.ElementIndex(3) % 7
    assert "NVRTC: ok" in jit_source(api, L, z, L.SET, e)
    u = fake_vec(1024, np.uint64, 0xb000)
    assert "NVRTC: ok" in jit_source(api, L, u, L.XOR, (u >> 3) | (u & 255) ^ (u / 3) + (u % 5) + vx.fmax(u, 7))


def test_a_broken_body_is_reported_with_the_compiler_log(env):
    vx, api, L, fake_vec = env
    x, z = fake_vec(64, np.float64, 0x1000), fake_vec(64, np.float64, 0x2000)
    bad = api.UserFunction(np.float64, "broken", [(np.float64, "x")], "return x +;")
    with pytest.raises(vx.VexbError) as ei:
        jit_source(api, L, z, L.SET, bad(x))
    assert "NVRTC could not compile" in str(ei.value) and "return x +;" in str(ei.value)
    with pytest.raises(vx.VexbError):
        api.UserFunction(np.float64, "not an identifier", [], "return 1;")


def _ccsr_source(L, row, col, val, idx_bytes=1, compile=True):
    row, col = np.ascontiguousarray(row, np.int32), np.ascontiguousarray(col, np.int32)
    val = np.ascontiguousarray(val)
    dt = L.F64 if val.dtype == np.float64 else L.F32
    n = C.c_size_t(0)
    args = (row.size - 1, row.ctypes.data, col.ctypes.data, val.ctypes.data, dt, idx_bytes)
    L.check(L.lib().vexb_ccsr_jit_source(*args, None, C.byref(n), 0))
    buf = C.create_string_buffer(n.value + 256)
    cap = C.c_size_t(len(buf))
    L.check(L.lib().vexb_ccsr_jit_source(*args, buf, C.byref(cap), int(compile)))
    return buf.value.decode()


def test_ccsr_specialised_kernel_source_compiles(env):
    """The matrix-specialised CCSR kernel (csrc/ccsr.cu, tunable ccsr.jit): unique rows become code.  NVRTC compiles it
    for sm_100a without a device; running it is a GPU test (tests/test_gpu_ccsr.py)."""
    vx, api, L, _ = env
    from vexcl_b200 import gen
    idx, row, col, val = gen.poisson_ccsr(32)
    src = _ccsr_source(L, row, col, val)
    assert "NVRTC: ok" in src and "vexb_ccsr_jit" in src and "const unsigned char *__restrict__ idx" in src
    # 4 rows per thread (rows of <= 8 entries): one straight-line path for threads whose rows share a unique row, one
    # switch per row otherwise -> every unique row appears 1 + 4 times as a case, its gathers 4 + 4 times
    assert "4 rows per thread" in src and src.count("case ") == 2 * 5 and src.count("__ldg(") == 8 * 8
    assert "__ldg(xi0 + (-1024))" in src and "__ldg(xi3 + (1024))" in src                 # +-n^2 as address immediates
    assert float.fromhex(src.split("__dmul_rn(")[1].split(",")[0]) == 1.0                  # boundary row: 1 * x[i]
    slow = src.split("} else {")[1]                                                        # the per-row switches: row 0 first
    lits = [float.fromhex(t.split(",")[0]) for t in slow.split("__dmul_rn(")[1:9]]
    assert lits == list(val)                                                                # values survive exactly (hex literals)
    # single precision, 2-byte idx, a row longer than one gather group, an empty row
    rng = np.random.default_rng(2)
    row = np.array([0, 0, 11, 14])
    col = rng.integers(-50, 50, 14)
    valf = rng.random(14).astype(np.float32)
    src = _ccsr_source(L, row, col, valf, idx_bytes=2)
    assert "NVRTC: ok" in src and "const unsigned short *__restrict__ idx" in src and "__fmul_rn(" in src
    assert "2 rows per thread" in src
    slow = src.split("} else {")[1]
    assert [np.float32(float.fromhex(t.split("f,")[0])) for t in slow.split("__fmul_rn(")[1:15]] == list(valf)
    with pytest.raises(vx.VexbError, match="too large"):
        _ccsr_source(L, np.arange(41), np.zeros(40, np.int32), np.ones(40), compile=False)


def test_user_defined_stencil_operator_source_compiles(env):
    """VEX_STENCIL_OPERATOR (stencil.hpp:510-680): the generated kernel compiles for sm_100a without a device."""
    vx, api, L, _ = env
    lib = L.lib()

    def source(dtype, width, center, body, compile=True):
        k = C.c_int(-1)
        L.check(lib.vexb_stencil_operator_register(dtype, width, center, body.encode(), C.byref(k)))
        n = C.c_size_t(0)
        L.check(lib.vexb_stencil_operator_source(k.value, None, C.byref(n), 0))
        buf = C.create_string_buffer(n.value + 256)
        cap = C.c_size_t(len(buf))
        L.check(lib.vexb_stencil_operator_source(k.value, buf, C.byref(cap), int(compile)))
        return k.value, buf.value.decode()

    body = "return sin(X[1] - X[0]) + sin(X[0] - X[-1]);"                       # tests/stencil.cpp:187-189
    k1, src = source(L.F64, 3, 1, body)
    assert "NVRTC: ok" in src and "typedef double T;" in src and "#define WIDTH 3" in src and "#define CENTER 1" in src
    assert body in src and "stencil_oper(win + CENTER + threadIdx.x)" in src
    assert source(L.F64, 3, 1, body, compile=False)[0] == k1                      # same definition, same id
    k2, src = source(L.F32, 5, 0, "return X[0] + powf(X[1] + X[4], 3.0f);")
    assert k2 != k1 and "NVRTC: ok" in src and "typedef float T;" in src
    with pytest.raises(vx.VexbError, match="NVRTC could not compile"):
        source(L.F64, 3, 1, "return X[0] +;")
    k = C.c_int(-1)
    assert lib.vexb_stencil_operator_register(L.F64, 3, 3, b"return X[0];", C.byref(k)) == 2      # center outside the stencil
    assert lib.vexb_stencil_operator_register(L.I32, 3, 1, b"return X[0];", C.byref(k)) == 2      # integer operators
    assert lib.vexb_stencil_operator_source(12345, None, C.byref(C.c_size_t(0)), 0) == 2


def test_unregistered_call_is_rejected(env):
    vx, api, L, fake_vec = env
    e = L.Expr()
    e.n_terms, e.n_code = 0, 1
    e.code[0].op, e.code[0].type, e.code[0].arg = L.OP["CALL"], L.F64, 60000
    buf = C.create_string_buffer(64)
    assert L.lib().vexb_eval_path(L.F64, L.SET, C.byref(e), buf, 64) == 2
    assert b"unregistered function" in L.lib().vexb_last_error()


def test_compound_shifts_keep_the_type_of_the_left_operand(env):
    """`a >>= b`: C/C++ (and the reference's emitted `lhs[i] >>= rhs`) shift in the promoted type of the LEFT operand; the
    count's type does not matter.  With a signed a and an unsigned b the shift must stay arithmetic."""
    vx, api, L, fake_vec = env
    a, b = fake_vec(64, np.int32, 0x1000), fake_vec(64, np.uint32, 0x2000)
    src = jit_source(api, L, a, L.RSH, b)
    assert "return (int)((int)lhs[i] >> (int)" in src and "NVRTC: ok" in src
    w = fake_vec(64, np.uint64, 0x3000)
    src = jit_source(api, L, a, L.LSH, w)
    assert "return (int)((int)lhs[i] << (int)" in src
    # other compound operators still use the common type
    src = jit_source(api, L, a, L.ADD, b)
    assert "return (int)((unsigned int)lhs[i] + (unsigned int)" in src


def test_sixteen_terminals_with_converted_scalars_normalise(env):
    """8 vectors + 8 int literals used as doubles: the conversions of the scalars fold in place (no extra terminal slots),
    so a full 16-terminal expression still fits (exprhost.hpp normalize_expr)."""
    vx, api, L, fake_vec = env
    vs = [fake_vec(256, np.float64, 0x1000 * (k + 1)) for k in range(8)]
    e = vs[0] * 2
    for k in range(1, 8):
        e = e + vs[k] * (k + 2)                    # int scalars -> CVT to double
    z = fake_vec(256, np.float64, 0x20000)
    src = jit_source(api, L, z, L.SET, e, compile=False)
    assert "vexb_jit_kernel" in src
    # a scalar pushed twice with different conversions keeps both versions
    s = api.Scalar(3)
    i = fake_vec(256, np.int32, 0x30000)
    src = jit_source(api, L, z, L.SET, (vs[0] * s) + (i * s), compile=False)
    assert "vexb_jit_kernel" in src


def test_process_may_exit_while_background_compilations_run(built):
    """A host program that leaves main() while NVRTC is compiling on a background thread must exit cleanly (NVRTC's own
    lazily registered exit handlers used to run first and pull its statics from under the compilation: SIGSEGV on every
    cold start of tests/cpp/test_vector_arithmetics, profiles/r02_exit_crash.md)."""
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    code = r"""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, %r)
import vexcl_b200 as vx
from vexcl_b200 import api, _lib as L
class FakeCtx:
    nparts, local, is_distributed = 1, [0], False
    def partition(self, n): return vx.partition(n, 1)
def fake_vec(n, dt, addr):
    v = api.vector.__new__(api.vector)
    v.ctx, v.n, v.np_dtype, v.dtype, v.part, v.bufs = FakeCtx(), n, np.dtype(dt), api._vdt(dt), vx.partition(n, 1), {0: C.c_void_p(addr)}
    return v
x, y, z = (fake_vec(1024, np.float64, 0x1000 * (k + 1)) for k in range(3))
def lowered(expr):
    low = api._Lowering(0, 0); low.size = 1024; low.lower(api.wrap(expr)); return low
first = lowered(x * y + z)
L.check(L.lib().vexb_jit_precompile(z.dtype, L.SET, C.byref(first.e), 1))      # the first compilation is a background one, as in a real run:
pend = C.c_int(1)                                                                # our atexit handler is registered BEFORE NVRTC's own
while pend.value:
    L.check(L.lib().vexb_jit_pending(C.byref(pend)))
for k, e in enumerate((vx.sin(x) * y + z / 3.0, (x - y) * (x + y) / z, vx.sqrt(x) + vx.cos(y) * z, x / y / z + 1.0)):
    low = lowered(e)
    L.check(L.lib().vexb_jit_precompile(z.dtype, L.SET, C.byref(low.e), 1))    # background, as the first use of a new shape
pend = C.c_int(0); L.check(L.lib().vexb_jit_pending(C.byref(pend)))
print("pending", pend.value, flush=True)
sys.exit(0)                                                                      # leave while they compile
""" % str(root)
    for _ in range(3):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, f"exit status {r.returncode}\n{r.stdout[-500:]}\n{r.stderr[-1500:]}"
        assert "pending 1" in r.stdout                                                # the exit really overlapped a compilation
