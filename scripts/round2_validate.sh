#!/bin/bash
# Round-2 first GPU call: everything written in round 1 after the GPU budget ran out, in one go (single GPU).
#   gpurun --timeout 600 -- 'bash scripts/round2_validate.sh'
set -u
mkdir -p gpurun_out
export VEXB_RUN_UNVERIFIED=1
# 1. NVRTC-specialised CCSR kernel + the skipped two-slice C++ stencil run
timeout 300 python -m pytest tests/test_gpu_unverified.py tests/test_gpu_ccsr.py tests/test_gpu_cpp_frontend.py -q -k "unverified or ccsr or stencil or scalar or 16_bit or pattern or constants" 2>&1 | tail -25 > gpurun_out/r02_unverified_tests.log
# 1b. the C++ stencil binary incl. the NVRTC user-defined operator case (one slice, then two)
for p in 1 2; do VEXCL_TEST_PARTS=$p timeout 120 tests/cpp/bin/test_stencil 12345 2>&1 | tail -4 >> gpurun_out/r02_unverified_tests.log; done
# 2. the two-slice stencil binary under compute-sanitizer (it stopped after 'two_stencils' in round 1)
VEXCL_TEST_PARTS=2 timeout 200 compute-sanitizer --tool memcheck tests/cpp/bin/test_stencil 12345 > gpurun_out/r02_stencil_2slices_memcheck.log 2>&1
# 2b. host-side memory errors of the header front end: the multivector / stencil tests under ASan + UBSan
for t in test_multivector test_stencil; do
  g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -I include tests/cpp/$t.cpp -o /tmp/${t}_asan \
      -L vexcl_b200 -lvexb200 -Wl,-rpath,$PWD/vexcl_b200 -lpthread -ldl 2>&1 | tail -3
  ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0 VEXCL_TEST_PARTS=2 timeout 200 /tmp/${t}_asan 12345 2>&1 | tail -6 >> gpurun_out/r02_asan.log
done
# 3. timings: CCSR variants incl. ccsr.jit, stencil throughput (also in bench.py extra.stencil)
timeout 100 python - > gpurun_out/r02_ccsr_jit_probe.json 2>&1 <<'PY'
import json, sys
sys.path.insert(0, ".")
import vexcl_b200 as vx
from vexcl_b200 import gen
from bench import time_loop
n = 256; N = n ** 3
ctx = vx.Context([0])
idx, row, col, val = gen.poisson_ccsr(n)
A = vx.SpMatCCSR(ctx, N, idx, row, col, val)
x, y = vx.vector(ctx, N), vx.vector(ctx, N)
x.assign(vx.ElementIndex() * (1.0 / N) + 0.5); y.assign(0.0)
out = {}
for jit in (0, 1):
    vx.set_param("ccsr.jit", jit)
    for append in (True, False):
        ms = time_loop(ctx, lambda: A.apply(x, y, 1.0, append), 40, 3, ctx.finish) / 40
        out[f"jit={jit},append={append}"] = {"ms": ms, "gbs_compulsory": N * (25 if append else 17) / (ms * 1e-3) / 1e9}
# CSR variants on config 3 (2-D Poisson 3162^2): TMA tiles (0) vs thread per row (3)
row, col, val = gen.poisson_strip(2, 3162)
Nc = row.size - 1
Ac = vx.SpMat(ctx, Nc, Nc, row, col, val, vx.FMT_CSR)
xc, yc = vx.vector(ctx, Nc), vx.vector(ctx, Nc)
xc.assign(vx.ElementIndex() * (1.0 / Nc) + 0.5)
nbytes = gen.spmv_bytes(Nc, Nc, int(row[-1]))
for k in (0, 3):
    vx.set_param("spmv.kernel", k)
    ms = time_loop(ctx, lambda: Ac.apply(xc, yc, 1.0, False), 40, 3, ctx.finish) / 40
    out[f"csr spmv.kernel={k}"] = {"ms": ms, "gbs": nbytes / (ms * 1e-3) / 1e9}
vx.set_param("spmv.kernel", 0)
# the headline path with 16-bit ELL columns (spmv.col16 is read when the matrix is built)
for c16 in (0, 1):
    vx.set_param("spmv.col16", c16)
    Ah = vx.SpMat(ctx, Nc, Nc, row, col, val, vx.FMT_HELL)
    ms = time_loop(ctx, lambda: Ah.apply(xc, yc, 1.0, False), 100, 5, ctx.finish) / 100
    out[f"hell spmv.col16={c16}"] = {"ms": ms, "gbs": nbytes / (ms * 1e-3) / 1e9}
    del Ah
vx.set_param("spmv.col16", 0)
# row-pattern strip (VEXB_FMT_PATTERNS -> ccsr_kernel), generic and NVRTC-specialised
Ap = vx.SpMat(ctx, Nc, Nc, row, col, val, vx.FMT_PATTERNS)
for jit in (0, 1):
    vx.set_param("ccsr.jit", jit)
    ms = time_loop(ctx, lambda: Ap.apply(xc, yc, 1.0, False), 100, 5, ctx.finish) / 100
    out[f"patterns ccsr.jit={jit}"] = {"ms": ms, "gbs_effective": nbytes / (ms * 1e-3) / 1e9, "patterns": int(Ap.info().loc.n_tiles)}
vx.set_param("ccsr.jit", 0)
print(json.dumps(out, indent=1))
PY
tail -5 gpurun_out/r02_unverified_tests.log; tail -5 gpurun_out/r02_stencil_2slices_memcheck.log; cat gpurun_out/r02_ccsr_jit_probe.json
