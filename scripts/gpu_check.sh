#!/bin/bash
# One-GPU check: the GPU test suite, then a short bench line.   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -120 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "roofline", {k: d["roofline"][k] for k in ("kernel", "frac", "frac_by_format_bytes", "kernel_ms")})
    print("parity", d["parity"]); print("e2e", d["e2e"]["value"])
    x = d["extra"]
    print("csr", json.dumps(x.get("csr_kernels"), indent=0)[:1800])
    print("strong", json.dumps(x.get("strong"))[:1500])
    print("cg", json.dumps(x.get("cg_step"))[:2500])
    print("reduce", x.get("reduce_all"))
    for k in ("a=b+c*d", "sum(a*b)", "ccsr_spmv", "stencil"):
        print(k, {kk: vv for kk, vv in x.get(k, {}).items() if kk in ("gbs", "frac_of_peak", "ms", "gbs_compulsory")})
except Exception as e:
    print("bench parse failed:", e)
PY
