#!/bin/bash
# Two-GPU check: multi-GPU tests, a 2-rank bench line, an ncu launch list of the CG / CSR kernels.
#   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/gpu_check2.sh'
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_reduce.py tests/test_gpu_cpp_frontend.py tests/test_gpu_variants.py -x -q -k "not bench_one_process" 2>&1 | tail -15 > gpurun_out/pytest_multi.log
tail -6 gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 800 gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n2.json").read().splitlines() if l.startswith("{")][-1])
    print("value", d["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"], "parity", d["parity"])
    x = d["extra"]
    print("strong", json.dumps(x.get("strong"))[:2500])
    print("cg", json.dumps(x.get("cg_step"))[:2500])
    print("reduce", x.get("reduce_all"))
except Exception as e:
    print("bench parse failed:", e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cg.csv python scripts/prof_cg.py > gpurun_out/prof_cg.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/launches_cg.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki][:70]
    agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
for k, v in agg.items():
    print(f"{len(v):4d} x {sum(v)/len(v)/1000:9.1f} us  (min {min(v)/1000:.1f})  {k}")
PY
