#!/bin/bash
# N-GPU check (N = number of visible GPUs): multi-GPU tests, then the bench line at N ranks.
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_check8.sh'
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -k "not bench_one_process" 2>&1 | tail -15 > gpurun_out/pytest_multi_n$N.log
tail -4 gpurun_out/pytest_multi_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 600 gpurun_out/bench_n$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n$N.json").read().splitlines() if l.startswith("{")][-1])
    print("value", d["value"], "ms", d["ms_per_step"], "frac", d["frac_of_aggregate_hbm_peak"], "launches", d["gpu_launches"], "parity", d["parity"])
    x = d["extra"]
    print("strong", json.dumps(x.get("strong"))[:3000])
    print("cg", json.dumps(x.get("cg_step"))[:3000])
    print("reduce", x.get("reduce_all"))
except Exception as e:
    print("bench parse failed:", e)
PY
