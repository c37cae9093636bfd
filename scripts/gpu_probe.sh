#!/bin/bash
# One-GPU round trip: cold-start exit check, timing probe, GPU test suite, short bench line.
#   gpurun --timeout 1700 -- 'bash scripts/gpu_probe.sh [probe targets]'
set -u
mkdir -p gpurun_out
for i in 1 2; do VEXCL_TEST_PARTS=2 tests/cpp/bin/test_vector_arithmetics 12345 > gpurun_out/va_$i.log 2>&1; echo "va run $i rc=$?"; done
timeout 600 python scripts/probe_r02.py "$@" > gpurun_out/r02_probe.json 2> gpurun_out/r02_probe.err; echo "probe rc=$?"; tail -3 gpurun_out/r02_probe.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_n1.err
