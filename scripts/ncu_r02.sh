#!/bin/bash
# Round-2 ncu evidence, one GPU.   gpurun --timeout 1500 -- 'bash scripts/ncu_r02.sh [targets...]'
#   1. launch list of the default bench command (shares of a step)
#   2. --set full captures of the hot kernels on their bench workloads (scripts/prof_r02.py)
# Reports are written outside gpurun_out/ (they can exceed its 64 MiB limit); what comes back is the raw-metric CSV, the
# per-instruction source page and -- when small -- the report itself.  Summaries: scripts/ncu_summarise.py -> profiles/.
set -u
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --clock-control none"
want() { [ $# -eq 0 ] || return 0; }
TARGETS="${*:-launches hell csr_scalar csr_warp ccsr ccsr_jit stencil dist_apply cg_update interp hell_multi}"
KEEP_REP="${KEEP_REP:-sell stencil ccsr_jit}"
has() { case " $TARGETS " in *" $1 "*) return 0;; esac; return 1; }
if has launches; then
    timeout 500 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv \
        python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench.stdout 2> gpurun_out/r02_launches_bench.err
    echo "launch list rc=$? lines=$(wc -l < gpurun_out/r02_launches_bench.csv)"
fi
cap() {  # name regex count targets...
    local name=$1 rx=$2 cnt=$3; shift 3
    has $name || return 0
    timeout 400 $NCU --set full --import-source on -k "regex:$rx" -s 2 -c $cnt -f -o /tmp/ncu/r02_ncu_$name \
        python scripts/prof_r02.py "$@" > gpurun_out/r02_ncu_$name.log 2>&1
    local rc=$?
    ncu -i /tmp/ncu/r02_ncu_$name.ncu-rep --page raw --csv > gpurun_out/r02_ncu_${name}_raw.csv 2>/dev/null
    ncu -i /tmp/ncu/r02_ncu_$name.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_ncu_${name}_sass.csv 2>/dev/null
    gzip -f gpurun_out/r02_ncu_${name}_sass.csv
    local sz=$(stat -c %s /tmp/ncu/r02_ncu_$name.ncu-rep 2>/dev/null || echo 0)
    # gpurun_out/ comes back only if it stays under 64 MiB: keep a report only while the directory is under 30 MB
    local used=$(du -sm gpurun_out | cut -f1)
    case " $KEEP_REP " in *" $name "*) [ "$sz" -gt 0 ] && [ "$sz" -lt 9000000 ] && [ "$used" -lt 30 ] && cp /tmp/ncu/r02_ncu_$name.ncu-rep gpurun_out/;; esac
    echo "$name rc=$rc rep=$sz raw=$(stat -c %s gpurun_out/r02_ncu_${name}_raw.csv) sass.gz=$(stat -c %s gpurun_out/r02_ncu_${name}_sass.csv.gz)"
}
cap hell '^hell_kernel' 2 hell
cap csr_scalar 'csr_scalar_kernel' 2 csr_poisson
cap csr_warp 'csr_warp_kernel' 2 csr_irregular
cap ccsr '^ccsr_kernel' 2 ccsr
cap ccsr_jit 'vexb_ccsr_jit' 2 ccsr_jit
cap stencil 'stencil_kernel' 2 stencil
cap dist_apply 'dist_apply_kernel' 1 cg
cap cg_update 'cg_update' 2 cg
cap interp '^interp_kernel' 1 vec
cap hell_multi 'hell_multi_kernel' 2 hell multi_rhs
cap sell 'sell_kernel' 2 sell
if [ "$(du -sm gpurun_out | cut -f1)" -ge 60 ]; then rm -f gpurun_out/*.ncu-rep; echo "reports dropped to stay under the size limit"; fi
du -sh gpurun_out
