"""Summarise ncu reports (read here, no GPU needed):   python scripts/ncu_summarise.py gpurun_out/r02_ncu_*.ncu-rep > profiles/r02_ncu_summary.md

Per captured launch: duration, DRAM bytes (read + write), DRAM / L2 / L1 / issue utilisation, occupancy, registers, and
the top warp-stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active)."""
import csv
import io
import subprocess
import sys
from pathlib import Path

KEYS = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1_pct",
    "l1tex__t_bytes.sum": "l1_bytes",
    "sm__inst_executed.sum": "inst",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "lim_regs",
    "launch__occupancy_limit_shared_mem": "lim_smem",
    "launch__occupancy_limit_warps": "lim_warps",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit",
    "l1tex__t_sector_hit_rate.pct": "l1_hit",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "smsp__cycles_active.avg": "smsp_cycles",
}


def to_bytes(v, unit):
    u = unit.lower()
    f = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
    return v * f


def to_us(v, unit):
    u = unit.lower()
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "s": 1e6}.get(u, 1)


def main():
    for path in sys.argv[1:]:
        if path.endswith(".csv"):                              # already exported with `ncu -i rep --page raw --csv`
            rows = list(csv.reader(open(path)))
        else:
            r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
            if r.returncode != 0:
                print(f"## {path}\n\ncannot read: {r.stderr[:300]}\n")
                continue
            rows = list(csv.reader(io.StringIO(r.stdout)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        print(f"## {Path(path).name}\n")
        for row in rows[2:]:
            g = lambda k: row[ix[k]] if k in ix else ""
            name = g("Kernel Name")[:90]
            vals = {}
            for k, short in KEYS.items():
                if k in ix and row[ix[k]] not in ("", "n/a"):
                    try:
                        vals[short] = (float(row[ix[k]].replace(",", "")), units[ix[k]])
                    except ValueError:
                        pass
            t_us = to_us(*vals["time"]) if "time" in vals else float("nan")
            rd = to_bytes(*vals["dram_rd"]) if "dram_rd" in vals else 0.0
            wr = to_bytes(*vals["dram_wr"]) if "dram_wr" in vals else 0.0
            stalls = []
            for h in hdr:
                if h.startswith("smsp__average_warp") and "stalled_" in h and h.endswith(".ratio"):
                    try:
                        stalls.append((float(row[ix[h]].replace(",", "")), h.split("stalled_")[1].split("_per_")[0].replace(".ratio", "")))
                    except (ValueError, IndexError):
                        pass
            stalls.sort(reverse=True)
            f = lambda k, fmt="{:.1f}": fmt.format(vals[k][0]) if k in vals else "-"
            print(f"* `{name}` grid {f('grid', '{:.0f}')} x {f('block', '{:.0f}')}, {f('regs', '{:.0f}')} regs: **{t_us:.1f} us**, "
                  f"DRAM {rd / 1e6:.1f} + {wr / 1e6:.1f} MB = {(rd + wr) / 1e6:.1f} MB ({(rd + wr) / t_us / 1e3 if t_us == t_us else 0:.0f} GB/s, {f('dram_pct')} % of peak); "
                  f"L2 {f('l2_pct')} % (hit {f('l2_hit')} %), L1 {f('l1_pct')} % (hit {f('l1_hit')} %), issue {f('issue_pct')} %, FP64 pipe {f('fp64_pct')} %, LSU {f('lsu_pct')} %, "
                  f"warps active {f('occ_pct')} % (limits: regs {f('lim_regs', '{:.0f}')}, smem {f('lim_smem', '{:.0f}')}, warps {f('lim_warps', '{:.0f}')} blocks/SM); "
                  f"top stalls: " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:5]))
        print()


if __name__ == "__main__":
    main()
