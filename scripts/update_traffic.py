"""profiles/roofline_traffic.json from the last ncu --set full captures (gpurun_out/r02_ncu_*_raw.csv, scripts/ncu_r02.sh):
dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernels, stamped with the fingerprint of the
kernel sources they were measured on (bench.py quotes the figure only while that fingerprint still matches).

    python scripts/update_traffic.py [commit]"""
import csv
import datetime
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(path: Path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    vals = []
    for row in rows[2:]:
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(row[ix[k]].replace(",", "")) * UNIT.get(units[ix[k]], 1)
        vals.append(tot)
    return int(sum(vals) / len(vals)) if vals else None


out = {"kernel_source_sha16": bench.kernel_source_sha(),
       "measured_at_commit": sys.argv[1] if len(sys.argv) > 1 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip(),
       "date": datetime.date.today().isoformat(),
       "source": "profiles/r02_ncu_summary.md (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch, average of the captured launches)"}
for key, name in (("hell_kernel", "hell"), ("csr_kernel", "csr_scalar"), ("ccsr_kernel", "ccsr_jit")):
    f = ROOT / "gpurun_out" / f"r02_ncu_{name}_raw.csv"
    if f.exists():
        out[key + "_bytes_per_launch"] = dram_bytes(f)
(ROOT / "profiles" / "roofline_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
