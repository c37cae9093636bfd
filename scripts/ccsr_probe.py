"""Times vex::SpMatCCSR (y += A*x, 3-D Poisson 256^3) under the kernel's tunables.  --one: a few launches of the
default configuration (for ncu)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import vexcl_b200 as vx
from vexcl_b200 import gen
from bench import time_loop

n = 256
N = n ** 3
ctx = vx.Context([0])
idx, row, col, val = gen.poisson_ccsr(n)
A = vx.SpMatCCSR(ctx, N, idx, row, col, val)
x, y = vx.vector(ctx, N), vx.vector(ctx, N)
x.assign(vx.ElementIndex() * (1.0 / N) + 0.5)
y.assign(0.0)
if "--one" in sys.argv:
    for _ in range(6):
        A.apply(x, y, 1.0, True)
    ctx.finish()
    sys.exit(0)
out = {}
for kernel in (3,):
    vx.set_param("ccsr.kernel", kernel)
    for append in (True, False):
        ms = time_loop(ctx, lambda: A.apply(x, y, 1.0, append), 40, 3, ctx.finish) / 40
        out[f"kernel={kernel},append={append}"] = {"ms": ms, "gbs_compulsory": N * (25 if append else 17) / (ms * 1e-3) / 1e9}
vx.set_param("ccsr.kernel", 1)
for hoist in (0, 1):
    for batch in (1, 8):
        for append in (True, False):
            vx.set_param("ccsr.hoist", hoist)
            vx.set_param("ccsr.batch", batch)
            ms = time_loop(ctx, lambda: A.apply(x, y, 1.0, append), 40, 3, ctx.finish) / 40
            out[f"hoist={hoist},batch={batch},append={append}"] = {"ms": ms, "gbs_compulsory": N * (25 if append else 17) / (ms * 1e-3) / 1e9}
print(json.dumps(out, indent=1))
