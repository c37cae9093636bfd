import json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vexcl_b200 as vx
from vexcl_b200 import _lib as L
from vexcl_b200.api import Event, UserFunction
ctx = vx.Context([0])
def timeit(fn, reps=30, warm=3):
    for _ in range(warm): fn()
    ctx.finish(); e0, e1 = Event(ctx), Event(ctx); e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.sync(); return e0.elapsed_ms(e1) / reps
n = 100_000_000
a, b, c, d = (vx.vector(ctx, n) for _ in range(4))
for v in (a, b, c, d): v.assign(vx.ElementIndex() * 1e-8 + 0.25)
sq = UserFunction(np.float64, "sq", [(np.float64, "x")], "return x * x;")
cases = {"a=b+c*d": lambda: a.assign(b + c * d), "a=sin(b)*c+sqrt(d)": lambda: a.assign(vx.sin(b) * c + vx.sqrt(d)),
         "a=(b-c)*(b+c)/d+b*0.5": lambda: a.assign((b - c) * (b + c) / d + b * 0.5)}
for name, fn in cases.items():
    for mode in ("sweep/interp", "interp", "jit"):
        vx.set_param("eval.force_interp", 0 if mode == "sweep/interp" else 1)
        vx.set_param("eval.jit", 1 if mode == "jit" else 0)
        ms = timeit(fn)
        print(json.dumps(dict(case=name, mode=mode, path=a.eval_path(L.SET, 0) if False else mode, ms=ms, gbs=32 * n / ms / 1e6)), flush=True)
vx.set_param("eval.force_interp", 0); vx.set_param("eval.jit", 0)
ms = timeit(lambda: a.assign(sq(b) + c))
print(json.dumps(dict(case="a=sq(b)+c (VEX_FUNCTION)", mode="jit", ms=ms, gbs=24 * n / ms / 1e6)))
# auto mode (default): first uses interpreted, then specialised
vx.set_param("eval.force_interp", 0); vx.set_param("eval.jit", 2)
fn = lambda: a.assign((b - c) * (b + c) / d + b * 0.25)
ctx.finish()
import time
for k in range(6):
    t = time.perf_counter(); fn(); ctx.finish(); print(json.dumps(dict(case="auto", call=k, ms=(time.perf_counter() - t) * 1e3)))
print(json.dumps(dict(case="auto steady", ms=timeit(fn), gbs=32 * n / timeit(fn) / 1e6)))
