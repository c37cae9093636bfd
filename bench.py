#!/usr/bin/env python
"""bench.py -- CSR SpMV effective GB/s (plus fused axpy / reduction GB/s) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference path

A *step* is one `y = A*x` pass (vex::SpMat<double>, CSR) over the 2-D 5-point Poisson matrix of
BASELINE.json configs[2]: grid 3162 x 3162 per GPU -> 9 998 244 rows, 49 940 644 nonzeros per GPU.
With N GPUs the grid grows along y to 3162 x (3162*N) (weak scaling: every rank keeps a 10M-row
slab and exchanges one grid line of ghosts with each neighbour over NCCL), so
value = N * algorithmic bytes per slab / time.  Algorithmic bytes:
nnz*12 + (nrows+1)*4 + ncols*8 + nrows*8 = 799 252 612 B per slab.

Cache policy: the per-GPU working set (0.8 GB) is 6x the 126 MB L2, so consecutive passes cannot
be served from L2 ("inputs larger than L2").

One JSON line on stdout (rank 0).  Extra keys: roofline, cpu_baseline (N=1), e2e, clocks, gpu_launches,
parity (at EVERY N: every row of every rank's slab against the oracle, max over ranks), extra:
  config[1] numbers (fused a=b+c*d, a+=b+c*d, saxpy, sum(a*b) at N=1e8; N=1 only),
  strong    configs[2] (the 10M-row matrix split over the N GPUs) and configs[3] (3-D 7-pt 256^3 split over the N GPUs):
            us per product back to back and with an L2 flush before every product, with parity,
  cg_step   configs[4] on an SPD matrix: unfused composition and the 4-launch fused iteration, with parity.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

GRID = 3162                       # configs[2]: 3162^2 = 9 998 244 rows
VEC_N = 100_000_000               # configs[1]: N = 1e8 doubles
METRIC = "csr_spmv_effective_gbs"


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_sha() -> str:
    """Fingerprint of the SpMV kernel sources: ncu traffic figures are only quoted for the sources they were measured on."""
    import hashlib
    h = hashlib.sha256()
    for name in ("spmv.cu", "spmv_dev.cuh", "ccsr.cu", "distapply.cu"):
        f = ROOT / "vexcl_b200" / "csrc" / name
        if f.exists():
            h.update(f.read_bytes())
    return h.hexdigest()[:16]


def load_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the last `ncu --set full` capture, or None when the
    kernel sources changed since (profiles/roofline_traffic.json records the fingerprint it was measured on)."""
    tp = ROOT / "profiles" / "roofline_traffic.json"
    try:
        d = json.loads(tp.read_text())
        v = d.get(kernel + "_bytes_per_launch")
        if v is None:
            return None, "no ncu capture for this kernel"
        if d.get("kernel_source_sha16") != kernel_source_sha():
            return None, f"stale: measured at {d.get('measured_at_commit')} ({d.get('date')}), kernel sources changed since"
        return v, f"ncu --set full at {d.get('measured_at_commit')} ({d.get('date')}): {d.get('source')}"
    except Exception as e:                                    # no file: no claim
        return None, f"unavailable ({type(e).__name__})"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in Path(self.path).read_text().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_spmv_setup():
    """The reference's CPU path for configs[2], ready to time: the oracle's restatement of the OpenCL-CPU
    execution model (8 x cores work-groups of one work-item over contiguous row chunks, OpenMP), on buffers
    placed first-touch by the threads that will read them.  Returns (step, y, nbytes, description)."""
    import ctypes as C
    import oracle
    G = oracle.default_groups()
    row, col, val = oracle.poisson(2, GRID)                  # the per-GPU slab of configs[2]
    n = row.size - 1
    nnz = int(row[-1])
    x = oracle.uniform_real(7, n)
    from vexcl_b200 import gen
    nbytes = gen.spmv_bytes(n, n, nnz)
    Lo = oracle.lib()
    Lo.orc_numa_clone.restype = C.c_void_p
    Lo.orc_numa_clone.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    Lo.orc_csr_spmv_raw.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rowp = row.ctypes.data
    c_row = Lo.orc_numa_clone(rowp, n + 1, 8, None, 0, G)
    c_col = Lo.orc_numa_clone(col.ctypes.data, nnz, 8, rowp, n, G)
    c_val = Lo.orc_numa_clone(val.ctypes.data, nnz, 8, rowp, n, G)
    c_x = Lo.orc_numa_clone(x.ctypes.data, n, 8, None, 0, G)
    y = np.zeros(n)
    c_y = Lo.orc_numa_clone(y.ctypes.data, n, 8, None, 0, G)
    del col, val

    def step():
        Lo.orc_csr_spmv_raw(n, c_row, c_col, c_val, c_x, c_y, G)

    def result():
        C.memmove(y.ctypes.data, c_y, n * 8)
        return y

    desc = (f"oracle port of the OpenCL-CPU backend model: {G} work-groups of one work-item over contiguous row chunks, "
            f"OpenMP over {oracle.num_threads()} threads, 64-bit indices as vex::SpMat<double> defaults, first-touch placement")
    return step, result, nbytes, desc, oracle.num_threads()


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def run_reference(args, rank: int, world: int):
    """The reference's own implementation of the path on the host cores (kind "port": the real reference needs
    Boost + OpenCL, absent in this image).  Each step is one full pass over the 10M-row matrix."""
    if rank != 0:
        return
    step, result, nbytes, desc, cores = cpu_spmv_setup()
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    gbs = nbytes * args.steps / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[2]: y = A*x, 2-D 5-pt Poisson CSR, 3162^2 = 9998244 rows, nnz 49940644",
                   "algorithmic_bytes_per_step": nbytes},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": f"full 10M-row matrix, {args.steps} passes; {desc}"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ this repo's arm (GPU)
def cpu_baseline_sample():
    step, result, nbytes, desc, cores = cpu_spmv_setup()
    step()
    reps, t0 = 0, time.perf_counter()
    while reps < 3 or (time.perf_counter() - t0 < 5.0 and reps < 200):
        step()
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": nbytes * reps / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": f"same matrix and x, {reps} passes; {desc}"}, result()


# ------------------------------------------------------------------------------------------------ parity against the oracle
def oracle_slab_product(row, col, val, part, seed0):
    """The oracle's y for this rank's rows: x is U[0,1) from the oracle's libstdc++-equivalent generator, seeded per
    owner part (seed0 + owner), so every rank can rebuild the x values of the neighbours whose columns it touches.
    row/col/val: the strip with global column ids.  Returns (y, row magnitudes)."""
    import oracle                                            # checker only
    col = np.asarray(col, dtype=np.int64)
    if col.size:
        o_lo = int(np.searchsorted(part, int(col.min()), side="right") - 1)
        o_hi = int(np.searchsorted(part, int(col.max()), side="right") - 1)
    else:
        o_lo = o_hi = 0
    xs = [oracle.uniform_real(seed0 + o, int(part[o + 1] - part[o])) for o in range(o_lo, o_hi + 1)]
    x_ext = np.concatenate(xs) if xs else np.empty(0)
    lcol = col - int(part[o_lo])
    lrow = np.asarray(row, dtype=np.int64) - int(row[0])
    return oracle.csr_spmv(lrow, lcol, val, x_ext), oracle.csr_absrow(lrow, lcol, val, x_ext)


def parity_of_product(vx, ctx, A, x, y, row, col, val, part, rank, seed0, allmax, allsum):
    """Write the oracle's x, multiply once through the public call, compare EVERY local row with the oracle.
    Relative to the row magnitude sum_j |a_ij x_j| (the reference's own tolerance is 1e-10, tests/spmv.cpp:33)."""
    import oracle
    k = ctx.local[0]
    n_loc = int(part[k + 1] - part[k])
    x.write(oracle.uniform_real(seed0 + k, n_loc), local_only=True)
    y.assign(-1.0)
    ctx.finish()
    A.apply(x, y, 1.0, False)
    ctx.finish()
    got = y.read() if ctx.is_distributed else y.read()[int(part[k]):int(part[k + 1])]
    want, mag = oracle_slab_product(row, col, val, part, seed0)
    err = float(np.max(np.abs(got - want) / np.maximum(mag, 1e-300))) if n_loc else 0.0
    if not np.all(np.isfinite(got)):
        err = float("inf")
    lo, hi = int(part[k]), int(part[k + 1])
    c = np.asarray(col)
    gi = np.flatnonzero((c < lo) | (c >= hi))                # entries that needed a value from another GPU
    ghost_rows = int(np.unique(np.searchsorted(np.asarray(row, dtype=np.int64) - int(row[0]), gi, side="right")).size)
    return {"max_rel_err_vs_oracle": allmax(err), "rows_checked": int(allsum(n_loc)), "rows_with_ghost_columns_checked": int(allsum(ghost_rows)),
            "tolerance": 1e-10, "x": f"oracle.uniform_real(seed {seed0} + part)"}


class L2Flusher:
    """Writes a buffer larger than L2 (512 MB vs 126 MB) before a timed product."""
    def __init__(self, ctx, L):
        import ctypes as C
        self.ctx, self.L, self.k = ctx, L, ctx.local[0]
        self.n = 512 << 20
        self.p = C.c_void_p()
        L.check(L.lib().vexb_malloc(ctx.devs[self.k], self.n, C.byref(self.p)))

    def flush(self):
        self.L.check(self.L.lib().vexb_memset(self.ctx.devs[self.k], self.p, 1, self.n, self.ctx.streams[self.k]))

    def close(self):
        self.L.check(self.L.lib().vexb_free(self.ctx.devs[self.k], self.p))


def time_flushed(ctx, fn, steps, warmup, barrier, flusher):
    """Sum over `steps` of the device time of fn() alone, each preceded by an (untimed) L2 flush."""
    from vexcl_b200.api import Event
    for _ in range(warmup):
        fn()
    ctx.finish(); barrier()
    pairs = [(Event(ctx), Event(ctx)) for _ in range(steps)]
    for e0, e1 in pairs:
        flusher.flush()
        e0.record()
        fn()
        e1.record()
    pairs[-1][1].sync(); ctx.finish(); barrier()
    return sum(e0.elapsed_ms(e1) for e0, e1 in pairs)


def bench_strong(ctx, vx, L, gen, rank, world, args, barrier, allmax, allsum, peak):
    """The named multi-GPU configurations on the SAME matrices at every N: configs[2] 10M-row 2-D
    Poisson and configs[3] 3-D 7-pt 256^3, rows split over the N GPUs by vex::partition.  Per-GPU working sets at N = 8 are
    100 / 215 MB against a 126 MB L2, so both cache policies are reported: products back to back (what an iterative solver
    sees) and with L2 flushed before every product."""
    from vexcl_b200.api import Graph
    out = {}
    k = ctx.local[0]
    flusher = L2Flusher(ctx, L)
    steps = max(args.steps, 20)
    for name, dim, n in (("configs[2] 2-D 5-pt 3162^2", 2, GRID), ("configs[3] 3-D 7-pt 256^3", 3, 256)):
        N = n ** dim
        part = ctx.partition(N)
        r0, r1 = int(part[k]), int(part[k + 1])
        row, col, val = gen.poisson_strip(dim, n, r0=r0, r1=r1)
        _, nnz_total = gen.poisson_nnz(dim, n)
        A = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_AUTO, strip=True)
        x, y = vx.vector(ctx, N), vx.vector(ctx, N)
        res = {"rows": N, "nnz": nnz_total, "rows_per_gpu": r1 - r0, "halo": "peer-memory push inside the product kernel" if A.peer_halo else
               ("none" if world == 1 else "NCCL send/recv")}
        res["parity"] = parity_of_product(vx, ctx, A, x, y, row, col, val, part, rank, 7, allmax, allsum)
        del row, col, val
        nbytes = gen.spmv_bytes(N, N, nnz_total)

        def step():
            A.apply(x, y, 1.0, False)

        ms = allmax(time_loop(ctx, step, steps, args.warmup, barrier)) / steps
        res["us_per_product"] = ms * 1e3
        res["gbs"] = nbytes / (ms * 1e-3) / 1e9
        res["frac_of_aggregate_hbm_peak"] = res["gbs"] / (peak * world)
        try:
            g = Graph(ctx, step)
            msg = allmax(time_loop(ctx, g.launch, steps, args.warmup, barrier)) / steps
            res["us_per_product_cuda_graph"] = msg * 1e3
            # ten products per graph launch: what an iterative solver replaying whole iterations sees -- one host launch per
            # ten products, so the figure no longer contains the host's enqueue rate (Python, ~10 us per call)
            g10 = Graph(ctx, lambda: [step() for _ in range(10)])
            ms10 = allmax(time_loop(ctx, g10.launch, max(steps // 10, 5), 2, barrier)) / max(steps // 10, 5) / 10
            res["us_per_product_cuda_graph_of_10"] = ms10 * 1e3
            res["frac_of_aggregate_hbm_peak_cuda_graph_of_10"] = nbytes / (ms10 * 1e-3) / 1e9 / (peak * world)
            del g10
            msf = allmax(time_flushed(ctx, g.launch, steps, 3, barrier, flusher)) / steps
            res["us_per_product_l2_flushed"] = msf * 1e3
            res["gbs_l2_flushed"] = nbytes / (msf * 1e-3) / 1e9
            res["frac_of_aggregate_hbm_peak_l2_flushed"] = res["gbs_l2_flushed"] / (peak * world)
            del g
        except vx.VexbError as e:
            res["graph_error"] = str(e)
        res["algorithmic_bytes"] = nbytes
        res["per_gpu_working_set_mb"] = nbytes / world / 1e6
        out[name] = res
        del A, x, y
    flusher.close()
    out["cache_policy"] = ("us_per_product / _cuda_graph: products back to back, no flush (per-GPU working set above L2 only at small N); "
                          "_l2_flushed: a 512 MB memset before every product, product timed alone with events, summed")
    out["timing"] = "CUDA events on the launching stream, max over ranks"
    return out


def bench_reduce_bits(ctx, vx, L, rank, world, dist):
    """vexb_reduce_all across the N GPUs: every rank must hold the same bits, and they must match the oracle's sum."""
    import oracle
    from vexcl_b200.api import DeviceScalar
    n = 1_000_003 * world
    part = ctx.partition(n)
    k = ctx.local[0]
    xs = oracle.uniform_real(11 + k, int(part[k + 1] - part[k]))
    v = vx.vector(ctx, n)
    v.write(xs, local_only=True)
    d = DeviceScalar(ctx)
    vx.Reductor(ctx, np.float64, L.SUM).device(v * v, d)
    got = float(d.get())
    bits = [np.float64(got).view(np.uint64)]
    ref = float(np.dot(xs, xs))
    if dist is not None:
        allb = [None] * world
        dist.all_gather_object(allb, int(bits[0]))
        allr = [None] * world
        dist.all_gather_object(allr, ref)
        bits, ref = allb, float(np.sum(allr))
    return {"identical_bits_on_all_ranks": len(set(int(b) for b in bits)) == 1, "rel_err_vs_numpy": abs(got - ref) / abs(ref), "n": n,
            "combine": "peer memory inside the reduction kernel" if (ctx.peers is not None and world > 1) else ("ncclAllReduce" if world > 1 else "single device")}



def time_loop(ctx, fn, steps, warmup, barrier):
    import vexcl_b200 as vx
    from vexcl_b200.api import Event
    for _ in range(warmup):
        fn()
    ctx.finish()
    barrier()
    e0, e1 = Event(ctx), Event(ctx)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    e1.sync()
    ctx.finish()
    barrier()
    return e0.elapsed_ms(e1)


def run_ours(args, rank: int, world: int, local_rank: int):
    import vexcl_b200 as vx
    from vexcl_b200 import gen, _lib as L
    from vexcl_b200.api import PinnedArray, copy_h2d_async, copy_d2h_async

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = [vx.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)

        def allgather(arr):
            out = [None] * world
            dist.all_gather_object(out, np.asarray(arr))
            return out

        ctx = vx.Context.distributed(rank, world, local_rank, uid[0], allgather, use_peer=not args.no_peer,
                                     peer_halo=False if args.no_peer_halo else None)
        tok = torch.zeros(1, device="cuda")

        def barrier():
            dist.all_reduce(tok)
            torch.cuda.synchronize()

        def max_over_ranks(v):
            t = torch.tensor([v], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        def sum_over_ranks(v):
            t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            return float(t.item())
    else:
        ctx = vx.Context([local_rank])

        def barrier():
            ctx.finish()

        def max_over_ranks(v):
            return v

        def sum_over_ranks(v):
            return float(v)

    peak, peak_src = load_peaks()
    k = ctx.local[0]

    # ---- workload: this rank's 10M-row slab of the 3162 x (3162*N) grid ----------------------------
    nx, ny = GRID, GRID * world
    N = nx * ny
    part = ctx.partition(N)
    r0, r1 = int(part[k]), int(part[k + 1])
    row, col, val = gen.poisson_strip(2, nx, ny, r0=r0, r1=r1, index_dtype=np.int64)
    slab_rows, slab_nnz = r1 - r0, int(row[-1])
    fmt = {"auto": vx.FMT_AUTO, "csr": vx.FMT_CSR, "hell": vx.FMT_HELL, "patterns": vx.FMT_PATTERNS, "sell": vx.FMT_SELL}[args.format]
    A = vx.SpMat(ctx, N, N, row, col, val, fmt, strip=True)
    info = A.info()
    A_alt = None
    if world == 1 and info.loc.fmt != vx.FMT_CSR:
        A_alt = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_CSR, strip=True)     # the TMA row-block stream kernel, reported in extra
    xh = PinnedArray(slab_rows)
    xh.a[:] = np.random.default_rng(7 + rank).random(slab_rows)
    yh = PinnedArray(slab_rows)
    x = vx.vector(ctx, N)
    y = vx.vector(ctx, N)
    x.write(xh.a, local_only=True)
    step_bytes_local = gen.spmv_bytes(slab_rows, slab_rows, slab_nnz)
    # whole-job algorithmic bytes: every rank's slab (identical up to boundary rows); summed exactly below
    step_bytes = sum_over_ranks(step_bytes_local)

    def spmv_step():
        A.apply(x, y, 1.0, False)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = vx.launch_count()
    ms = time_loop(ctx, spmv_step, args.steps, args.warmup, barrier)
    launches = vx.launch_count() - n0 - 0
    # launches counted over warm-up + timed; scale to the timed region
    launches_timed = round(launches * args.steps / (args.steps + args.warmup))
    ms = max_over_ranks(ms)
    value = step_bytes * args.steps / (ms * 1e-3) / 1e9

    # clock probe: keep the same kernel busy long enough for >= 3 samples if the timed region was short
    clocks = None
    if ms < 600:
        # every rank runs the SAME number of further products (the peer-memory halo is a lock-step protocol), enough for
        # ~1.2 s of the same loop, so that rank 0's sampler sees the clocks under this load
        extra_steps = int(min(60000, max(200, 1200.0 / max(ms / args.steps, 1e-3))))
        for _ in range(extra_steps):
            spmv_step()
        ctx.finish()
        barrier()
    if rank == 0:
        clocks = sampler.stop()
        if ms < 600:
            clocks["note"] = "timed region shorter than the sampling period; samples include a ~1.2 s continuation of the same loop on every rank"

    # ---- dominant kernel duration (single-GPU SpMV kernel timed alone with events) ------------------
    from vexcl_b200.api import Event
    ev0, ev1 = Event(ctx), Event(ctx)
    reps = max(args.steps, 20)
    lib = L.lib()
    import ctypes as C
    ctx.finish()
    ev0.record()
    for _ in range(reps):
        L.check(lib.vexb_dspmat_mul_local(A.parts[k], ctx.streams[k], x.bufs[k], y.bufs[k], 1.0, 0))
    ev1.record(); ev1.sync()
    kern_ms = ev0.elapsed_ms(ev1) / reps
    loc = info.loc
    kern_bytes = gen.spmv_bytes(slab_rows, slab_rows, int(loc.nnz))
    achieved = kern_bytes / (kern_ms * 1e-3) / 1e9
    is_hell = loc.fmt == vx.FMT_HELL
    is_patterns = loc.fmt == vx.FMT_PATTERNS
    is_sell = loc.fmt == vx.FMT_SELL
    kname = (f"hell_kernel<double,{int(loc.ell_width)}>" if is_hell else
             f"ccsr_kernel<double> ({int(loc.n_tiles)} row patterns)" if is_patterns else
             "sell_kernel<double>" if is_sell else
             "csr_scalar_kernel<double> (thread per row: the strip's own choice for short even rows)")
    traffic, traffic_note = load_traffic("hell_kernel" if is_hell else "ccsr_kernel" if is_patterns else "csr_kernel")
    # what the stored format must move at least: hybrid ELL has no row pointers, and its columns are 16-bit offsets from
    # the diagonal when the band allows (configs[2]); padding slots are read as columns only
    if is_hell:
        col_bytes = 2 if int(loc.device_bytes) < int(loc.ell_pitch) * int(loc.ell_width) * 12 else 4
        format_bytes = int(loc.ell_pitch) * int(loc.ell_width) * col_bytes + int(loc.nnz) * 8 + int(loc.csr_tail_nnz) * 4 + slab_rows * 16
    else:
        col_bytes = 4
        format_bytes = kern_bytes
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kern_bytes, "kernel_ms": kern_ms,
                "format_bytes_per_launch": format_bytes, "format_column_bytes": col_bytes,
                "frac_by_format_bytes": format_bytes / (kern_ms * 1e-3) / 1e9 / peak,
                "note": "achieved/frac use the canonical CSR figure of BASELINE.md section 3 (12 B per nonzero + row pointers + x + y); "
                        "format_bytes is what the device format actually has to move -- a frac above 1 means the format is more "
                        "compact than canonical CSR, not that work is skipped (see parity)"}

    # ---- end to end through the public call with HOST buffers ---------------------------------------
    # Every step copies its x slab from pinned host memory to the device, multiplies, and copies its y slab
    # back.  The three stages run on three streams with double-buffered device vectors, so the H2D of step
    # i+1 and the D2H of step i-1 overlap the product of step i (PCIe is full duplex).
    import ctypes as C
    dev = ctx.devs[k]
    nb = slab_rows * 8

    def new_stream():
        s = C.c_void_p(); L.check(lib.vexb_stream_create(dev, C.byref(s))); return s

    def new_event():
        e = C.c_void_p(); L.check(lib.vexb_event_create(dev, C.byref(e))); return e

    s_in, s_out, s_main = new_stream(), new_stream(), ctx.streams[k]
    xs, ys = [x, vx.vector(ctx, N)], [y, vx.vector(ctx, N)]
    xhs, yhs = [xh, PinnedArray(slab_rows)], [yh, PinnedArray(slab_rows)]
    xhs[1].a[:] = xh.a
    ev_in, ev_comp, ev_out = ([new_event() for _ in range(2)] for _ in range(3))

    def e2e_step(i):
        b = i & 1
        L.check(lib.vexb_stream_wait_event(dev, s_in, ev_comp[b]))            # x[b] free again (step i-2 multiplied)
        L.check(lib.vexb_h2d(dev, xs[b].bufs[k], xhs[b].a.ctypes.data, nb, s_in, 0))
        L.check(lib.vexb_event_record(dev, ev_in[b], s_in))
        L.check(lib.vexb_stream_wait_event(dev, s_main, ev_in[b]))
        L.check(lib.vexb_stream_wait_event(dev, s_main, ev_out[b]))           # y[b] of step i-2 is on the host
        A.apply(xs[b], ys[b], 1.0, False)
        L.check(lib.vexb_event_record(dev, ev_comp[b], s_main))
        L.check(lib.vexb_stream_wait_event(dev, s_out, ev_comp[b]))
        L.check(lib.vexb_d2h(dev, yhs[b].a.ctypes.data, ys[b].bufs[k], nb, s_out, 0))
        L.check(lib.vexb_event_record(dev, ev_out[b], s_out))

    def e2e_run(n):
        for i in range(n):
            e2e_step(i)
        for b in range(2):
            L.check(lib.vexb_stream_wait_event(dev, s_main, ev_out[b]))       # the last results have landed

    e2e_steps = max(4, min(args.steps, 50))
    e2e_run(4)
    ctx.finish(); barrier()
    t0, t1 = Event(ctx), Event(ctx)
    t0.record()
    e2e_run(e2e_steps)
    t1.record(); t1.sync(); ctx.finish(); barrier()
    ms_e2e = max_over_ranks(t0.elapsed_ms(t1))
    e2e = {"value": step_bytes * e2e_steps / (ms_e2e * 1e-3) / 1e9, "unit": "GB/s",
           "h2d_bytes_per_step": slab_rows * 8 * world, "d2h_bytes_per_step": slab_rows * 8 * world,
           "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
           "note": "per step: x slab pinned host -> device, y = A*x, y slab device -> pinned host; copies of "
                   "neighbouring steps overlap the product (3 streams, double-buffered device vectors)"}
    del xs, ys

    # ---- parity of what was just timed, at every N: all rows of every slab against the oracle -----------------
    extra = {}
    cpu_base = None
    parity = parity_of_product(vx, ctx, A, x, y, row, col, val, part, rank, 7, max_over_ranks, sum_over_ranks)
    parity["halo"] = "peer-memory push inside the product kernel" if A.peer_halo else ("none" if world == 1 else "NCCL send/recv")
    del row, col, val
    try:
        extra["reduce_all"] = bench_reduce_bits(ctx, vx, L, rank, world, dist)
    except vx.VexbError as e:
        extra["reduce_all"] = {"error": str(e)}
    if world == 1:
        try:
            extra["multi_rhs"] = bench_multi_rhs(ctx, vx, A, N, step_bytes, args, barrier, peak)
        except vx.VexbError as e:
            extra["multi_rhs"] = {"error": str(e)}
    if world == 1:
        try:
            extra["fused_product"] = bench_fused_product(ctx, vx, A, N, step_bytes, args, barrier, peak)
        except vx.VexbError as e:
            extra["fused_product"] = {"error": str(e)}
    if A_alt is not None:
        extra["csr_kernels"] = bench_csr_kernels(ctx, vx, gen, A_alt, x, y, step_bytes, args, barrier, peak)
        extra["csr_stream_kernel"] = extra["csr_kernels"]["configs[2] forced to CSR"]["default"]
        del A_alt
    if world == 1:
        # config[1]: vector arithmetic + Reductor, N = 1e8 doubles
        try:
            extra.update(bench_vectors(ctx, vx, args, peak))
        except vx.VexbError as e:
            extra["error"] = str(e)
        try:
            extra["ccsr_spmv"] = bench_ccsr(ctx, vx, args, peak)
        except vx.VexbError as e:
            extra["ccsr_spmv"] = {"error": str(e)}
        try:
            extra["stencil"] = bench_stencil(ctx, vx, args, peak)
        except Exception as e:                                 # first measured by the round-end run: never lose the line
            extra["stencil"] = {"error": repr(e)}
        if not args.no_cpu_baseline:
            cpu_base, _ = cpu_baseline_sample()
    del A, x, y
    if not args.no_strong:
        try:
            extra["strong"] = bench_strong(ctx, vx, L, gen, rank, world, args, barrier, max_over_ranks, sum_over_ranks, peak)
        except vx.VexbError as e:
            extra["strong"] = {"error": str(e)}
    if not args.no_cg:
        # configs[4]: one CG iteration (SpMV + 2 axpy + 2 dot + p update) on the 3-D 7-pt Poisson matrix
        try:
            extra["cg_step"] = bench_cg(ctx, vx, rank, world, args, barrier, max_over_ranks, sum_over_ranks, peak)
        except vx.VexbError as e:
            extra["cg_step"] = {"error": str(e)}
    extra["parity_max_rel_err_vs_oracle"] = parity["max_rel_err_vs_oracle"]

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"configs[2] per GPU: y = A*x, vex::SpMat<double> CSR, 2-D 5-pt Poisson, grid {nx} x {ny} "
                                   f"({N} rows, {slab_rows} per GPU), halo over NCCL send/recv",
                       "algorithmic_bytes_per_step": step_bytes,
                       "format": (f"CSR in; device format chosen by SpMat (as the reference does, spmat.hpp:92-103): hybrid ELL width "
                                  f"{int(loc.ell_width)}, CSR tail {int(loc.csr_tail_nnz)} nnz, 32-bit columns" if is_hell else
                                  f"CSR in; device format: {int(loc.n_tiles)} unique row patterns + one pattern id per row "
                                  f"(VEXB_FMT_PATTERNS, requested explicitly)" if is_patterns else
                                  "CSR in; device format: sliced ELL (SELL-32-sigma)" if is_sell else
                                  "CSR in; device format: CSR (32-bit indices), thread-per-row kernel"),
                       "requested_format": args.format,
                       "cache": "inputs larger than L2 (0.8 GB per GPU vs 126 MB L2)", "partition": "equal weights",
                       "reference_arm": "bench.py --impl reference always times the 10M-row configs[2] matrix on the host cores; at "
                                        "N > 1 this arm is weak-scaled (10M rows PER GPU), so only the rate (GB/s) is comparable, "
                                        "not the matrix; the same-matrix numbers at N GPUs are in extra.strong"},
            "frac_of_aggregate_hbm_peak": value / (peak * world),
            "roofline": roofline, "e2e": e2e, "gpu_launches": launches_timed, "clocks": clocks, "parity": parity,
            "cpu_baseline": cpu_base, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        barrier()
        dist.destroy_process_group()


def bench_csr_kernels(ctx, vx, gen, A_csr, x, y, nbytes, args, barrier, peak):
    """The CSR kernels of csrc/spmv.cu (what vex::sparse::csr and format=csr use; SpMat's AUTO picks hybrid ELL for these
    matrices as the reference does): configs[2] forced to CSR, and an irregular matrix (4M rows, widths U[0,32))."""
    names = {-1: "default", 3: "thread_per_row", 4: "warp_tiles", 5: "warp_rings_tma", 6: "cta_tiles_x_window", 0: "tma_cta_tiles"}
    steps = max(args.steps, 20)
    out = {}

    def sweep(A, xv, yv, nb):
        r = {}
        for variant, nm in names.items():
            vx.set_param("spmv.kernel", variant)
            ms = time_loop(ctx, lambda: A.apply(xv, yv, 1.0, False), steps, 3, barrier) / steps
            r[nm] = {"ms": ms, "gbs": nb / (ms * 1e-3) / 1e9, "frac_of_peak": nb / (ms * 1e-3) / 1e9 / peak}
        vx.set_param("spmv.kernel", -1)
        return r

    out["configs[2] forced to CSR"] = sweep(A_csr, x, y, nbytes)
    n = 4_000_000
    row, col, val = gen.irregular_rows(n, 0, 32, seed=1)
    nb = gen.spmv_bytes(n, n, int(row[-1]))
    xi, yi = vx.vector(ctx, n), vx.vector(ctx, n)
    xi.assign(vx.ElementIndex() * (1.0 / n) + 0.5)
    Ai = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_CSR)
    res = sweep(Ai, xi, yi, nb)
    del Ai
    Ah = vx.SpMat(ctx, n, n, row, col, val, vx.FMT_AUTO)
    ms = time_loop(ctx, lambda: Ah.apply(xi, yi, 1.0, False), steps, 3, barrier) / steps
    res["spmat_auto_format"] = {"ms": ms, "gbs": nb / (ms * 1e-3) / 1e9, "frac_of_peak": nb / (ms * 1e-3) / 1e9 / peak,
                                "fmt": {vx.FMT_CSR: "csr", vx.FMT_HELL: "hybrid ell", vx.FMT_SELL: "sliced ell (SELL-32-sigma)"}.get(int(Ah.info().loc.fmt), "?"),
                                "ell_width": int(Ah.info().loc.ell_width), "csr_tail_nnz": int(Ah.info().loc.csr_tail_nnz)}
    res["rows"], res["nnz"], res["algorithmic_bytes"] = n, int(row[-1]), nb
    out["irregular 4M rows, widths U[0,32)"] = res
    out["note"] = ("default = the strip's own choice (thread per row for short even rows, warp tiles otherwise); "
                   "tma_cta_tiles = the round-1 one-shot TMA kernel")
    return out


def bench_multi_rhs(ctx, vx, A, N, nbytes_one, args, barrier, peak):
    """vex::SpMat * vex::multivector<double,4> on configs[2]: the matrix is streamed once for the four right-hand sides."""
    steps = max(args.steps, 20)
    xs = [vx.vector(ctx, N) for _ in range(4)]
    ys = [vx.vector(ctx, N) for _ in range(4)]
    for r, v in enumerate(xs):
        v.assign(vx.ElementIndex() * (1.0 / N) + 0.25 * r)
    one = time_loop(ctx, lambda: A.apply(xs[0], ys[0], 1.0, False), steps, 3, barrier) / steps
    four = time_loop(ctx, lambda: A.apply_multi(xs, ys, 1.0, False), steps, 3, barrier) / steps
    vx.set_param("spmv.no_multi", 1)
    sep = time_loop(ctx, lambda: A.apply_multi(xs, ys, 1.0, False), steps, 3, barrier) / steps
    vx.set_param("spmv.no_multi", 0)
    return {"ms_one_vector": one, "ms_four_vectors_one_pass": four, "ms_four_separate_products": sep, "ratio_to_one_vector": four / one,
            "gbs_as_four_products": 4 * nbytes_one / (four * 1e-3) / 1e9,
            "note": "per row: matrix entries once + 4 x (x + y); four separate products is what the reference does (operations.hpp:876-880)"}


def bench_fused_product(ctx, vx, A, N, nbytes, args, barrier, peak):
    """`y = x + A*x` as one generated kernel (the product inlined into the consumer, sparse/product.hpp:45-130) against the
    plain product and against the unfused composition (vector part, then y += A*x) on configs[2]."""
    steps = max(args.steps, 20)
    x, y = vx.vector(ctx, N), vx.vector(ctx, N)
    x.assign(vx.ElementIndex() * (1.0 / N) + 0.5)
    plain = time_loop(ctx, lambda: A.apply(x, y, 1.0, False), steps, 3, barrier) / steps
    fused = time_loop(ctx, lambda: y.assign(x + A * x), steps, 3, barrier) / steps
    ctx.fuse_products = False
    try:
        unfused = time_loop(ctx, lambda: y.assign(x + A * x), steps, 3, barrier) / steps
    finally:
        ctx.fuse_products = True
    return {"ms_y=A*x": plain, "ms_y=x+A*x_one_kernel": fused, "ms_y=x+A*x_two_kernels": unfused, "ratio_fused_to_plain": fused / plain,
            "gbs_fused": (nbytes + 8 * N) / (fused * 1e-3) / 1e9,
            "note": "fused: NVRTC kernel with the hybrid-ELL row loop generated into it (VEXB_TERM_SPMV); x[i] is one more 8-byte read per row"}


def bench_cg(ctx, vx, rank, world, args, barrier, max_over_ranks, sum_over_ranks, peak):
    """configs[4]: CG iterations on the 3-D 7-point Poisson operator, SPD form (gen.poisson_strip(spd=True): the
    benchmark generator's identity boundary rows make its matrix non-symmetric, on which CG diverges).
    Grid: 512^3 when 8 ranks (the named configuration: 16.7M rows per GPU), else 256 x 256 x (256 * ranks)
    (the same slab per GPU).  Two solvers: the unfused composition (7 vector kernels + scalar kernels per iteration,
    device-resident alpha / beta) and the fused iteration (4 launches per GPU), each replayed as CUDA graphs."""
    from vexcl_b200 import gen
    from vexcl_b200.solvers import CGDevice, CGFused, cg_bytes_per_iteration, cg_fused_bytes_per_iteration
    if world == 8:
        nx = ny = nz = 512
    else:
        nx = ny = 256
        nz = 256 * world
    N = nx * ny * nz
    k = ctx.local[0]
    part = ctx.partition(N)
    r0, r1 = int(part[k]), int(part[k + 1])
    row, col, val = gen.poisson_strip(3, nx, ny, nz, r0=r0, r1=r1, spd=True)
    nnz_total = int(sum_over_ranks(int(row[-1])))
    A = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_AUTO, strip=True)
    del row, col, val
    b, x = vx.vector(ctx, N), vx.vector(ctx, N)
    iters = max(10, min(args.steps, 50))
    spmv_b = gen.spmv_bytes(N, N, nnz_total)
    out = {"grid": [nx, ny, nz], "rows": N, "nnz": nnz_total, "matrix": "SPD 7-point Laplacian (Dirichlet neighbours dropped), O(1) entries",
           "halo": "peer-memory push inside the product kernel" if A.peer_halo else ("none" if world == 1 else "NCCL send/recv")}
    for name, cls in (("unfused", CGDevice), ("fused", CGFused)):
        # right-hand side: a hash of the index in [-0.5, 0.5) -- rough, so the residual of the Laplacian falls from the
        # first iterations (a smooth b makes |r| grow for a long while although the A-norm of the error falls)
        b.assign(((vx.ElementIndex() * 2654435761) % 1000003) * (1.0 / 1000003) - 0.5)
        x.assign(0.0)
        cg = cls(A, b, x)
        rho0 = cg.residual2()
        res = {}
        for mode in ("stream", "graph"):
            if mode == "graph":
                try:
                    cg.capture()
                except vx.VexbError as e:
                    res["graph_error"] = str(e)
                    break
            ms = max_over_ranks(time_loop(ctx, lambda: cg.run(1), iters, 3, barrier))
            conv = cg_bytes_per_iteration(N, spmv_b)
            res[mode] = {"ms_per_iteration": ms / iters, "gbs_unfused_convention": conv * iters / (ms * 1e-3) / 1e9,
                         "frac_of_aggregate_hbm_peak_unfused_convention": conv * iters / (ms * 1e-3) / 1e9 / (peak * world)}
            if cls is CGFused:
                comp = cg_fused_bytes_per_iteration(N, spmv_b)
                res[mode]["gbs_compulsory"] = comp * iters / (ms * 1e-3) / 1e9
                res[mode]["frac_of_aggregate_hbm_peak_compulsory"] = comp * iters / (ms * 1e-3) / 1e9 / (peak * world)
        res["residual2_start"], res["residual2_after"] = rho0, cg.residual2()
        if cls is CGFused:
            res["product_and_dot_in_one_kernel"] = bool(cg.fused_product)
        out[name] = res
        del cg
    out["bytes_per_iteration_unfused_convention"] = cg_bytes_per_iteration(N, spmv_b)
    out["bytes_per_iteration_fused_compulsory"] = cg_fused_bytes_per_iteration(N, spmv_b)
    out["convention"] = ("unfused reference-equivalent traffic: SpMV + dot 16N + axpy 24N + axpy 24N + dot 8N + p-update 24N; "
                         "fused compulsory: SpMV (p read, q written) + r sweep 24N + x/p sweep 40N")
    # round-1 keys, kept for comparison across rounds (best of the two solvers)
    for mode in ("stream", "graph"):
        cands = [out[n][mode] for n in ("unfused", "fused") if mode in out.get(n, {})]
        if cands:
            best = min(cands, key=lambda r: r["ms_per_iteration"])
            out[mode] = {"ms_per_iteration": best["ms_per_iteration"], "gbs": best["gbs_unfused_convention"],
                         "frac_of_aggregate_hbm_peak": best["frac_of_aggregate_hbm_peak_unfused_convention"]}
    del A, b, x
    try:
        out["parity"] = cg_parity(ctx, vx, rank, world, max_over_ranks)
    except vx.VexbError as e:
        out["parity"] = {"error": str(e)}
    return out


def cg_parity(ctx, vx, rank, world, max_over_ranks):
    """Both solvers against oracle.cg on a small SPD problem split over the same N GPUs: the history of rho = (r, r) over
    12 iterations must agree to 1e-8 relative (sums are associated differently, nothing else differs)."""
    import oracle
    from vexcl_b200 import gen
    from vexcl_b200.solvers import CGDevice, CGFused
    nx, ny, nz = 40, 36, 24 * world
    N = nx * ny * nz
    k = ctx.local[0]
    part = ctx.partition(N)
    row, col, val = gen.poisson_strip(3, nx, ny, nz, spd=True)
    bh = oracle.uniform_real(21, N)
    _, want = oracle.cg(row, col, val, bh, np.zeros(N), 12)
    r0, r1 = int(part[k]), int(part[k + 1])
    srow = row[r0:r1 + 1]
    A = vx.SpMat(ctx, N, N, srow, col[srow[0]:srow[-1]], val[srow[0]:srow[-1]], vx.FMT_AUTO, strip=True) if ctx.is_distributed else \
        vx.SpMat(ctx, N, N, row, col, val, vx.FMT_AUTO)
    out = {"grid": [nx, ny, nz], "iterations": 12, "tolerance": 1e-8}
    for name, cls in (("unfused", CGDevice), ("fused", CGFused)):
        b, x = vx.vector(ctx, N), vx.vector(ctx, N)
        b.write(bh[r0:r1] if ctx.is_distributed else bh, local_only=ctx.is_distributed)
        x.assign(0.0)
        cg = cls(A, b, x)
        hist = []
        for _ in range(12):
            cg.run(1)
            hist.append(cg.residual2())
        err = max(abs(h - w) / abs(w) for h, w in zip(hist, want))
        out[name + "_max_rel_err_of_rho_history"] = max_over_ranks(float(err) if np.all(np.isfinite(hist)) else float("inf"))
        del cg, b, x
    out["rho_first_last_oracle"] = [float(want[0]), float(want[-1])]
    return out


def bench_ccsr(ctx, vx, args, peak):
    """examples/benchmark.cpp:481-606: y += A*x with vex::SpMatCCSR on the 3-D Poisson matrix; n = 256 here (the
    reference uses 128, whose 50 MB working set would sit in the 126 MB L2)."""
    from vexcl_b200 import gen
    n = 256
    N = n ** 3
    idx, row, col, val = gen.poisson_ccsr(n)
    A = vx.SpMatCCSR(ctx, N, idx, row, col, val)
    x, y = vx.vector(ctx, N), vx.vector(ctx, N)
    x.assign(vx.ElementIndex() * (1.0 / N) + 0.5)
    y.assign(0.0)
    steps = max(10, min(args.steps, 40))
    variants = {}
    for name, prm in (("no_hoist", ("ccsr.hoist", 0)), ("table_from_global", ("ccsr.smem", 0)),
                      ("two_rows_per_thread", ("ccsr.kernel", 3))):         # csrc/ccsr.cu tunables, for the record
        vx.set_param(prm[0], prm[1])
        variants[name + "_ms"] = time_loop(ctx, lambda: A.apply(x, y, 1.0, True), steps, 3, ctx.finish) / steps
        vx.set_param(prm[0], 1 if prm[0] != "ccsr.kernel" else 0)
    variants["y=A*x_ms"] = time_loop(ctx, lambda: A.apply(x, y, 1.0, False), steps, 3, ctx.finish) / steps
    # the library default is what is reported
    ms = time_loop(ctx, lambda: A.apply(x, y, 1.0, True), steps, 3, ctx.finish)
    _, nnz = gen.poisson_nnz(3, n)
    t = ms * 1e-3 / steps
    compulsory = N * (A.info().idx_bytes + 24)               # idx as stored on the device + x + y read + y written
    return {"grid": n, "rows": N, "ms": ms / steps, "rows_per_s": N / t, "variants": variants,
            "gbs_compulsory": compulsory / t / 1e9, "frac_of_peak": compulsory / t / 1e9 / peak,
            "gbs_by_reference_formula": (nnz * 20 + 4 * N * 8) / t / 1e9,
            "note": "y += A*x; compulsory bytes = (1-byte idx + x + y in + y out) per row; the reference's own figure "
                    "(benchmark.cpp:563) counts the matrix as if it were CSR"}


def bench_stencil(ctx, vx, args, peak):
    """examples/benchmark.cpp:281-349: b = a * s with a 21-point stencil of 1/21; N = 2^26 doubles here (the reference
    uses 2^20, whose 16 MB working set would sit in L2)."""
    n, width = 1 << 26, 21
    S = vx.stencil(ctx, np.full(width, 1.0 / width), width // 2)
    a, b = vx.vector(ctx, n), vx.vector(ctx, n)
    a.assign(vx.ElementIndex() * (1.0 / n) + 0.5)
    steps = max(10, min(args.steps, 40))
    ms = time_loop(ctx, lambda: S.apply(a, b, 1.0, False), steps, 3, ctx.finish)
    t = ms * 1e-3 / steps
    return {"n": n, "width": width, "ms": ms / steps, "gbs_compulsory": 16 * n / t / 1e9, "frac_of_peak": 16 * n / t / 1e9 / peak,
            "gflops": 2.0 * width * n / t / 1e9, "gbs_by_reference_formula": 2.0 * width * n * 8 / t / 1e9,
            "note": "compulsory bytes = x read once + y written once; the reference's figure (benchmark.cpp:308) counts every tap as a memory access"}


def bench_vectors(ctx, vx, args, peak):
    """configs[1]: examples/benchmark.cpp vector arithmetic + Reductor<double,SUM> at N = 1e8."""
    from vexcl_b200 import _lib as L
    n = VEC_N
    rng = np.random.default_rng(1)
    a, b, c, d = (vx.vector(ctx, n) for _ in range(4))
    a.assign(0.0)
    for v in (b, c, d):
        v.assign(vx.ElementIndex() * 1e-8 + 0.25)          # device-side fill, values in [0.25, 1.25)
    ssum = vx.Reductor(ctx, np.float64, L.SUM)
    steps = max(10, min(args.steps, 40))
    cases = {
        "a=b+c*d": (lambda: a.assign(b + c * d), 32),
        "a+=b+c*d": (lambda: a.__iadd__(b + c * d), 40),
        "a=alpha*a+b": (lambda: a.assign(0.5 * a + b), 24),
        "sum(a*b)": (lambda: ssum(a * b), 16),
    }
    out = {}
    for name, (fn, bpe) in cases.items():
        ms = time_loop(ctx, fn, steps, 3, ctx.finish)
        gbs = bpe * n * steps / (ms * 1e-3) / 1e9
        out[name] = {"gbs": gbs, "frac_of_peak": gbs / peak, "ms": ms / steps, "bytes_per_elem": bpe, "n": n}
    # an expression without a hand-written sweep: interpreter, then the NVRTC kernel built in the background at first use
    import ctypes as C
    gen_expr = lambda: a.assign((b - c) * (b + c) / d + b * 0.5 + c * d)
    vx.set_param("eval.jit", 0)
    ms_i = time_loop(ctx, gen_expr, steps, 3, ctx.finish) / steps
    vx.set_param("eval.jit", 2)
    t0 = time.perf_counter()
    gen_expr(); ctx.finish()                               # first use in default mode: interpreter serves, compilation starts
    first_ms = (time.perf_counter() - t0) * 1e3
    pend = C.c_int(1)
    while pend.value and time.perf_counter() - t0 < 60:
        L.check(L.lib().vexb_jit_pending(C.byref(pend)))
        time.sleep(0.005)
    ready_ms = (time.perf_counter() - t0) * 1e3
    ms_j = time_loop(ctx, gen_expr, steps, 3, ctx.finish) / steps
    g = lambda ms: 32 * n / (ms * 1e-3) / 1e9
    out["generic_expression"] = {"expr": "a = (b-c)*(b+c)/d + b*0.5 + c*d", "bytes_per_elem": 32, "n": n,
                                 "interpreter": {"ms": ms_i, "gbs": g(ms_i), "frac_of_peak": g(ms_i) / peak},
                                 "specialised_nvrtc": {"ms": ms_j, "gbs": g(ms_j), "frac_of_peak": g(ms_j) / peak},
                                 "first_call_ms_default_mode": first_ms, "specialised_kernel_ready_after_ms": ready_ms,
                                 "note": "default mode: the first call is served by the interpreter while NVRTC compiles on a background thread"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cg", action="store_true")
    ap.add_argument("--no-peer", action="store_true", help="combine reductions with ncclAllReduce instead of the fused peer-memory exchange")
    ap.add_argument("--no-peer-halo", action="store_true", help="exchange SpMat halos with NCCL send/recv instead of the in-kernel peer-memory push")
    ap.add_argument("--no-strong", action="store_true", help="skip extra.strong (the named configurations split over the N GPUs)")
    ap.add_argument("--format", default="auto", choices=["auto", "csr", "hell", "patterns", "sell"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world > 1:
        print(f"warning: WORLD_SIZE={world} != --gpus {args.gpus}", file=sys.stderr)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
