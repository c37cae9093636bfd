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

One JSON line on stdout (rank 0).  Extra keys: roofline, cpu_baseline (N=1), e2e, clocks,
gpu_launches, extra (config[1] numbers: fused a=b+c*d, a+=b+c*d, saxpy, sum(a*b) at N=1e8).
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

        ctx = vx.Context.distributed(rank, world, local_rank, uid[0], allgather, use_peer=not args.no_peer)
        tok = torch.zeros(1, device="cuda")

        def barrier():
            dist.all_reduce(tok)
            torch.cuda.synchronize()

        def max_over_ranks(v):
            t = torch.tensor([v], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
    else:
        ctx = vx.Context([local_rank])

        def barrier():
            ctx.finish()

        def max_over_ranks(v):
            return v

    peak, peak_src = load_peaks()
    k = ctx.local[0]

    # ---- workload: this rank's 10M-row slab of the 3162 x (3162*N) grid ----------------------------
    nx, ny = GRID, GRID * world
    N = nx * ny
    part = ctx.partition(N)
    r0, r1 = int(part[k]), int(part[k + 1])
    row, col, val = gen.poisson_strip(2, nx, ny, r0=r0, r1=r1, index_dtype=np.int64)
    slab_rows, slab_nnz = r1 - r0, int(row[-1])
    fmt = {"auto": vx.FMT_AUTO, "csr": vx.FMT_CSR, "hell": vx.FMT_HELL, "patterns": vx.FMT_PATTERNS}[args.format]
    A = vx.SpMat(ctx, N, N, row, col, val, fmt, strip=True)
    info = A.info()
    A_alt = None
    if world == 1 and info.loc.fmt != vx.FMT_CSR:
        A_alt = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_CSR, strip=True)     # the TMA row-block stream kernel, reported in extra
    del row, col, val
    xh = PinnedArray(slab_rows)
    xh.a[:] = np.random.default_rng(7 + rank).random(slab_rows)
    yh = PinnedArray(slab_rows)
    x = vx.vector(ctx, N)
    y = vx.vector(ctx, N)
    x.write(xh.a, local_only=True)
    step_bytes_local = gen.spmv_bytes(slab_rows, slab_rows, slab_nnz)
    # whole-job algorithmic bytes: every rank's slab (identical up to boundary rows); summed exactly below
    if dist is not None:
        import torch
        t = torch.tensor([float(step_bytes_local)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        step_bytes = float(t.item())
    else:
        step_bytes = float(step_bytes_local)

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
    if rank == 0:
        if ms < 600:
            t_end = time.perf_counter() + 1.2
            if world == 1:
                while time.perf_counter() < t_end:
                    for _ in range(50):
                        spmv_step()
                    ctx.finish()
        clocks = sampler.stop()
        if ms < 600 and world == 1:
            clocks["note"] = "timed region shorter than the sampling period; samples include a 1.2 s continuation of the same loop"

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
    kname = (f"hell_kernel<double,{int(loc.ell_width)}>" if is_hell else
             f"ccsr_kernel<double> ({int(loc.n_tiles)} row patterns)" if is_patterns else "csr_stream_kernel<double>")
    traffic = None
    tp = ROOT / "profiles" / "roofline_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("hell_kernel_bytes_per_launch" if is_hell else
                                                     "ccsr_kernel_bytes_per_launch" if is_patterns else "csr_stream_kernel_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kern_bytes, "kernel_ms": kern_ms}

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

    # ---- parity spot check of what was just timed (rank-local rows against numpy on 4096 rows) -------
    extra = {}
    cpu_base = None
    if A_alt is not None:
        ms_alt = time_loop(ctx, lambda: A_alt.apply(x, y, 1.0, False), max(args.steps, 20), 3, barrier)
        n_alt = max(args.steps, 20)
        extra["csr_stream_kernel"] = {"gbs": step_bytes * n_alt / (ms_alt * 1e-3) / 1e9, "ms": ms_alt / n_alt,
                                      "frac_of_peak": step_bytes * n_alt / (ms_alt * 1e-3) / 1e9 / peak,
                                      "note": "same matrix forced to the TMA-staged CSR row-block kernel (format=csr)"}
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
            cpu_base, y_cpu = cpu_baseline_sample()
            import oracle                      # checker only: parity of the timed kernel on the oracle's x
            x.write(oracle.uniform_real(7, slab_rows))
            y.assign(A * x)
            got = y.read()
            err = float(np.max(np.abs(got - y_cpu) / (np.abs(y_cpu) + 1e-300)))
            extra["parity_max_rel_err_vs_oracle"] = err
    if not args.no_cg:
        # configs[4]: one CG iteration (SpMV + 2 axpy + 2 dot + p update) on the 3-D 7-pt Poisson matrix
        del A, x, y
        try:
            extra["cg_step"] = bench_cg(ctx, vx, rank, world, args, barrier, max_over_ranks, peak)
        except vx.VexbError as e:
            extra["cg_step"] = {"error": str(e)}

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
                                  f"CSR in; device format: CSR row-block stream (TMA-staged tiles of {int(loc.tile_nnz)} nnz), 32-bit indices"),
                       "requested_format": args.format,
                       "cache": "inputs larger than L2 (0.8 GB per GPU vs 126 MB L2)", "partition": "equal weights"},
            "frac_of_aggregate_hbm_peak": value / (peak * world),
            "roofline": roofline, "e2e": e2e, "gpu_launches": launches_timed, "clocks": clocks,
            "cpu_baseline": cpu_base, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        barrier()
        dist.destroy_process_group()


def bench_cg(ctx, vx, rank, world, args, barrier, max_over_ranks, peak):
    """configs[4]: CG iteration on the 3-D 7-point Poisson matrix (examples/benchmark.cpp:357-415 generator).
    Grid: 512^3 when 8 ranks (the named configuration: 16.7M rows per GPU), else 256 x 256 x (256 * ranks)
    (the same slab per GPU).  Device-resident alpha / beta, one CUDA graph per iteration."""
    from vexcl_b200 import gen
    from vexcl_b200.solvers import CGDevice, cg_bytes_per_iteration
    if world == 8:
        nx = ny = nz = 512
    else:
        nx = ny = 256
        nz = 256 * world
    N = nx * ny * nz
    k = ctx.local[0]
    part = ctx.partition(N)
    r0, r1 = int(part[k]), int(part[k + 1])
    row, col, val = gen.poisson_strip(3, nx, ny, nz, r0=r0, r1=r1)
    val /= float((nx - 1) ** 2)                                   # O(1) entries so that the iteration stays finite
    A = vx.SpMat(ctx, N, N, row, col, val, vx.FMT_AUTO, strip=True)
    _, nnz_total = gen.poisson_nnz(3, nx, ny, nz)
    del row, col, val
    b, x = vx.vector(ctx, N), vx.vector(ctx, N)
    b.assign(vx.ElementIndex() * (1.0 / N) + 0.5)
    x.assign(0.0)
    cg = CGDevice(A, b, x)
    iters = max(10, min(args.steps, 50))
    out = {}
    for mode in ("stream", "graph"):
        if mode == "graph":
            try:
                cg.capture()
            except vx.VexbError as e:
                out["graph_error"] = str(e)
                break
        ms = max_over_ranks(time_loop(ctx, lambda: cg.run(1), iters, 3, barrier))
        nbytes = cg_bytes_per_iteration(N, gen.spmv_bytes(N, N, nnz_total))
        gbs = nbytes * iters / (ms * 1e-3) / 1e9
        out[mode] = {"ms_per_iteration": ms / iters, "gbs": gbs, "frac_of_aggregate_hbm_peak": gbs / (peak * world)}
    out.update({"grid": [nx, ny, nz], "rows": N, "nnz": nnz_total, "bytes_per_iteration": nbytes,
                "convention": "unfused reference-equivalent traffic: SpMV + dot 16N + axpy 24N + axpy 24N + dot 8N + p-update 24N",
                "residual2_after": cg.residual2()})
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
    ap.add_argument("--format", default="auto", choices=["auto", "csr", "hell", "patterns"])
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
