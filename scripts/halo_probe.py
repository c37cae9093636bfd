"""Where does the multi-GPU step time go?  torchrun --nproc-per-node 2 scripts/halo_probe.py"""
import os, sys, json
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch, torch.distributed as dist
import vexcl_b200 as vx
from vexcl_b200 import gen, _lib as L
from vexcl_b200.api import Event
import ctypes as C

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
uid = [vx.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
def allgather(a):
    out = [None] * world; dist.all_gather_object(out, np.asarray(a)); return out
ctx = vx.Context.distributed(rank, world, lr, uid[0], allgather)
k = rank
nx, ny = 3162, 3162 * world
N = nx * ny
part = ctx.partition(N)
row, col, val = gen.poisson_strip(2, nx, ny, r0=int(part[k]), r1=int(part[k + 1]))
fmt = {"auto": vx.FMT_AUTO, "csr": vx.FMT_CSR}[sys.argv[1] if len(sys.argv) > 1 else "auto"]
A = vx.SpMat(ctx, N, N, row, col, val, fmt, strip=True)
x, y = vx.vector(ctx, N), vx.vector(ctx, N)
x.assign(vx.ElementIndex() * 1e-9 + 0.5)
tok = torch.zeros(1, device="cuda")
def barrier():
    dist.all_reduce(tok); torch.cuda.synchronize()
def timeit(fn, reps=200):
    for _ in range(5): fn()
    ctx.finish(); barrier()
    e0, e1 = Event(ctx), Event(ctx); e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.sync(); ctx.finish(); barrier()
    t = torch.tensor([e0.elapsed_ms(e1) / reps], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3
lib = L.lib()
res = {}
res["apply_us"] = timeit(lambda: A.apply(x, y, 1.0, False))
h = A.parts[k]
info = A.info()
res["interior_plus_bnd_us"] = timeit(lambda: L.check(lib.vexb_dspmat_mul_local(h, ctx.streams[k], x.bufs[k], y.bufs[k], 1.0, 0)))
res["pack_us"] = timeit(lambda: L.check(lib.vexb_dspmat_pack(h, ctx.streams[k], x.bufs[k])))
res["remote_us"] = timeit(lambda: L.check(lib.vexb_dspmat_mul_remote(h, ctx.streams[k], y.bufs[k], 1.0)))
comms = ctx._arr(ctx.comms); parts = ctx._arr(A.parts); streams = ctx._arr(ctx.streams)
res["exchange_only_us"] = timeit(lambda: L.check(lib.vexb_halo_exchange(1, comms, parts, streams)))
vx.set_param("dspmat.no_peer_halo", 1)            # the NCCL send/recv path (pack, exchange, interior, two boundary kernels)
res["apply_nccl_path_us"] = timeit(lambda: A.apply(x, y, 1.0, False))
vx.set_param("dspmat.no_peer_halo", 0)
res["n_ghost"] = int(info.n_ghost); res["n_send"] = int(info.n_send); res["loc_fmt"] = int(info.loc.fmt); res["loc_rows"] = int(info.loc.nrows)
if rank == 0: print(json.dumps(res))
barrier(); dist.destroy_process_group()
