"""Where the distributed apply spends its time (run under torchrun, one rank per GPU):
interface batches, interior batches, ghost reduction, and the overlapped step of bench.py.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/mgpu_breakdown.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wave_fenics_b200 as wfx
from wave_fenics_b200 import partition

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
P, N, L = 4, 64, 0.1
grid = partition.rank_grid(world)
gshape = tuple(N * g for g in grid)
ctx = wfx.Context(lr)
mesh = partition.create_box_hex_partition(gshape, P, tuple(L * g for g in grid), grid, rank, perturb=0.15)
halo = partition.make_halo(mesh, ctx, np.float64)
geo = wfx.Geometry(mesh, P, ctx=ctx)
stiff = wfx.StiffnessOperator(mesh, P, ctx=ctx, geometry=geo)
mass = wfx.MassOperator(mesh, P, ctx=ctx, geometry=geo)
mass.assemble(halo)
minv = mass.inverse_diagonal_ptr()
dev = torch.device("cuda", lr)
x = torch.randn(mesh.ndofs, dtype=torch.float64, device=dev)
y = torch.empty_like(x)
comm = torch.cuda.Stream(device=dev, priority=-1)


def iface():
    stiff.apply_part(x, y, 0, beta=0, scale_ptr=minv)


def interior():
    stiff.apply_part(x, y, 1, beta=0, scale_ptr=minv)


def ghost():
    halo.update_rev_fwd_scaled(y, minv, stream=torch.cuda.current_stream().cuda_stream)


def serial():
    iface(); ghost(); interior()


def overlapped():
    main = torch.cuda.current_stream()
    iface()
    comm.wait_stream(main)
    halo.update_rev_fwd_scaled(y, minv, stream=comm.cuda_stream)
    interior()
    main.wait_stream(comm)


def whole():
    stiff.apply_part(x, y, -1, beta=0, scale_ptr=minv)


# the same operator without the interface / interior split: one pass, then the ghost reduction
flat = wfx.StiffnessOperator(mesh, P, ctx=ctx, geometry=geo, split=False)


def flat_all():
    flat.apply_part(x, y, -1, beta=0, scale_ptr=minv)


def flat_serial():
    flat_all(); ghost()


def timeit(f, n=20):
    for _ in range(5):
        f()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


info = stiff.info()
res = {k: timeit(f) for k, f in [("interface", iface), ("interior", interior), ("ghost", ghost),
                                 ("all batches, no halo", whole), ("serial", serial), ("overlapped", overlapped),
                                 ("unsplit plan, no halo", flat_all), ("unsplit plan + ghost", flat_serial)]}
if rank == 0:
    print(f"ranks {world} grid {grid} launches {info['nlaunches']} halo transport {halo.transport}")
    for k, v in res.items():
        print(f"  {k:24s} {v * 1e3:8.1f} us")
dist.destroy_process_group()
