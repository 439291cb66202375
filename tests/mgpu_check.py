"""Multi-GPU check, run under torchrun (one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py
Every rank also solves the whole problem on its own GPU (one-rank path, already checked against
the oracle) and compares its partition's entries with it."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wave_fenics_b200 as wfx  # noqa: E402
from wave_fenics_b200 import partition  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = wfx.Context.get(local)
    P, L, c0, f0, p0 = 4, 0.1, 1500.0, 0.5e6, 6e4
    grid = partition.rank_grid(world)
    gshape = (8, 6, 4)
    lengths = (L, L * 6 / 8, L * 4 / 8)
    gmesh = wfx.create_box_hex(gshape, P, lengths, perturb=0.15)
    mesh = partition.create_box_hex_partition(gshape, P, lengths, grid, rank, perturb=0.15)
    halo = partition.Halo(mesh, ctx)
    dev = torch.device("cuda", local)

    # operator apply + halo reduction
    xg = np.random.default_rng(42).standard_normal(gmesh.ndofs)
    gop = wfx.StiffnessOperator(gmesh, P, ctx=ctx)
    yg = torch.zeros(gmesh.ndofs, dtype=torch.float64, device=dev)
    gop(torch.from_numpy(xg).to(dev), yg)
    op = wfx.StiffnessOperator(mesh, P, ctx=ctx)
    y = torch.zeros(mesh.ndofs, dtype=torch.float64, device=dev)
    op(torch.from_numpy(xg[mesh.global_dofs]).to(dev), y)
    halo.update_rev_fwd(y)
    torch.cuda.synchronize()
    ref = yg.cpu().numpy()[mesh.global_dofs]
    err = np.linalg.norm(y.cpu().numpy() - ref) / np.linalg.norm(ref)
    assert err < 1e-12, f"rank {rank}: halo-reduced apply differs, rel L2 {err:.3e}"

    # split form: interface cells, scaled ghost reduction on a side stream, interior cells with the
    # fused 1/m -- must equal (assembled K x) / (assembled m) on every copy
    geo = wfx.Geometry(mesh, P, ctx=ctx)
    mass = wfx.MassOperator(mesh, P, ctx=ctx, geometry=geo)
    mass.assemble(halo)
    gmass = wfx.MassOperator(gmesh, P, ctx=ctx)
    want = (yg.cpu().numpy() / gmass.diagonal())[mesh.global_dofs]
    assert np.allclose(mass.diagonal(), gmass.diagonal()[mesh.global_dofs], rtol=1e-14, atol=0)
    y2 = torch.full((mesh.ndofs,), float("nan"), dtype=torch.float64, device=dev)
    xl = torch.from_numpy(xg[mesh.global_dofs]).to(dev)
    side = torch.cuda.Stream(device=dev)
    minv = mass.inverse_diagonal_ptr()
    op.apply_part(xl, y2, 0, beta=0, scale_ptr=minv)
    side.wait_stream(torch.cuda.current_stream())
    halo.update_rev_fwd_scaled(y2, minv, stream=side.cuda_stream)
    op.apply_part(xl, y2, 1, beta=0, scale_ptr=minv)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    err2 = np.linalg.norm(y2.cpu().numpy() - want) / np.linalg.norm(want)
    assert err2 < 1e-12, f"rank {rank}: split/scaled apply differs, rel L2 {err2:.3e}"

    # forward update alone: ghosts take the owner's value
    z = torch.from_numpy(xg[mesh.global_dofs].copy()).to(dev)
    z[mesh.size_local:] = -1.0
    halo.update_fwd(z)
    torch.cuda.synchronize()
    assert np.array_equal(z.cpu().numpy(), xg[mesh.global_dofs])

    # RK4 with the halo against the one-rank solve
    dt = wfx.cfl_timestep(gmesh.h_min, c0, P, f0)
    geqn = wfx.LinearGLLOpt(gmesh, None, P, c0, f0, p0, ctx=ctx)
    geqn.init()
    geqn.rk4(0.0, 1.0, dt, max_steps=25)
    ug, vg = geqn.get_state()
    eqn = wfx.LinearGLLOpt(mesh, None, P, c0, f0, p0, ctx=ctx, halo=halo)
    eqn.init()
    eqn.rk4(0.0, 1.0, dt, max_steps=25)
    u, v = eqn.get_state()
    eu = np.linalg.norm(u - ug[mesh.global_dofs]) / np.linalg.norm(ug)
    ev = np.linalg.norm(v - vg[mesh.global_dofs]) / np.linalg.norm(vg)
    assert eu < 1e-12 and ev < 1e-12, f"rank {rank}: RK4 mismatch {eu:.3e} {ev:.3e}"
    # replicas of shared dofs are bitwise identical across ranks
    full = torch.full((gmesh.ndofs,), float("nan"), dtype=torch.float64, device=dev)
    full[torch.from_numpy(mesh.global_dofs).to(dev)] = torch.from_numpy(u).to(dev)
    gathered = [torch.empty_like(full) for _ in range(world)]
    dist.all_gather(gathered, full)
    st = torch.stack(gathered)
    lo = torch.where(torch.isnan(st), torch.full_like(st, float("inf")), st).min(0).values
    hi = torch.where(torch.isnan(st), torch.full_like(st, float("-inf")), st).max(0).values
    assert torch.equal(lo, hi), "copies of a shared dof differ between ranks"
    dist.barrier()
    if rank == 0:
        print(f"mgpu_check ok: world {world} grid {grid} apply {err:.2e} rk4 u {eu:.2e} v {ev:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
