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

    # transports: the fused ghost reduction over NVLink peer memory (one kernel writing the neighbours'
    # buffers) must equal the NCCL send/recv path bitwise (same neighbour-order summation), also when
    # repeated back to back (the flag epochs) and with the owner-side scaling
    os.environ["WFX_HALO_TRANSPORT"] = "nccl"
    halo_nccl = partition.Halo(mesh, ctx, comm=halo.comm)
    del os.environ["WFX_HALO_TRANSPORT"]
    assert halo_nccl.transport == "nccl"
    transport = halo.transport
    want_t = os.environ.get("WFX_EXPECT_TRANSPORT")
    assert want_t is None or transport == want_t, f"halo transport is {transport}, expected {want_t}"
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    for it in range(6):
        w = torch.randn(mesh.ndofs, dtype=torch.float64, device=dev, generator=gen)
        wa, wb = w.clone(), w.clone()
        if it % 2:
            halo.update_rev_fwd_scaled(wa, minv)
            halo_nccl.update_rev_fwd_scaled(wb, minv)
        else:
            halo.update_rev_fwd(wa)
            halo_nccl.update_rev_fwd(wb)
        torch.cuda.synchronize()
        assert torch.equal(wa, wb), f"rank {rank}: {transport} and nccl ghost reductions differ (iteration {it})"
    # fp32 halo + fp32 model: masses assembled through a separate fp64 set-up halo
    halo32 = partition.Halo(mesh, ctx, np.float32, comm=halo.comm)
    w32 = torch.randn(mesh.ndofs, dtype=torch.float32, device=dev, generator=gen)
    os.environ["WFX_HALO_TRANSPORT"] = "nccl"
    halo32n = partition.Halo(mesh, ctx, np.float32, comm=halo.comm)
    del os.environ["WFX_HALO_TRANSPORT"]
    wa, wb = w32.clone(), w32.clone()
    halo32.update_rev_fwd(wa)
    halo32n.update_rev_fwd(wb)
    torch.cuda.synchronize()
    assert torch.equal(wa, wb)
    try:
        mass.__class__(mesh, P, ctx=ctx, geometry=geo).assemble(halo32)
        raise AssertionError("assembling the fp64 diagonal through an fp32 halo must fail")
    except wfx.WfxError:
        pass
    dt32 = wfx.cfl_timestep(gmesh.h_min, c0, P, f0)
    geqn32 = wfx.LinearGLLOpt(gmesh, None, P, c0, f0, p0, dtype=np.float32, ctx=ctx)
    geqn32.init()
    geqn32.rk4(0.0, 1.0, dt32, max_steps=10)
    ug32, _ = geqn32.get_state()
    eqn32 = wfx.LinearGLLOpt(mesh, None, P, c0, f0, p0, dtype=np.float32, ctx=ctx, halo=halo32)
    eqn32.init()
    eqn32.rk4(0.0, 1.0, dt32, max_steps=10)
    u32, _ = eqn32.get_state()
    e32 = np.linalg.norm(u32.astype(np.float64) - ug32[mesh.global_dofs]) / max(np.linalg.norm(ug32), 1e-300)
    assert e32 < 2e-5, f"rank {rank}: fp32 distributed RK4 mismatch {e32:.3e}"
    del eqn32, geqn32

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
        print(f"mgpu_check ok: world {world} grid {grid} transport {transport} apply {err:.2e} rk4 u {eu:.2e} v {ev:.2e} fp32 {e32:.1e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
