#!/usr/bin/env python
"""Benchmark of the waveFEniCS hot path on B200 (see the contract in DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N = 1 workload: BASELINE.json configs[1] -- single-B200 stiffness + mass operator apply,
64^3 hex cells, P4 (16 974 593 dofs), fp64.  One "step" = one fused apply
kv = M^-1 (-c0^2 K u).  `value` = GDoF/s with all inputs resident in HBM; `e2e` = the same
apply through the C-ABI host entry point (pinned host x -> device -> apply -> host y).
N > 1: the same per-GPU block on every rank of a cartesian partition (weak scaling), the
interface dofs of K u reduced with the NCCL halo exchange before the mass inverse.

--impl reference times the CPU restatement of the reference operator (oracle/, built with the
reference's own compiler flags) on the box's host cores, on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GDoF/s stiffness+mass apply (P4 hex, fp64)"
UNIT = "GDoF/s"
L = 0.1


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi SM clock / throttle-reason sampler running during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].startswith("Active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_operator_sample(cells, P, threads, repeats=1):
    """Times the reference CPU operator (dense skernel + b/m) on a cells^3 sample mesh."""
    import wave_fenics_b200 as wfx
    from oracle import oracle
    mesh = wfx.create_box_hex(cells, P, (L, L, L), perturb=0.0)
    G, detJ = oracle.precompute_geometric_data(mesh, P)
    m = np.zeros(mesh.ndofs)
    oracle.mass_apply(mesh, P, detJ, np.ones(mesh.ndofs), m)
    x = np.random.default_rng(42).standard_normal(mesh.ndofs)
    y = np.zeros(mesh.ndofs)
    oracle.stiffness_apply(mesh, P, G, x, y, dense=True, nthreads=threads, fast=True)  # warm-up
    best = 1e300
    for _ in range(repeats):
        y[:] = 0
        t0 = time.perf_counter()
        oracle.stiffness_apply(mesh, P, G, x, y, dense=True, nthreads=threads, fast=True)
        kv = y / m
        best = min(best, time.perf_counter() - t0)
    assert np.isfinite(kv).all()
    return mesh.ndofs, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    threads = min(oracle.max_threads(), 32)  # thread-private y copies: bound the memory
    cells = args.ref_cells
    for _ in range(args.warmup):
        cpu_operator_sample(min(cells, 8), args.P, threads)
    # bounded: if K steps of the requested sample would not end within ~2.5 minutes, shrink the sample
    # mesh (the cost is proportional to the number of cells; the rate in DoF/s is what is reported)
    ndofs, dt = cpu_operator_sample(cells, args.P, threads)
    budget = 150.0
    if dt * args.steps > budget and cells > 8:
        cells = max(8, int(cells * (budget / (dt * args.steps)) ** (1.0 / 3.0)))
        ndofs, dt = cpu_operator_sample(cells, args.P, threads)
    t = [dt]
    for _ in range(args.steps - 1):
        ndofs, dt = cpu_operator_sample(cells, args.P, threads)
        t.append(dt)
    ms = 1e3 * float(np.mean(t))
    value = ndofs / (ms * 1e-3) / 1e9
    sample = f"{cells}^3 cells P{args.P} ({ndofs} dofs) dense skernel + b/m, {threads} OpenMP threads, -Ofast -march=native"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"stiffness+mass apply, P{args.P} hex, fp64 (CPU restatement of the reference operator; bounded sample)",
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_b200(args):
    import torch
    import torch.distributed as dist
    import wave_fenics_b200 as wfx

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = wfx.Context.get(local_rank)
    P, N = args.P, args.cells

    halo = None
    if world == 1:
        mesh = wfx.create_box_hex(N, P, (L, L, L), perturb=args.perturb)
    else:
        from wave_fenics_b200 import partition
        grid = partition.rank_grid(world)
        gshape = tuple(N * g for g in grid)
        mesh = partition.create_box_hex_partition(gshape, P, tuple(L * g for g in grid), grid, rank, perturb=args.perturb)
        halo = partition.make_halo(mesh, ctx, np.float64)
    geo = wfx.Geometry(mesh, P, ctx=ctx)
    stiff = wfx.StiffnessOperator(mesh, P, ctx=ctx, geometry=geo)
    mass = wfx.MassOperator(mesh, P, ctx=ctx, geometry=geo)
    info = stiff.info()
    if halo is not None:
        mass.assemble(halo)
    minv_ptr = mass.inverse_diagonal_ptr()

    dev = torch.device("cuda", local_rank)
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    x = torch.randn(mesh.ndofs, dtype=torch.float64, device=dev, generator=g)
    y = torch.empty_like(x)

    import ctypes as C

    # high priority: the small pack / NCCL / unpack kernels must not queue behind the interior batches
    comm_stream = torch.cuda.Stream(device=dev, priority=-1) if halo is not None else None

    def step():
        if halo is None:
            stiff.apply_scaled(x, minv_ptr, y)
        else:
            # interface cells -> their ghost reduction (+ 1/m on the owner) on a side stream while
            # the interior cells run with the fused 1/m; shared dofs are never scaled in-kernel
            main = torch.cuda.current_stream()
            stiff.apply_part(x, y, 0, beta=0, scale_ptr=minv_ptr)
            comm_stream.wait_stream(main)
            halo.update_rev_fwd_scaled(y, minv_ptr, stream=comm_stream.cuda_stream)
            stiff.apply_part(x, y, 1, beta=0, scale_ptr=minv_ptr)
            main.wait_stream(comm_stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms = ms_total / args.steps
    ndofs_global = mesh.ndofs_global
    value = ndofs_global / (ms * 1e-3) / 1e9

    # end-to-end through the C-ABI host entry point, pinned host buffers, copies in the timed region
    e2e = None
    if world == 1:
        xh = torch.empty(mesh.ndofs, dtype=torch.float64).pin_memory()
        yh = torch.empty(mesh.ndofs, dtype=torch.float64).pin_memory()
        xh.copy_(x)
        call = lambda: wfx.capi.call("wfx_stiffness_mass_apply_host", stiff.handle, mass.handle,
                                     C.c_void_p(xh.data_ptr()), C.c_void_p(yh.data_ptr()))
        for _ in range(3):
            call()
        n_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            call()
        dt = (time.perf_counter() - t0) / n_e2e
        assert torch.allclose(yh, y.cpu(), rtol=0, atol=0), "host-path result differs from device path"
        e2e = {"value": mesh.ndofs / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": mesh.ndofs * 8,
               "d2h_bytes_per_step": mesh.ndofs * 8, "ms_per_step": dt * 1e3}
    else:
        # every rank: its part of x from pinned host memory, the distributed apply (halo included),
        # its part of the result back to pinned host memory
        xh = torch.empty(mesh.ndofs, dtype=torch.float64).pin_memory()
        yh = torch.empty(mesh.ndofs, dtype=torch.float64).pin_memory()
        xh.copy_(x)

        def host_step():
            x.copy_(xh, non_blocking=True)
            step()
            yh.copy_(y, non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(3):
            host_step()
        n_e2e = max(3, min(args.steps, 10))
        sync_all()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            host_step()
        sync_all()
        t = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        nb = torch.tensor([mesh.ndofs * 8], dtype=torch.int64, device=dev)
        dist.all_reduce(nb)
        e2e = {"value": ndofs_global / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(nb.item()),
               "d2h_bytes_per_step": int(nb.item()), "ms_per_step": dt * 1e3,
               "note": "per-rank pinned buffers; bytes summed over ranks"}

    # second half of the BASELINE metric: one full RK4 step (4 x stiffness + boundary term +
    # halo + fused stage update) of the wave model on the same mesh, through wfx_wave_rk4
    rk4 = None
    if not args.no_rk4:
        del x, y
        eqn = wfx.LinearGLLOpt(mesh, None, P, 1500.0, 0.5e6, 6e4, ctx=ctx, halo=halo)
        eqn.init()
        dt_w = wfx.cfl_timestep(mesh.h_min, 1500.0, P, 0.5e6)
        eqn.rk4(0.0, 1.0, dt_w, max_steps=2)
        sync_all()
        k_rk = max(3, min(args.steps, 10))
        ev0.record()
        eqn.rk4(2 * dt_w, 1.0, dt_w, max_steps=k_rk)
        ev1.record()
        sync_all()
        ms_rk = ev0.elapsed_time(ev1) / k_rk
        if world > 1:
            t = torch.tensor([ms_rk], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_rk = float(t.item())
        u_chk, _ = eqn.get_state()
        # (the wave starts at x = 0: ranks away from the source face are still at rest)
        assert np.isfinite(u_chk).all() and (rank != 0 or np.abs(u_chk).max() > 0)
        # algorithmic bytes per step: 4 stages x (G + dofmap + read un +
        # write b) + the fused stage updates (10 + 11 + 11 + 7 vector passes)
        s8 = 8
        b_step = 4 * (info["num_cells"] * info["num_dofs"] * (6 * s8 + 4) + mesh.ndofs * 2 * s8) + 39 * mesh.ndofs * s8
        rk4 = {"ms_per_step": ms_rk, "steps": k_rk, "dofs_global": int(ndofs_global),
               "gdof_steps_per_s": ndofs_global / (ms_rk * 1e-3) / 1e9,
               "bytes_per_step_per_gpu": b_step, "achieved_gbs_per_gpu": b_step / (ms_rk * 1e-3) / 1e9}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    if rk4:
        rk4["frac_of_hbm_peak"] = rk4["achieved_gbs_per_gpu"] / peak
    # dominant kernel = stiff_brick_kernel: the step is its `nlaunches` colour launches, so the
    # kernel's average launch duration is ms / nlaunches and its algorithmic bytes per launch are
    # bytes / nlaunches; achieved = bytes per step / step time.
    achieved = info["bytes"] / (ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        threads = min(oracle.max_threads(), 32)
        nd_cpu, t_cpu = cpu_operator_sample(args.ref_cells, P, threads)
        cpu = {"value": nd_cpu / t_cpu / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.ref_cells}^3 cells P{P} ({nd_cpu} dofs), dense skernel + b/m, {threads} OpenMP threads, -Ofast -march=native"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"stiffness+mass apply kv=M^-1(-c0^2 K u), {N}^3 hex cells per GPU, P{P}, fp64, "
                                   f"{'perturbed (non-affine)' if args.perturb else 'affine'} geometry, general 6-entry G per point",
                       "cells_per_gpu": N ** 3, "dofs_global": int(ndofs_global), "degree": P,
                       "l2_policy": "inputs larger than L2 (G 1.57 GB + vectors 0.4 GB per apply vs 126 MB L2)",
                       "partition": "1" if world == 1 else "x".join(map(str, partition.rank_grid(world)))},
            "clocks": clocks,
            "e2e": e2e if e2e else {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                    "note": "multi-GPU e2e = device-resident path"},
            "gpu_launches": args.steps * (info["nlaunches"] + (0 if halo is None else 5)),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "stiff_brick_kernel",
                         "bytes_per_step": info["bytes"], "launches_per_step": info["nlaunches"]},
            "cpu_baseline": cpu, "rk4": rk4}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 20 (b200 arm), 5 (reference arm)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=64, help="cells per axis per GPU")
    ap.add_argument("--P", type=int, default=4)
    ap.add_argument("--perturb", type=float, default=0.15)
    ap.add_argument("--ref-cells", type=int, default=48, help="CPU sample mesh (cells per axis)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rk4", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        # 20 applies = 9 ms: boost clocks.  Sustained load (1000 applies, 0.5 s) runs into the board's
        # software power cap: SM clock 1.70 instead of 1.97 GHz, 0.50 instead of 0.46 ms per apply
        # (profiles/r1_scaling.md) -- the kernel is latency-bound, so it follows the SM clock.
        args.steps = 20 if args.impl == "b200" else 5
    args.steps = max(1, args.steps)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
