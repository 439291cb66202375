"""One pass over EVERY kernel of the hot path at BASELINE config 2's size (64^3 cells, P4, fp64, perturbed
geometry), for a per-kernel ncu table (DRAM bytes, time, achieved fraction of the DRAM peak):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,\
sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --csv \
        --log-file gpurun_out/all_kernels.csv python tools/all_kernels.py
    python tools/all_kernels.py --summarise gpurun_out/all_kernels.csv > profiles/r2_all_kernels.md

Numbers taken under ncu are cold-cache and serialised: the table is evidence of bytes moved and of each kernel's
share, never a bench value."""
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(ncell=64, P=4):
    import numpy as np
    import torch
    import wave_fenics_b200 as wfx
    mesh = wfx.create_box_hex(ncell, P, (0.1,) * 3, perturb=0.15)
    geo = wfx.Geometry(mesh, P)                                   # geom_gll_kernel, affine_detect_kernel
    mass = wfx.MassOperator(mesh, P, geometry=geo)                # segsum_kernel, mass_finish_kernel
    stiff = wfx.StiffnessOperator(mesh, P, geometry=geo)
    x = torch.randn(mesh.ndofs, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    for _ in range(2):
        stiff.apply_scaled(x, mass.inverse_diagonal_ptr(), y)     # stiff_brick_kernel (REG), 8 colour launches
    mass(x, y)                                                    # diag_apply_kernel
    cell = wfx.StiffnessOperator(mesh, P, geometry=geo, mode=wfx.capi.STIFF_CELL_COLOUR)
    cell.apply(x, y, beta=1)                                      # stiff_cell_kernel
    del cell, stiff
    # affine mesh: the AFF instantiation of the brick kernel
    amesh = wfx.create_box_hex(ncell, P, (0.1,) * 3, perturb=0.0)
    ageo = wfx.Geometry(amesh, P)
    astiff = wfx.StiffnessOperator(amesh, P, geometry=ageo)
    astiff.apply(x, y, beta=0)
    del astiff, ageo, amesh
    # the streamed-cell kernel where WFX_STIFF_AUTO takes it: P7 fp32 (37^3 cells, ~17.6 M dofs)
    smesh = wfx.create_box_hex(37, 7, (0.1,) * 3, perturb=0.15)
    sgeo = wfx.Geometry(smesh, 7, np.float32)
    sstiff = wfx.StiffnessOperator(smesh, 7, dtype=np.float32, geometry=sgeo)   # permute_g_axes_kernel at set-up
    xs = torch.randn(smesh.ndofs, dtype=torch.float32, device="cuda")
    ys = torch.empty_like(xs)
    sstiff.apply(xs, ys, beta=0)                                  # stiff_cell2_kernel, 8 colour launches
    del sstiff, sgeo, smesh, xs, ys
    # the time loop: stiffness, boundary_kernel, rk_stage_kernel<1..4>, set_source_kernel
    os.environ.setdefault("WFX_WAVE_GRAPH", "0")
    model = wfx.LinearGLLOpt(mesh, None, P, 1500.0, 0.5e6, 60000.0)
    model.init()
    dt = 1e-9
    model.rk4(0.0, 2.5 * dt, dt)
    torch.cuda.synchronize()


def summarise(path):
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        key = (r[ci["ID"]], r[ci["Kernel Name"]])
        per.setdefault(key, {})[r[ci["Metric Name"]]] = (float(r[ci["Metric Value"]].replace(",", "")), r[ci["Metric Unit"]])
    agg = {}
    for (_, name), m in per.items():
        clean = name.replace("void ", "").replace("<unnamed>::", "")
        short = clean.split("(")[0].split("<")[0]
        if short == "stiff_brick_kernel":  # template arguments: T, N, SLOT, W, MINB, REG, layout, AFF, IDS
            targs = clean.split("<", 1)[1].split(">(")[0].replace("LayoutStd<", "LayoutStd[").split(",")
            aff = len(targs) >= 3 and targs[-2].strip() in ("1", "true", "(bool)1")
            short += "<" + targs[0].strip() + ", N=" + targs[1].strip() + (", AFF" if aff else "") + ">"
        elif short in ("stiff_cell2_kernel", "stiff_cell_kernel", "rk_stage_kernel"):
            targs = clean.split("<", 1)[1].split(">(")[0].split(",")
            short += "<" + ", ".join(t.strip() for t in targs[:2]) + ">"
        a = agg.setdefault(short, dict(n=0, t=0.0, rd=0.0, wr=0.0, pct=0.0, regs=0, occ=0.0))
        t, tu = m["gpu__time_duration.sum"]
        t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(tu, 1.0)            # -> us
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = m["dram__bytes_read.sum"][0] * scale.get(m["dram__bytes_read.sum"][1], 1.0)
        wr = m["dram__bytes_write.sum"][0] * scale.get(m["dram__bytes_write.sum"][1], 1.0)
        a["n"] += 1
        a["t"] += t
        a["rd"] += rd
        a["wr"] += wr
        a["pct"] += m["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0]
        a["regs"] = int(m["launch__registers_per_thread"][0])
        a["occ"] += m["sm__warps_active.avg.pct_of_peak_sustained_active"][0]
    print("| kernel | launches | avg time us | DRAM read MB / launch | DRAM write MB / launch | DRAM GB/s | % of DRAM peak (ncu) | regs | warps active % |")
    print("|---|---|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        n = a["n"]
        print(f"| `{k}` | {n} | {a['t'] / n:.1f} | {a['rd'] / n / 1e6:.1f} | {a['wr'] / n / 1e6:.1f} | "
              f"{(a['rd'] + a['wr']) / a['t'] / 1e3:.0f} | {a['pct'] / n:.1f} | {a['regs']} | {a['occ'] / n:.0f} |")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
        summarise(sys.argv[2])
    else:
        run()
