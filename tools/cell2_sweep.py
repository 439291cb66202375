"""Brick kernel against the streamed-cell kernel per degree and scalar type on one B200 (config 3 sizes,
~17 M dofs, perturbed mesh, fused stiffness + mass apply), with the pipeline depth (cells per slot) as a
parameter.  Prints one JSON line per case; every streamed-cell result is checked against the brick result
on the same inputs.

    python tools/cell2_sweep.py [--degrees 5,6,7] [--cps 2,4,8] [--dtypes f64,f32] [--tag name]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wave_fenics_b200 as wfx  # noqa: E402

L = 0.1
CELLS = {2: 128, 3: 86, 4: 64, 5: 51, 6: 43, 7: 37}


def time_ms(fn, reps, windows=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(windows):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / reps)
    return float(np.median(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--degrees", default="5,6,7")
    ap.add_argument("--cps", default="2,4,8")
    ap.add_argument("--dtypes", default="f64,f32")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--tag", default="")
    ap.add_argument("--no-brick", action="store_true")
    ap.add_argument("--orders", default="brick", help="cell orders of the streamed kernel: brick,colour")
    ap.add_argument("--perm", default="1", help="axis relabelling on (1) / off (0), comma separated")
    args = ap.parse_args()
    peak = 6538.6
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    for P in [int(p) for p in args.degrees.split(",")]:
        mesh = wfx.create_box_hex(CELLS[P], P, (L, L, L), perturb=0.15)
        for name in args.dtypes.split(","):
            dt, tdt = (np.float64, torch.float64) if name == "f64" else (np.float32, torch.float32)
            tol = 1e-12 if name == "f64" else 2e-5
            geo = wfx.Geometry(mesh, P, dt)
            mass = wfx.MassOperator(mesh, P, dtype=dt, geometry=geo)
            x = torch.randn(mesh.ndofs, dtype=tdt, device="cuda")
            yb = torch.full_like(x, float("nan"))
            row = dict(tag=args.tag, P=P, dtype=name, dofs=mesh.ndofs)
            os.environ["WFX_CELL2"] = "0"
            ob = wfx.StiffnessOperator(mesh, P, dtype=dt, geometry=geo)
            info = ob.info()
            ob.apply_scaled(x, mass.inverse_diagonal_ptr(), yb)
            if not args.no_brick:
                ms = time_ms(lambda: ob.apply_scaled(x, mass.inverse_diagonal_ptr(), yb), args.reps)
                row["brick_ms"] = round(ms, 4)
                row["brick_frac"] = round(info["bytes"] / (ms * 1e-3) / 1e9 / peak, 4)
            del ob
            for order in [o for o in args.orders.split(",") if o]:
                for perm in [q for q in args.perm.split(",") if q]:
                    for cps in ([int(c) for c in args.cps.split(",") if c] if order == "colour" else [0]):
                        os.environ["WFX_STREAM_ORDER"] = order
                        os.environ["WFX_AXIS_PERM"] = perm
                        os.environ["WFX_CELL2_CPS"] = str(max(cps, 1))
                        oc = wfx.StiffnessOperator(mesh, P, dtype=dt, geometry=geo, mode=wfx.capi.STIFF_CELL_STREAM)
                        yc = torch.full_like(x, float("nan"))
                        oc.apply_scaled(x, mass.inverse_diagonal_ptr(), yc)
                        err = float((yc.double() - yb.double()).norm() / yb.double().norm())
                        ms = time_ms(lambda: oc.apply_scaled(x, mass.inverse_diagonal_ptr(), yc), args.reps)
                        k = f"{order}{'' if order == 'brick' else cps}_perm{perm}"
                        row[k + "_ms"] = round(ms, 4)
                        row[k + "_frac"] = round(info["bytes"] / (ms * 1e-3) / 1e9 / peak, 4)
                        row[k + "_err"] = err
                        assert err < tol, (P, name, k, err)
                        del oc, yc
            print(json.dumps(row), flush=True)
            del geo, mass, x, yb
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
