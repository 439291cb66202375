"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the CPU
restatement on the host cores and prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--ref-cells", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1
    for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "GDoF/s" and d["value"] > 0 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]


def test_default_arm_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device the product arm must exit non-zero, not print a number."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1",
                        "--cells", "4", "--no-cpu-baseline", "--no-rk4"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not any(l.startswith("{") and '"value"' in l for l in r.stdout.splitlines())


def test_reference_arm_is_independent_of_the_product_library():
    """The CPU arm must not import the product package: it runs with the library path pointing
    nowhere (capi raises on import when libwavefx.so is missing), under torch.distributed.run's
    OMP_NUM_THREADS=1, and still uses the host cores it may run on."""
    env = dict(os.environ, WFX_LIB="/nonexistent/libwavefx.so", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-cells", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == max(1, min(len(os.sched_getaffinity(0)), 32))


def test_oracle_mesh_generator_matches_the_product_generator(wfx):
    """oracle/refmesh.py (numpy + the oracle's permutation) and wave-fenics_b200/mesh.py (libwavefx's
    permutation) are independent and produce the same DOLFINx-layout arrays."""
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import refmesh
    for n, P, pert in (((3, 4, 5), 3, 0.1), (4, 4, 0.15), (2, 7, 0.0)):
        a = wfx.create_box_hex(n, P, (0.1, 0.2, 0.3), perturb=pert)
        b = refmesh.box(n, P, (0.1, 0.2, 0.3), perturb=pert)
        assert np.array_equal(a.x, b.x) and np.array_equal(a.xdofs, b.xdofs) and np.array_equal(a.dofmap, b.dofmap)
        assert a.ndofs == b.ndofs
