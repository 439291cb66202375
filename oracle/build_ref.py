"""Builds oracle/_ref/libwfref_cuda.so from the REFERENCE's own CUDA sources (read in place under
/root/reference, never copied) plus the forwarding shim oracle/ref_cuda_shim.cu.

TEST INFRASTRUCTURE ONLY.  build_cpu() does the same for the two cell kernels of the CPU operators
(skernel, mkernel of common/operators.hpp), see oracle/ref_cpu_shim.cpp.  What compiles of the reference without its un-vendored dependencies
(DOLFINx, Basix, xtensor, FFCx, MPI) are the device primitives of the hackathon GPU operators:
common/cuda/scatter.cu (gather, atomicAdd scatter) and common/cuda/transform.cu (pointwise detJ
multiply).  Everything else on the hot path includes <dolfinx.h> / <basix/...> / <xtensor/...> and is
unbuildable here.  The library travels to the GPU box with the snapshot
(oracle/_ref/ is git-ignored, not gpurun-ignored); it needs a GPU to run, so only `-m gpu` tests use it.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("WFX_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libwfref_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build(force=False):
    """Returns the library path, or None when the reference sources are not present (GPU box)."""
    srcs = [os.path.join(REF, "common", "cuda", f) for f in ("scatter.cu", "transform.cu")]
    if not all(os.path.exists(s) for s in srcs):
        return LIB if os.path.exists(LIB) else None
    shim = os.path.join(HERE, "ref_cuda_shim.cu")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(s) for s in srcs + [shim]):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    cmd = [NVCC, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
           "-ccbin", "/usr/bin/g++", "-I", os.path.join(REF, "common", "cuda"), *srcs, shim, "-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("building oracle/_ref/libwfref_cuda.so failed")
    return LIB


LIB_CPU = os.path.join(OUT, "libwfref_cpu.so")
# the same code with the reference's own optimisation level (demo/cpu_planar3d/CMakeLists.txt:27: -Ofast
# -march=native -mprefer-vector-width=512) for the CPU baseline of bench.py; -march is the portable
# x86-64-v3 here because the library is built in this container and runs on the GPU box's host
LIB_CPU_FAST = os.path.join(OUT, "libwfref_cpu_fast.so")


def _braces(lines, i):
    """Index of the line that closes the first brace opened on or after line i."""
    depth, seen, j = 0, False, i
    while True:
        for ch in lines[j]:
            if ch == "{":
                depth += 1
                seen = True
            elif ch == "}":
                depth -= 1
        if seen and depth == 0:
            return j
        j += 1


def _cut(text, marker, after=None, with_template=True):
    """The text of the definition whose first line contains `marker` (the first one after a line
    containing `after`, if given): from its `template <...>` line, when it has one, to its closing brace."""
    lines = text.split("\n")
    lo = 0
    if after is not None:
        lo = next(i for i, l in enumerate(lines) if after in l)
    hits = [i for i in range(lo, len(lines)) if marker in lines[i]]
    if not hits:
        raise RuntimeError("reference header: no definition of %r" % marker)
    i = hits[0]
    start = i - 1 if (with_template and "template" in lines[i - 1]) else i
    return "\n".join(lines[start:_braces(lines, i) + 1]) + "\n"


def build_cpu(force=False):
    """oracle/_ref/libwfref_cpu.so: the reference's own CPU code of the hot path that needs none of its
    dependencies' arithmetic -- skernel / mkernel, the two operators' call loops (common/operators.hpp),
    kernels::copy / axpy and LinearGLLOpt::init / f0 / f1 / rk4 (common/LinearGLL.hpp) -- cut out of the
    headers where they lie and compiled against the stand-ins of oracle/ref_cpu_shim.cpp with the parity
    flags of the oracle (-O2 -ffp-contract=off).  Returns the library path, or None when the reference
    sources are not present and no prebuilt library exists."""
    hdr = os.path.join(REF, "common", "operators.hpp")
    whdr = os.path.join(REF, "common", "LinearGLL.hpp")
    shim = os.path.join(HERE, "ref_cpu_shim.cpp")
    if not (os.path.exists(hdr) and os.path.exists(whdr)):
        return LIB_CPU if os.path.exists(LIB_CPU) else None
    if not force and os.path.exists(LIB_CPU) and os.path.exists(LIB_CPU_FAST) and min(
            os.path.getmtime(LIB_CPU), os.path.getmtime(LIB_CPU_FAST)) >= max(
            os.path.getmtime(hdr), os.path.getmtime(whdr), os.path.getmtime(shim), os.path.getmtime(__file__)):
        return LIB_CPU
    os.makedirs(OUT, exist_ok=True)
    ops, wave = open(hdr).read(), open(whdr).read()
    pieces = {
        "ref_cell_kernels.inc": _cut(ops, "inline void mkernel(") + _cut(ops, "inline void skernel("),
        "ref_mass_call.inc": _cut(ops, "void operator()(", after="class MassOperatorCPU"),
        "ref_stiffness_call.inc": _cut(ops, "void operator()(", after="class StiffnessOperator"),
        "ref_wave_kernels.inc": _cut(wave, "namespace kernels", with_template=False),
        "ref_wave_methods.inc": "".join(_cut(wave, sig, after="class LinearGLLOpt", with_template=False)
                                        for sig in ("void init()", "void f0(", "void f1(", "void rk4(")),
    }
    written = []
    try:
        for name, text in pieces.items():
            path = os.path.join(OUT, name)
            with open(path, "w") as fh:
                fh.write(text)
            written.append(path)
        for flags, lib in ((["-O2", "-ffp-contract=off"], LIB_CPU), (["-Ofast"], LIB_CPU_FAST)):
            cmd = ["g++", *flags, "-std=c++17", "-march=x86-64-v3", "-shared", "-fPIC", "-I", OUT, shim, "-o", lib]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("building %s failed" % lib)
    finally:
        for path in written:
            os.remove(path)  # reference text never stays in the tree, not even in the ignored directory
    return LIB_CPU


LIB_MESH = os.path.join(OUT, "libwfref_mesh.so")


def build_mesh(force=False):
    """oracle/_ref/libwfref_mesh.so: the reference's partition arithmetic decompose3d /
    compute_cartesian_indices (demo/gpu_cg/mesh.hpp:37-62), cut out of the header and compiled against the
    stand-in of oracle/ref_mesh_shim.cpp.  None when neither the sources nor a prebuilt library exist."""
    hdr = os.path.join(REF, "demo", "gpu_cg", "mesh.hpp")
    shim = os.path.join(HERE, "ref_mesh_shim.cpp")
    if not os.path.exists(hdr):
        return LIB_MESH if os.path.exists(LIB_MESH) else None
    if not force and os.path.exists(LIB_MESH) and os.path.getmtime(LIB_MESH) >= max(
            os.path.getmtime(hdr), os.path.getmtime(shim), os.path.getmtime(__file__)):
        return LIB_MESH
    os.makedirs(OUT, exist_ok=True)
    text = open(hdr).read()
    inc = os.path.join(OUT, "ref_mesh_functions.inc")
    with open(inc, "w") as fh:
        fh.write(_cut(text, "decompose3d(int x)", with_template=False))
        fh.write(_cut(text, "compute_cartesian_indices(", with_template=False))
        fh.write(_cut(open(os.path.join(REF, "common", "permute.hpp")).read(), "void reorder_dofmap(", with_template=False))
        fh.write(_cut(open(os.path.join(REF, "common", "precompute.hpp")).read(), "void dot(const U& A"))
    # the statements between "// Temporal parameters" and the first print of demo/cpu_planar3d/main.cpp
    demo = open(os.path.join(REF, "demo", "cpu_planar3d", "main.cpp")).read().split("\n")
    a = next(i for i, l in enumerate(demo) if "// Temporal parameters" in l)
    b = next(i for i in range(a, len(demo)) if "if (rank == 0)" in demo[i])
    inc2 = os.path.join(OUT, "ref_demo_params.inc")
    with open(inc2, "w") as fh:
        fh.write("\n".join(demo[a + 1:b]) + "\n")
    try:
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I", OUT, shim, "-o", LIB_MESH]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("building oracle/_ref/libwfref_mesh.so failed")
    finally:
        os.remove(inc)
        os.remove(inc2)
    return LIB_MESH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build_cpu(force="--force" in sys.argv))
    print(build_mesh(force="--force" in sys.argv))
