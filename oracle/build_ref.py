"""Builds oracle/_ref/libwfref_cuda.so from the REFERENCE's own CUDA sources (read in place under
/root/reference, never copied) plus the forwarding shim oracle/ref_cuda_shim.cu.

TEST INFRASTRUCTURE ONLY.  What compiles of the reference without its un-vendored dependencies
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


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
